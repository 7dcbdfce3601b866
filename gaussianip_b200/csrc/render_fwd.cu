// Stage 3: per-tile front-to-back alpha blending.  Replaces renderCUDA (forward) of the
// external operator ("Forward blend").
//
// One CTA per 16x16 tile, 8 warps; warp w owns an 8x4 pixel block, and inside the warp lanes 0-15 / 16-31 own its
// left / right 4x4 pixels (lane_px / lane_py).  The two HALVES take their own hits: a ~3.7 px splat overlaps 1.4
// of the two 4x4 blocks on average, so letting each half evaluate a different Gaussian in the same instruction
// stream cuts the evaluation iterations of a chunk from |hits(A) u hits(B)| to max(|hits(A)|, |hits(B)|).  The warps of a CTA are fully INDEPENDENT (no __syncthreads): each walks
// the tile's sorted instance list in chunks of 32, gathering point_list -> 48 B Geom record
// (L2-resident) into its own shared-memory ring with cp.async, STAGES chunks ahead of the
// blend.  Lane l first tests whether instance l of the chunk can reach any of the warp's 32
// pixels (conservative extent test, see preprocess.cu); the ballot gives the hit list and only
// hits are evaluated, in list order, so results are identical to evaluating everything.  A
// warp stops as soon as all of its pixels are saturated.
// Hit records (RECORD): whenever at least one of the warp's pixels BLENDS a Gaussian, the warp appends
// {Gaussian id, mask of the blending lanes} to its own record list in `saved` (GsbLayout.off_hits): per hit one
// ballot and one shared-memory store of the mask (slot = the instance's position in the chunk), per chunk one
// compaction of the non-zero masks into consecutive records.  The backward blend replays these lists back to
// front and never re-tests or culls anything (render_bwd.cu).
// Measured alternatives (profiles/r1_experiments.md): block-synchronous 256-instance batches lost
// 44 % of issue slots to CTA barriers (r1a); deeper private rings are SLOWER (4 stages 241 us, 8
// stages 318 us vs 232 us at 2: shared memory is taken from L1, which serves the 8 warps' re-reads
// of the same records); one producer warp feeding a ring shared by the 8 consumers was slower too
// (257-297 us: a single warp's gather latency cannot feed eight consumers).
#include <atomic>

#include "gsb_common.cuh"

namespace gsb {

namespace {

constexpr int WARPS = 8;
#ifndef GSB_FWD_HB
#define GSB_FWD_HB 2
#endif
constexpr int HB = GSB_FWD_HB;   // hits evaluated together (ILP)
#ifndef GSB_FWD_STAGES
#define GSB_FWD_STAGES 2
#endif
constexpr int STAGES = GSB_FWD_STAGES;   // chunks in flight per warp (power of two)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <bool RECORD>
__global__ void __launch_bounds__(WARPS * 32)
render_fwd_kernel(View v, const Geom* __restrict__ geom, const uint32_t* __restrict__ point_list,
                  const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                  float* __restrict__ out_color,
                  float* __restrict__ out_depth, float* __restrict__ out_alpha,
                  uint32_t* __restrict__ n_contrib, float* __restrict__ final_T,
                  uint2* __restrict__ hits, uint32_t* __restrict__ hit_count) {
  extern __shared__ float4 smem_dyn[];             // 12 KB per stage per CTA
  float4 (*s_rec)[STAGES][3][32] = reinterpret_cast<float4 (*)[STAGES][3][32]>(smem_dyn);
  // RECORD: per warp and half, the blend masks of the current chunk's 32 instances
  uint32_t (*s_mask)[2][32] = reinterpret_cast<uint32_t (*)[2][32]>(smem_dyn + WARPS * STAGES * 3 * 32);

  const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles are launched first
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4;                      // 0: left 4x4 pixels of the warp's 8x4 block, 1: right 4x4
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int lx = lane_px(lane), ly = lane_py(lane);
  const int pix_x = tx * TILE_X + wx + lx;
  const int pix_y = ty * TILE_Y + wy + ly;
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  // Cull rectangles, one per half = bounding box of the half's pixels that are still accumulating.  They start as
  // the whole 4x4 blocks and shrink as pixels saturate, so the long-running warps (a few unsaturated silhouette
  // pixels) stop paying for splats that only reach finished pixels.  Every lane keeps both: lane l tests
  // instance l of a chunk against the two halves.
  float cxA = (float)(tx * TILE_X + wx) + 1.5f, cxB = cxA + 4.0f, cyA = (float)(ty * TILE_Y + wy) + 1.5f, cyB = cyA;
  float hwxA = 1.5f, hwyA = 1.5f, hwxB = 1.5f, hwyB = 1.5f;
  uint32_t alive_prev = 0xffffffffu;

  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const int chunks = (n + 31) >> 5;
  const uint32_t* pl = point_list + range.x;

  // A pixel is "done" (saturated, or outside the image) exactly when T == 0: a live pixel always has
  // T >= T_MIN, and with T = 0 every later test T*(1-alpha) >= T_MIN fails by itself, so the hit loop needs
  // no separate flag.  T_live follows T while the pixel accumulates and keeps the last value afterwards
  // (the transmittance the reference reports as final_T).
  float T = inside ? 1.0f : 0.0f, T_live = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dz = 0.f, A = 0.f;
  uint32_t last = 0;
  // hit records of this warp: capacity n (an instance yields at most one record per warp)
  uint2* rec_base = RECORD ? hits + ((size_t)range.x * WARPS + (size_t)warp * (size_t)n) : nullptr;
  int rec_n = 0;

  if (chunks > 0 && __any_sync(0xffffffffu, T != 0.0f)) {
    float4 (*ring)[3][32] = s_rec[warp];
    auto issue = [&](int c, uint32_t gid) {      // stage chunk c (lane's instance) into the ring
      if (c < chunks) {
        if (c * 32 + lane < n) {
          const float4* src = reinterpret_cast<const float4*>(geom + gid);
          float4 (*st)[32] = ring[c & (STAGES - 1)];
          cp_async16(&st[0][lane], src);
          cp_async16(&st[1][lane], src + 1);
          cp_async16(&st[2][lane], src + 2);
        }
      }
      cp_async_commit();                         // always commit: keeps the group count uniform
    };
    auto fetch_gid = [&](int c) -> uint32_t {
      const int e = c * 32 + lane;
      return (c < chunks && e < n) ? pl[e] : 0u;
    };
    // prologue: STAGES-1 chunks in flight, ids of the next one in a register
    static_assert(!RECORD || STAGES == 2, "the record path keeps the ids of exactly one chunk in flight");
    uint32_t gid_cur = fetch_gid(0);              // this lane's instance of the chunk being blended (RECORD)
#pragma unroll
    for (int c = 0; c < STAGES - 1; ++c) issue(c, c == 0 ? gid_cur : fetch_gid(c));
    uint32_t gid_next = fetch_gid(STAGES - 1);

    for (int c = 0; c < chunks; ++c) {
      const uint32_t gid_issued = gid_next;
      issue(c + STAGES - 1, gid_next);
      gid_next = fetch_gid(c + STAGES);
      if (RECORD) { s_mask[warp][0][lane] = 0u; s_mask[warp][1][lane] = 0u; }
      cp_async_wait<STAGES - 1>();               // chunk c has landed (for this lane)
      __syncwarp();                              // ... and for every lane of the warp
      float4 (*st)[32] = ring[c & (STAGES - 1)];
      const int e = c * 32 + lane;
      const bool done = T == 0.0f;
      const uint32_t alive = __ballot_sync(0xffffffffu, !done);
      if (alive != alive_prev) {
        const uint32_t changed = alive ^ alive_prev;
        alive_prev = alive;
        const int hx = lane & 3;                 // x inside the half
        if (changed & 0x0000ffffu) {
          const bool on = !done && half == 0;
          const int x0 = __reduce_min_sync(0xffffffffu, on ? hx : 64), x1 = __reduce_max_sync(0xffffffffu, on ? hx : -1);
          const int y0 = __reduce_min_sync(0xffffffffu, on ? ly : 64), y1 = __reduce_max_sync(0xffffffffu, on ? ly : -1);
          hwxA = 0.5f * (float)(x1 - x0); hwyA = 0.5f * (float)(y1 - y0);
          cxA = (float)(tx * TILE_X + wx + x0) + hwxA; cyA = (float)(ty * TILE_Y + wy + y0) + hwyA;
        }
        if (changed & 0xffff0000u) {
          const bool on = !done && half == 1;
          const int x0 = __reduce_min_sync(0xffffffffu, on ? hx : 64), x1 = __reduce_max_sync(0xffffffffu, on ? hx : -1);
          const int y0 = __reduce_min_sync(0xffffffffu, on ? ly : 64), y1 = __reduce_max_sync(0xffffffffu, on ? ly : -1);
          hwxB = 0.5f * (float)(x1 - x0); hwyB = 0.5f * (float)(y1 - y0);
          cxB = (float)(tx * TILE_X + wx + 4 + x0) + hwxB; cyB = (float)(ty * TILE_Y + wy + y0) + hwyB;
        }
      }
      bool hitA = false, hitB = false;
      if (e < n) {
        const float4 a = st[0][lane];
        hitA = (fabsf(a.x - cxA) <= a.z + hwxA) && (fabsf(a.y - cyA) <= a.w + hwyA);
        hitB = (fabsf(a.x - cxB) <= a.z + hwxB) && (fabsf(a.y - cyB) <= a.w + hwyB);
      }
      // a half without live pixels takes no hits (its rectangle is empty: negative half-widths can still pass
      // the test for degenerate splats of infinite extent)
      // every lane keeps the pending hits of ITS half (a half without live pixels takes none: its rectangle is
      // empty, but negative half-widths can still pass the test for degenerate splats of infinite extent)
      const uint32_t maskA = __ballot_sync(0xffffffffu, hitA), maskB = __ballot_sync(0xffffffffu, hitB);
      uint32_t pend = half ? maskB : maskA;
      if (!(alive & (half ? 0xffff0000u : 0x0000ffffu))) pend = 0u;
      // The two halves take their own hits, in list order, HB at a time each: an iteration evaluates up to
      // 2 x HB (half, Gaussian) pairs with one warp instruction stream.  Alphas (LDS, conic, ex2) of the HB hits
      // are independent and overlap; only the short transmittance chain is applied in order.
      while (__any_sync(0xffffffffu, pend != 0u)) {
        int k[HB];
#pragma unroll
        for (int i = 0; i < HB; ++i) {
          k[i] = __ffs(pend) - 1;               // -1 when this half has no hit left
          pend &= pend - 1;
        }
        float al[HB];
        float4 ff[HB];
#pragma unroll
        for (int i = 0; i < HB; ++i) {
          al[i] = 0.0f;
          if (k[i] >= 0) {
            const float4 a = st[0][k[i]];     // x, y, -, -
            const float4 q = st[1][k[i]];     // pre-scaled conic (qa, qb, qc), opacity
            ff[i] = st[2][k[i]];              // depth, r, g, b
            const float dx = a.x - pxf, dy = a.y - pyf;
            const float e2 = gauss_exponent2(q.x, q.y, q.z, dx, dy);      // log2 of the Gaussian weight
            al[i] = e2 <= 0.0f ? fminf(ALPHA_CAP, q.w * exp2_blend(e2)) : 0.0f;
          }
        }
#pragma unroll
        for (int i = 0; i < HB; ++i) {
          bool ok = false;
          if (k[i] >= 0 && al[i] >= ALPHA_MIN) {
            const float test_T = T * (1.0f - al[i]);
            ok = test_T >= T_MIN;
            if (ok) {
              const float w = al[i] * T;
              C0 += ff[i].y * w; C1 += ff[i].z * w; C2 += ff[i].w * w;
              Dz += ff[i].x * w; A += w;
              T_live = test_T;
              last = (uint32_t)(c * 32 + k[i] + 1);
            }
            T = ok ? test_T : 0.0f;       // a saturating splat (or a finished pixel) leaves T at 0
          }
          if (RECORD) {
            // bits 0-15: the left pixels that blended the left half's Gaussian; bits 16-31: the right half's
            const uint32_t vb = __ballot_sync(0xffffffffu, ok);
            if ((lane & 15) == 0 && k[i] >= 0) s_mask[warp][half][k[i]] = half ? (vb & 0xffff0000u) : (vb & 0x0000ffffu);
          }
        }
      }
      if (RECORD) {
        // append this chunk's records: instance l of the chunk was blended by the pixels in its two masks
        __syncwarp();
        const uint32_t m = s_mask[warp][0][lane] | s_mask[warp][1][lane];
        const uint32_t nz = __ballot_sync(0xffffffffu, m != 0u);
        if (m) rec_base[rec_n + __popc(nz & ((1u << lane) - 1u))] = make_uint2(gid_cur, m);
        rec_n += __popc(nz);
        gid_cur = gid_issued;
      }
      if (__all_sync(0xffffffffu, T == 0.0f)) break;
      __syncwarp();                              // ring slot c is free before it is refilled
    }
    cp_async_wait<0>();
  }
  if (RECORD && lane == 0) hit_count[tile * WARPS + warp] = (uint32_t)rec_n;

  if (inside) {
    T = T_live;
    const size_t hw = (size_t)v.H * v.W;
    const size_t pix = (size_t)pix_y * v.W + pix_x;
    out_color[pix] = C0 + T * v.bg[0];
    out_color[hw + pix] = C1 + T * v.bg[1];
    out_color[2 * hw + pix] = C2 + T * v.bg[2];
    out_depth[pix] = Dz;
    out_alpha[pix] = A;
    n_contrib[pix] = last;
    final_T[pix] = T;
  }
}

}  // namespace

int launch_render_fwd(const View& v, const Geom* geom, const uint32_t* point_list,
                      const uint2* ranges, const uint32_t* tile_order, float* color, float* depth, float* alpha,
                      uint32_t* n_contrib, float* final_T, uint2* hits, uint32_t* hit_count, bool debug,
                      cudaStream_t st) {
  const int T = v.gx * v.gy;
  if (T == 0) return GSB_OK;
#ifndef GSB_FWD_SMEM_PAD
#define GSB_FWD_SMEM_PAD 0
#endif
  constexpr size_t smem = (size_t)WARPS * STAGES * 3 * 32 * sizeof(float4) + (size_t)WARPS * 2 * 32 * sizeof(uint32_t) + GSB_FWD_SMEM_PAD;
  static std::atomic<unsigned long long> configured{0};   // bit per device: the attribute is per device
  int dev = 0;
  GSB_CUDA(cudaGetDevice(&dev));
  if (!(configured.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  if (hits && hit_count)
    render_fwd_kernel<true><<<T, WARPS * 32, smem, st>>>(v, geom, point_list, ranges, tile_order, color, depth,
                                                         alpha, n_contrib, final_T, hits, hit_count);
  else
    render_fwd_kernel<false><<<T, WARPS * 32, smem, st>>>(v, geom, point_list, ranges, tile_order, color, depth,
                                                          alpha, n_contrib, final_T, nullptr, nullptr);
  GSB_POST_LAUNCH(debug, st, "render_fwd_kernel");
  return GSB_OK;
}

}  // namespace gsb
