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
#include <cstring>

#include <cuda.h>            // CUtensorMap (type only: the encoder is fetched with cudaGetDriverEntryPoint)
#include <cudaTypedefs.h>

#include "gsb_common.cuh"

namespace gsb {

namespace {

constexpr int WARPS = 8;
// GSB_FWD_NOBRANCH: 0 = per-slot `if (hit)` branches in the hit loop (rounds 1-2), 1 = the alpha evaluation of a
// slot is branch-free, 2 (default) = the blend of a slot, the per-chunk cull test and the record store as well.  The
// divergent regions cost far more than their instructions: r2y measured 222.8 -> 210.8 -> 195.3 us per view for
// 0 / 1 / 2 with bit-identical results.
#ifndef GSB_FWD_NOBRANCH
#define GSB_FWD_NOBRANCH 2
#endif
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

// ---- TMA row gather (sm_100 cp.async.bulk.tensor ... tile::gather4) ------------------------------------------
// G4 variant of the kernel below: Geom is described to the TMA unit as a 2-D fp32 tensor [P rows][12], box {16, 1}
// (the 4 floats past a row are out of bounds and arrive as zeros, which gives the rows a 64-byte pitch in shared
// memory: row k of a stage sits at k * 64, one multiply like the cp.async layout).  Eight lanes of a warp issue one
// gather4 each (4 rows = 256 bytes, 128-byte aligned destination) instead of 3 x LDGSTS.128 on every lane;
// completion is a per-(warp, stage) mbarrier with a transaction count.  scripts/micro/tma_gather4_probe.cu
// measured the instruction in isolation (1.45-2x the LDGSTS row rate from L2).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tm, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}
constexpr int G4_ROW_FLOAT4 = 4;      // 64-byte row pitch of a gather4 stage
constexpr size_t G4_STAGE_BYTES = 32 * G4_ROW_FLOAT4 * sizeof(float4);

#ifndef GSB_FWD_MINB
#define GSB_FWD_MINB 1
#endif
template <bool RECORD, bool G4>
__global__ void __launch_bounds__(WARPS * 32, GSB_FWD_MINB)
render_fwd_kernel(View v, const __grid_constant__ CUtensorMap geom_map, const Geom* __restrict__ geom,
                  const uint32_t* __restrict__ point_list,
                  const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                  float* __restrict__ out_color,
                  float* __restrict__ out_depth, float* __restrict__ out_alpha,
                  uint32_t* __restrict__ n_contrib, float* __restrict__ final_T,
                  uint2* __restrict__ hits, uint32_t* __restrict__ hit_count) {
  extern __shared__ __align__(128) float4 smem_dyn[];   // cp.async: 12 KB per stage per CTA; gather4: 16 KB
  float4 (*s_rec)[STAGES][3][32] = reinterpret_cast<float4 (*)[STAGES][3][32]>(smem_dyn);
  constexpr size_t RING_FLOAT4 = G4 ? (size_t)WARPS * STAGES * 32 * G4_ROW_FLOAT4 : (size_t)WARPS * STAGES * 3 * 32;
  // RECORD: per warp and half, the blend masks of the current chunk's 32 instances
  uint32_t (*s_mask)[2][32] = reinterpret_cast<uint32_t (*)[2][32]>(smem_dyn + RING_FLOAT4);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_dyn + RING_FLOAT4 + WARPS * 2 * 32 / 4);   // [WARPS][STAGES] (G4)

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
    float4* ring4 = smem_dyn + (size_t)warp * STAGES * 32 * G4_ROW_FLOAT4;      // G4: [STAGES][32 rows][4 float4]
    uint64_t* bars = s_bar + warp * STAGES;
    if (G4) {
      if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncwarp();
    }
    int in_flight = -1;                            // G4: last chunk handed to the TMA unit
    auto issue = [&](int c, uint32_t gid) {      // stage chunk c (lane's instance) into the ring
      if (G4) {
        if (c < chunks) {                          // warp-uniform; lanes past the end of the list fetch row 0
          const int id1 = (int)__shfl_down_sync(0xffffffffu, gid, 1), id2 = (int)__shfl_down_sync(0xffffffffu, gid, 2),
                    id3 = (int)__shfl_down_sync(0xffffffffu, gid, 3);
          if (lane == 0) mbar_expect_tx(&bars[c & (STAGES - 1)], (uint32_t)G4_STAGE_BYTES);
          if ((lane & 3) == 0) {
            // the slot was read with ordinary loads until the __syncwarp that ended the previous round
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tma_gather4(ring4 + ((size_t)(c & (STAGES - 1)) * 32 + lane) * G4_ROW_FLOAT4, &geom_map, (int)gid, id1, id2,
                        id3, &bars[c & (STAGES - 1)]);
          }
          in_flight = c;
        }
        return;
      }
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
      if (G4) {
        mbar_wait(&bars[c & (STAGES - 1)], (uint32_t)((c / STAGES) & 1));     // chunk c has landed
      } else {
        cp_async_wait<STAGES - 1>();             // chunk c has landed (for this lane)
      }
      __syncwarp();                              // ... and for every lane of the warp
      float4 (*st)[32] = ring[c & (STAGES - 1)];
      const float4* st4 = ring4 + (size_t)(c & (STAGES - 1)) * 32 * G4_ROW_FLOAT4;
      // record r of the stage: three float4 (cp.async: three planes of 32; gather4: one 64-byte row)
      auto rec0 = [&](int r) -> float4 { return G4 ? st4[r * G4_ROW_FLOAT4] : st[0][r]; };
      auto rec1 = [&](int r) -> float4 { return G4 ? st4[r * G4_ROW_FLOAT4 + 1] : st[1][r]; };
      auto rec2 = [&](int r) -> float4 { return G4 ? st4[r * G4_ROW_FLOAT4 + 2] : st[2][r]; };
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
#if GSB_FWD_NOBRANCH >= 2
      {
        // branch-free: lanes past the end of the list test whatever their (unwritten) slot holds and are masked
        // (the empty asm keeps the compiler from sinking the load back under a branch on `in_list`)
        float4 a = rec0(lane);
        asm volatile("" : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w));
        const bool in_list = e < n;
        hitA = in_list & (fabsf(a.x - cxA) <= a.z + hwxA) & (fabsf(a.y - cyA) <= a.w + hwyA);
        hitB = in_list & (fabsf(a.x - cxB) <= a.z + hwxB) & (fabsf(a.y - cyB) <= a.w + hwyB);
      }
#else
      if (e < n) {
        const float4 a = rec0(lane);
        hitA = (fabsf(a.x - cxA) <= a.z + hwxA) && (fabsf(a.y - cyA) <= a.w + hwyA);
        hitB = (fabsf(a.x - cxB) <= a.z + hwxB) && (fabsf(a.y - cyB) <= a.w + hwyB);
      }
#endif
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
#if GSB_FWD_NOBRANCH
          {
            // branch-free: a half without a hit in this slot evaluates record 0 and discards the result
            const int kk = k[i] < 0 ? 0 : k[i];
            const float4 a = rec0(kk);        // x, y, -, -
            const float4 q = rec1(kk);        // pre-scaled conic (qa, qb, qc), opacity
            ff[i] = rec2(kk);                 // depth, r, g, b
            const float dx = a.x - pxf, dy = a.y - pyf;
            const float e2 = gauss_exponent2(q.x, q.y, q.z, dx, dy);      // log2 of the Gaussian weight
            al[i] = (k[i] >= 0 && e2 <= 0.0f) ? fminf(ALPHA_CAP, q.w * exp2_blend(e2)) : 0.0f;
          }
#else
          al[i] = 0.0f;
          if (k[i] >= 0) {
            const float4 a = rec0(k[i]);      // x, y, -, -
            const float4 q = rec1(k[i]);      // pre-scaled conic (qa, qb, qc), opacity
            ff[i] = rec2(k[i]);               // depth, r, g, b
            const float dx = a.x - pxf, dy = a.y - pyf;
            const float e2 = gauss_exponent2(q.x, q.y, q.z, dx, dy);      // log2 of the Gaussian weight
            al[i] = e2 <= 0.0f ? fminf(ALPHA_CAP, q.w * exp2_blend(e2)) : 0.0f;
          }
#endif
        }
#pragma unroll
        for (int i = 0; i < HB; ++i) {
#if GSB_FWD_NOBRANCH >= 2
          // branch-free blend: al is 0 for a half without a hit in this slot, so `cand` covers k < 0 as well
          const bool cand = al[i] >= ALPHA_MIN;
          const float test_T = T * (1.0f - al[i]);
          const bool ok = cand && test_T >= T_MIN;
          const float w = al[i] * T;
          C0 = ok ? fmaf(ff[i].y, w, C0) : C0; C1 = ok ? fmaf(ff[i].z, w, C1) : C1; C2 = ok ? fmaf(ff[i].w, w, C2) : C2;
          Dz = ok ? fmaf(ff[i].x, w, Dz) : Dz; A = ok ? A + w : A;
          T_live = ok ? test_T : T_live;
          last = ok ? (uint32_t)(c * 32 + k[i] + 1) : last;
          T = cand ? (ok ? test_T : 0.0f) : T;     // a saturating splat (or a finished pixel) leaves T at 0
#else
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
#endif
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
#if GSB_FWD_NOBRANCH >= 2
        {
          // predicated store: the slot address is computed by every lane so that no divergent region is needed
          uint2* slot = rec_base + (rec_n + __popc(nz & ((1u << lane) - 1u)));
          asm volatile("" : "+l"(slot));
          if (m) *slot = make_uint2(gid_cur, m);
        }
#else
        if (m) rec_base[rec_n + __popc(nz & ((1u << lane) - 1u))] = make_uint2(gid_cur, m);
#endif
        rec_n += __popc(nz);
        gid_cur = gid_issued;
      }
      if (__all_sync(0xffffffffu, T == 0.0f)) {
        // a warp that stops early still owns the chunk in flight: the TMA unit must not write into shared memory
        // of a CTA that has exited
        if (G4 && in_flight > c) mbar_wait(&bars[in_flight & (STAGES - 1)], (uint32_t)((in_flight / STAGES) & 1));
        break;
      }
      __syncwarp();                              // ring slot c is free before it is refilled
    }
    if (!G4) cp_async_wait<0>();
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


// ---- forward blend with a CTA-wide pre-cull --------------------------------------------------------------------
// In the kernel above every one of a tile's 8 warps walks the WHOLE tile list: it gathers the 48-byte record of every
// instance and tests it against its own 8x4 pixels, although only ~19 % of the instances reach a given warp (ncu:
// ~110 warp instructions per 32-instance chunk and warp, 27 % of the kernel; every record is fetched 8 times per CTA).
// Here the CTA first classifies the list ONCE (phase 0: the 8 warps share the chunks; lane = instance loads only the
// first 16 bytes of the record and writes an 8-bit mask of the warp regions the splat's conservative extent box
// overlaps — the same test as below with the static 8x4 rectangles, so nothing that could blend is dropped).  After
// one __syncthreads each warp scans the byte masks (a load, a ballot and a count per 32 instances), queues the
// positions of ITS instances and gathers / tests / blends dense chunks of 32 of them exactly like the kernel above
// (same per-half dynamic rectangles, same arithmetic, same list order, so images, contributor counts and hit records
// are identical).  MEASURED (r2v / r2w, profiles/r2_experiments.md): 139.5 M warp instructions per launch instead of
// 149.2 M, but IPC 2.5 instead of 2.7 (the scan's id load sits on the warp's critical path): 224.5 us against 223.2 us
// per view — no gain, so the kernel above stays the default.  Staging the ids in shared memory as well (+32 KB per
// CTA, taken from L1) and four 4x2 hit streams per warp instead of two 4x4 ones were both slower (249 / 251 us).
constexpr int PC_CAP = 8192;           // instances of a tile that get a mask byte; the rest of a longer list is taken by every warp
constexpr int PC_QUEUE = 64;           // queued (Gaussian id, list position) pairs per warp (power of two, >= 63)
#ifndef GSB_PC_UNITS
#define GSB_PC_UNITS 2
#endif
// UNITS = hit streams per warp: 2 = the two 4x4 halves (as above), 4 = four 4x2 quarters (lanes 8u .. 8u+7), each
// taking its own hits from a chunk: fewer iterations when a splat reaches only part of a half.
constexpr int PC_UNITS = GSB_PC_UNITS;
static_assert(PC_UNITS == 2 || PC_UNITS == 4, "halves or quarters");
constexpr size_t PC_SMEM_BYTES = (size_t)WARPS * STAGES * 3 * 32 * sizeof(float4) + (size_t)WARPS * PC_UNITS * 32 * sizeof(uint32_t) +
                                 (size_t)WARPS * PC_QUEUE * sizeof(uint2) + PC_CAP;

template <bool RECORD>
__global__ void __launch_bounds__(WARPS * 32, GSB_FWD_MINB)
render_fwd_precull_kernel(View v, const Geom* __restrict__ geom, const uint32_t* __restrict__ point_list,
                          const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                          float* __restrict__ out_color, float* __restrict__ out_depth,
                          float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib,
                          float* __restrict__ final_T, uint2* __restrict__ hits, uint32_t* __restrict__ hit_count) {
  constexpr int U = PC_UNITS, LU = 32 / U;          // lanes per unit
  extern __shared__ __align__(128) float4 smem_dyn[];
  float4 (*s_rec)[STAGES][3][32] = reinterpret_cast<float4 (*)[STAGES][3][32]>(smem_dyn);
  uint32_t (*s_mask)[U][32] = reinterpret_cast<uint32_t (*)[U][32]>(smem_dyn + WARPS * STAGES * 3 * 32);
  uint2 (*s_q)[PC_QUEUE] = reinterpret_cast<uint2 (*)[PC_QUEUE]>(s_mask + WARPS);
  uint8_t* s_wm = reinterpret_cast<uint8_t*>(s_q + WARPS);

  const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles are launched first
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int unit = lane / LU;
  const uint32_t unit_lanes = (U == 4 ? 0xffu : 0xffffu) << (unit * LU);
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int lx = lane_px(lane), ly = lane_py(lane);
  const int pix_x = tx * TILE_X + wx + lx;
  const int pix_y = ty * TILE_Y + wy + ly;
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  // Cull rectangles, one per unit = bounding box of the unit's pixels that are still accumulating (pixel centres).
  // Unit u of 2: x0 = 4u, y0 = 0, 4x4.  Unit u of 4 (lanes 8u..8u+7): x0 = 4 (u >> 1), y0 = 2 (u & 1), 4x2.
  const float bx = (float)(tx * TILE_X + wx), by = (float)(ty * TILE_Y + wy);
  float cx[U], cy[U], hwx[U], hwy[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int x0 = U == 4 ? 4 * (u >> 1) : 4 * u, y0 = U == 4 ? 2 * (u & 1) : 0;
    hwx[u] = 1.5f; hwy[u] = U == 4 ? 0.5f : 1.5f;
    cx[u] = bx + (float)x0 + hwx[u]; cy[u] = by + (float)y0 + hwy[u];
  }
  uint32_t alive_prev = 0xffffffffu;

  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const uint32_t* pl = point_list + range.x;

  // ---- phase 0: which warp regions can each instance reach (once per CTA) ----
  {
    const int n_m = n < PC_CAP ? n : PC_CAP;
    const float tx0 = (float)(tx * TILE_X), ty0 = (float)(ty * TILE_Y);
#pragma unroll 2
    for (int e = warp * 32 + lane; e < n_m; e += WARPS * 32) {
      const uint32_t gid = pl[e];
      const float4 a = *reinterpret_cast<const float4*>(geom + gid);      // x, y, extent x, extent y
      // warp region (c, r): pixel centres x in [tx0 + 8c, +7], y in [ty0 + 4r, +3]
      const float ex = a.z + 3.5f, ey = a.w + 1.5f;
      const uint32_t c0 = fabsf(a.x - (tx0 + 3.5f)) <= ex, c1 = fabsf(a.x - (tx0 + 11.5f)) <= ex;
      const uint32_t cm = c0 | (c1 << 1);                                    // warps 0/1 of a row
      uint32_t m = 0u;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (fabsf(a.y - (ty0 + 1.5f + 4.0f * (float)r)) <= ey) m |= cm << (2 * r);
      s_wm[e] = (uint8_t)m;
    }
  }
  __syncthreads();

  float T = inside ? 1.0f : 0.0f, T_live = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dz = 0.f, A = 0.f;
  uint32_t last = 0;
  uint2* rec_base = RECORD ? hits + ((size_t)range.x * WARPS + (size_t)warp * (size_t)n) : nullptr;
  int rec_n = 0;
  const uint32_t lt = (1u << lane) - 1u;

  if (n > 0 && __any_sync(0xffffffffu, T != 0.0f)) {
    float4 (*ring)[3][32] = s_rec[warp];
    uint2* q = s_q[warp];
    int scan = 0, q_head = 0, q_cnt = 0;
    // queue this warp's instances of whole 32-instance chunks until a dense chunk is available
    auto scan_more = [&]() {
      while (q_cnt < 32 && scan < n) {
        const int e = scan + lane;
        uint32_t m = 0u;
        if (e < n) m = e < PC_CAP ? (uint32_t)s_wm[e] : 0xffu;
        const bool mine = (m >> warp) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, mine);
        if (bal) {
          if (mine) q[(q_head + q_cnt + __popc(bal & lt)) & (PC_QUEUE - 1)] = make_uint2(pl[e], (uint32_t)e);
          q_cnt += __popc(bal);
        }
        scan += 32;
      }
      __syncwarp();
    };
    // take up to 32 queued instances: returns how many; lane l < count receives entry l
    auto pop = [&](uint2& ent) -> int {
      const int cnt = q_cnt < 32 ? q_cnt : 32;
      ent = make_uint2(0u, 0u);
      if (lane < cnt) ent = q[(q_head + lane) & (PC_QUEUE - 1)];
      q_head = (q_head + cnt) & (PC_QUEUE - 1);
      q_cnt -= cnt;
      __syncwarp();                               // the entries are read before scan_more overwrites the slots
      return cnt;
    };
    auto issue = [&](int slot, int cnt, uint32_t gid) {      // stage a dense chunk (lane's instance) into the ring
      if (lane < cnt) {
        const float4* src = reinterpret_cast<const float4*>(geom + gid);
        float4 (*st)[32] = ring[slot];
        cp_async16(&st[0][lane], src);
        cp_async16(&st[1][lane], src + 1);
        cp_async16(&st[2][lane], src + 2);
      }
      cp_async_commit();                         // always commit: keeps the group count uniform
    };
    static_assert(STAGES == 2, "one dense chunk in flight while one is blended");
    uint2 cur, nxt;
    scan_more();
    int cnt_cur = pop(cur);
    issue(0, cnt_cur, cur.x);
    for (int c = 0; cnt_cur > 0; ++c) {
      scan_more();
      const int cnt_nxt = pop(nxt);
      issue((c + 1) & 1, cnt_nxt, nxt.x);
      if (RECORD) {
#pragma unroll
        for (int u = 0; u < U; ++u) s_mask[warp][u][lane] = 0u;
      }
      cp_async_wait<1>();                        // dense chunk c has landed (for this lane)
      __syncwarp();                              // ... and for every lane of the warp
      float4 (*st)[32] = ring[c & 1];
      const bool done = T == 0.0f;
      const uint32_t alive = __ballot_sync(0xffffffffu, !done);
      if (alive != alive_prev) {
        const uint32_t changed = alive ^ alive_prev;
        alive_prev = alive;
        const int hx = lane & 3;                             // x inside the unit
        const int hy = U == 4 ? ((lane >> 2) & 1) : ly;      // y inside the unit
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t um = (U == 4 ? 0xffu : 0xffffu) << (u * LU);
          if (changed & um) {                                // warp-uniform
            const bool on = !done && unit == u;
            const int x0 = __reduce_min_sync(0xffffffffu, on ? hx : 64), x1 = __reduce_max_sync(0xffffffffu, on ? hx : -1);
            const int y0 = __reduce_min_sync(0xffffffffu, on ? hy : 64), y1 = __reduce_max_sync(0xffffffffu, on ? hy : -1);
            const int ux = U == 4 ? 4 * (u >> 1) : 4 * u, uy = U == 4 ? 2 * (u & 1) : 0;
            hwx[u] = 0.5f * (float)(x1 - x0); hwy[u] = 0.5f * (float)(y1 - y0);
            cx[u] = bx + (float)(ux + x0) + hwx[u]; cy[u] = by + (float)(uy + y0) + hwy[u];
          }
        }
      }
      // every lane keeps the pending hits of ITS unit (a unit without live pixels takes none: its rectangle is
      // empty, but negative half-widths can still pass the test for degenerate splats of infinite extent)
      uint32_t pend = 0u;
      {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool have = lane < cnt_cur;
        if (have) a = st[0][lane];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool h = have && (fabsf(a.x - cx[u]) <= a.z + hwx[u]) && (fabsf(a.y - cy[u]) <= a.w + hwy[u]);
          const uint32_t mk = __ballot_sync(0xffffffffu, h);
          if (unit == u) pend = mk;
        }
      }
      if (!(alive & unit_lanes)) pend = 0u;
      while (__any_sync(0xffffffffu, pend != 0u)) {
        int k[HB];
        uint32_t posk[HB];
#pragma unroll
        for (int i = 0; i < HB; ++i) {
          k[i] = __ffs(pend) - 1;               // -1 when this unit has no hit left
          pend &= pend - 1;
          posk[i] = __shfl_sync(0xffffffffu, cur.y, k[i] & 31);   // list position of that instance
        }
        float al[HB];
        float4 ff[HB];
#pragma unroll
        for (int i = 0; i < HB; ++i) {
          al[i] = 0.0f;
          if (k[i] >= 0) {
            const float4 a = st[0][k[i]];     // x, y, -, -
            const float4 qq = st[1][k[i]];    // pre-scaled conic (qa, qb, qc), opacity
            ff[i] = st[2][k[i]];              // depth, r, g, b
            const float dx = a.x - pxf, dy = a.y - pyf;
            const float e2 = gauss_exponent2(qq.x, qq.y, qq.z, dx, dy);      // log2 of the Gaussian weight
            al[i] = e2 <= 0.0f ? fminf(ALPHA_CAP, qq.w * exp2_blend(e2)) : 0.0f;
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
              last = posk[i] + 1u;
            }
            T = ok ? test_T : 0.0f;       // a saturating splat (or a finished pixel) leaves T at 0
          }
          if (RECORD) {
            // the unit's bits of the vote: the pixels of this unit that blended its Gaussian
            const uint32_t vb = __ballot_sync(0xffffffffu, ok);
            if ((lane & (LU - 1)) == 0 && k[i] >= 0) s_mask[warp][unit][k[i]] = vb & unit_lanes;
          }
        }
      }
      if (RECORD) {
        __syncwarp();
        uint32_t m = 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) m |= s_mask[warp][u][lane];
        const uint32_t nz = __ballot_sync(0xffffffffu, m != 0u);
        if (m) rec_base[rec_n + __popc(nz & lt)] = make_uint2(cur.x, m);
        rec_n += __popc(nz);
      }
      if (__all_sync(0xffffffffu, T == 0.0f)) break;
      __syncwarp();                              // ring slot c is free before it is refilled
      cur = nxt;
      cnt_cur = cnt_nxt;
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

// ---- transposed forward blend ------------------------------------------------------------------------------------
// The kernel above evaluates one (half, Gaussian) hit per iteration with ~9 of 32 lanes on a pixel the splat reaches.
// This one collects the hits of successive chunks into blocks of 32 and handles a block in two phases with
// (nearly) every lane busy, like the transposed backward (render_bwd.cu):
//   phase 1, lane = HIT: each lane owns one Gaussian of the block, walks the pixels of the warp's 8x4 block inside
//     the Gaussian's conservative extent box (still-live pixels only) and writes their alphas — the same arithmetic,
//     bit for bit, as the per-pixel evaluation — into a packed slab, remembering which reached 1/255;
//   phase 2, lane = PIXEL: the 32x32 bit matrix {hit, pixel} is transposed with five shuffles and each lane blends,
//     front to back, only the hits that reach ITS pixel (alpha from the slab, colour/depth row of the hit from
//     shared memory): the transmittance chain, saturation test and contributor count are unchanged.
// The masks of the pixels that really blended are transposed back into the per-warp hit records.
constexpr int F2_CAP = 384;            // candidate (hit, pixel) pairs per pass; a hit has at most 32
constexpr size_t F2_WARP_BYTES = (size_t)STAGES * 3 * 32 * sizeof(float4) + 3 * 32 * sizeof(float4) + 32 * sizeof(uint2) +
                                 F2_CAP * sizeof(float) + 32 * sizeof(float2) + 2 * 32 * sizeof(uint32_t);

__device__ __forceinline__ uint32_t transpose32_fwd(uint32_t x, int lane) {
#pragma unroll
  for (int sft = 16; sft >= 1; sft >>= 1) {
    const uint32_t mlo = sft == 16 ? 0x0000ffffu : sft == 8 ? 0x00ff00ffu : sft == 4 ? 0x0f0f0f0fu
                         : sft == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, sft);
    x = (lane & sft) ? ((x & ~mlo) | ((y >> sft) & mlo)) : ((x & mlo) | ((y << sft) & ~mlo));
  }
  return x;
}

template <bool RECORD>
__global__ void __launch_bounds__(WARPS * 32)
render_fwd_transposed_kernel(View v, const Geom* __restrict__ geom, const uint32_t* __restrict__ point_list,
                             const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                             float* __restrict__ out_color, float* __restrict__ out_depth,
                             float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib,
                             float* __restrict__ final_T, uint2* __restrict__ hits, uint32_t* __restrict__ hit_count) {
  extern __shared__ float4 smem_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* wbase = reinterpret_cast<char*>(smem_dyn) + (size_t)warp * F2_WARP_BYTES;
  float4 (*ring)[3][32] = reinterpret_cast<float4 (*)[3][32]>(wbase);                       // [STAGES][3][32] candidates
  float4 (*hb)[32] = reinterpret_cast<float4 (*)[32]>(wbase + STAGES * 3 * 32 * sizeof(float4));   // [3][32] hit block
  uint2* hb_meta = reinterpret_cast<uint2*>(hb + 3);               // per hit: Gaussian id, list position + 1
  float* slab = reinterpret_cast<float*>(hb_meta + 32);            // alphas of the candidate pairs of a pass
  float2* s_pxy = reinterpret_cast<float2*>(slab + F2_CAP);        // per pixel (lane layout): x, y
  uint32_t* s_bb = reinterpret_cast<uint32_t*>(s_pxy + 32);        // per hit: candidate pixels (extent box, live)
  uint32_t* s_off = s_bb + 32;                                     // per hit: first slab slot

  const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles are launched first
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int lx = lane_px(lane), ly = lane_py(lane);
  const int pix_x = tx * TILE_X + wx + lx;
  const int pix_y = ty * TILE_Y + wy + ly;
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const float bx0 = (float)(tx * TILE_X + wx), by0 = (float)(ty * TILE_Y + wy);
  s_pxy[lane] = make_float2(pxf, pyf);
  // cull rectangle of the warp = bounding box of its pixels that are still accumulating
  float cxw = bx0 + 3.5f, cyw = by0 + 1.5f, hwx = 3.5f, hwy = 1.5f;
  uint32_t alive_prev = 0xffffffffu;

  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const int chunks = (n + 31) >> 5;
  const uint32_t* pl = point_list + range.x;

  float T = inside ? 1.0f : 0.0f, T_live = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dz = 0.f, A = 0.f;
  uint32_t last = 0;
  uint2* rec_base = RECORD ? hits + ((size_t)range.x * WARPS + (size_t)warp * (size_t)n) : nullptr;
  int rec_n = 0;
  const uint32_t lt = (1u << lane) - 1u;

  // one block of nb buffered hits (hb, hb_meta), in list order
  auto process_block = [&](int nb) {
    const uint32_t alive_now = __ballot_sync(0xffffffffu, T != 0.0f);
    // ---- the candidate pixels of every hit: its extent box inside the 8x4 block, live pixels only ----
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), q = a;
    uint32_t bb = 0u;
    if (lane < nb) {
      a = hb[0][lane];
      q = hb[1][lane];
      const int j0 = max(0, (int)ceilf(a.x - a.z - bx0)), j1 = min(7, (int)floorf(a.x + a.z - bx0));
      const int i0 = max(0, (int)ceilf(a.y - a.w - by0)), i1 = min(3, (int)floorf(a.y + a.w - by0));
      if (j1 >= j0 && i1 >= i0) {
        const uint32_t xm = (2u << j1) - (1u << j0), ym = (2u << i1) - (1u << i0);
        const uint32_t ysel = (ym & 1u) | ((ym & 2u) << 3) | ((ym & 4u) << 6) | ((ym & 8u) << 9);   // bits 0, 4, 8, 12
        bb = (((xm & 15u) * ysel) | (((xm >> 4) * ysel) << 16)) & alive_now;
      }
    }
    const int cntc = __popc(bb);
    int incl = cntc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const int off = incl - cntc;
    s_bb[lane] = bb;
    s_off[lane] = (uint32_t)off;
    int lo = 0;
    while (lo < nb) {
      // as many hits as fit the slab (their candidate counts are a prefix sum, so the fitting lanes are a run)
      const int base = __shfl_sync(0xffffffffu, off, lo);
      const uint32_t fm = __ballot_sync(0xffffffffu, lane >= lo && lane < nb && incl - base <= F2_CAP);
      const int hi = lo + __popc(fm);
      // ---- phase 1: lane = hit ----
      uint32_t vm = 0u;
      if (lane >= lo && lane < hi) {
        uint32_t m = bb;
        float* dst = slab + (off - base);
        while (m) {
          const int p = __ffs(m) - 1;
          m &= m - 1;
          const float2 xy = s_pxy[p];
          const float dx = a.x - xy.x, dy = a.y - xy.y;
          const float e2 = gauss_exponent2(q.x, q.y, q.z, dx, dy);      // log2 of the Gaussian weight
          const float al = e2 <= 0.0f ? fminf(ALPHA_CAP, q.w * exp2_blend(e2)) : 0.0f;
          *dst++ = al;
          if (al >= ALPHA_MIN) vm |= 1u << p;
        }
      }
      __syncwarp();
      // ---- phase 2: lane = pixel ----
      uint32_t pend = transpose32_fwd(vm, lane);           // hits of this pass that reach this lane's pixel
      uint32_t blended = 0u;
      while (__any_sync(0xffffffffu, pend != 0u)) {
        const int k = __ffs(pend) - 1;                      // front to back; -1: nothing left for this pixel
        if (k >= 0) {
          pend &= pend - 1;
          const float al = slab[(int)s_off[k] - base + __popc(s_bb[k] & lt)];
          const float test_T = T * (1.0f - al);
          const bool ok = test_T >= T_MIN;
          if (ok) {
            const float4 ff = hb[2][k];                     // depth, r, g, b
            const float w = al * T;
            C0 += ff.y * w; C1 += ff.z * w; C2 += ff.w * w;
            Dz += ff.x * w; A += w;
            T_live = test_T;
            last = hb_meta[k].y;
            blended |= 1u << k;
          }
          T = ok ? test_T : 0.0f;       // a saturating splat (or a finished pixel) leaves T at 0
        }
      }
      if (RECORD) {
        const uint32_t rm = transpose32_fwd(blended, lane);   // lane = hit: the pixels that blended it
        const uint32_t nz = __ballot_sync(0xffffffffu, rm != 0u);
        if (rm) rec_base[rec_n + __popc(nz & lt)] = make_uint2(hb_meta[lane].x, rm);
        rec_n += __popc(nz);
      }
      __syncwarp();                                         // the slab is rewritten by the next pass
      lo = hi;
    }
  };

  if (chunks > 0 && __any_sync(0xffffffffu, T != 0.0f)) {
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
      cp_async_commit();
    };
    auto fetch_gid = [&](int c) -> uint32_t {
      const int e = c * 32 + lane;
      return (c < chunks && e < n) ? pl[e] : 0u;
    };
    static_assert(STAGES == 2, "the ids of exactly one chunk are kept in flight");
    uint32_t gid_cur = fetch_gid(0);
    issue(0, gid_cur);
    uint32_t gid_next = fetch_gid(1);
    int nh = 0;                                   // hits buffered in hb
    bool all_done = false;
    for (int c = 0; c < chunks; ++c) {
      const uint32_t gid_issued = gid_next;
      issue(c + 1, gid_next);
      gid_next = fetch_gid(c + 2);
      cp_async_wait<STAGES - 1>();
      __syncwarp();
      float4 (*st)[32] = ring[c & (STAGES - 1)];
      const int e = c * 32 + lane;
      const bool done = T == 0.0f;
      const uint32_t alive = __ballot_sync(0xffffffffu, !done);
      if (alive != alive_prev) {
        alive_prev = alive;
        const int x0 = __reduce_min_sync(0xffffffffu, done ? 64 : lx), x1 = __reduce_max_sync(0xffffffffu, done ? -1 : lx);
        const int y0 = __reduce_min_sync(0xffffffffu, done ? 64 : ly), y1 = __reduce_max_sync(0xffffffffu, done ? -1 : ly);
        hwx = 0.5f * (float)(x1 - x0); hwy = 0.5f * (float)(y1 - y0);
        cxw = bx0 + (float)x0 + hwx; cyw = by0 + (float)y0 + hwy;
      }
      bool hit = false;
      float4 r0, r1, r2;
      if (e < n) {
        r0 = st[0][lane];
        hit = (fabsf(r0.x - cxw) <= r0.z + hwx) && (fabsf(r0.y - cyw) <= r0.w + hwy);
      }
      const uint32_t hm = __ballot_sync(0xffffffffu, hit);
      if (hm) {
        const int cnt = __popc(hm);
        const int pos = nh + __popc(hm & lt);
        if (hit) { r1 = st[1][lane]; r2 = st[2][lane]; }
        if (hit && pos < 32) {
          hb[0][pos] = r0; hb[1][pos] = r1; hb[2][pos] = r2;
          hb_meta[pos] = make_uint2(gid_cur, (uint32_t)(e + 1));
        }
        if (nh + cnt >= 32) {
          __syncwarp();
          process_block(32);
          if (hit && pos >= 32) {
            hb[0][pos - 32] = r0; hb[1][pos - 32] = r1; hb[2][pos - 32] = r2;
            hb_meta[pos - 32] = make_uint2(gid_cur, (uint32_t)(e + 1));
          }
          nh = nh + cnt - 32;
        } else {
          nh += cnt;
        }
      }
      gid_cur = gid_issued;
      if (__all_sync(0xffffffffu, T == 0.0f)) { all_done = true; break; }
      __syncwarp();                              // ring slot c is free before it is refilled
    }
    if (!all_done && nh > 0) {
      __syncwarp();
      process_block(nh);
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

// Tensor map of the Geom array for the gather4 variant: fp32 [P rows][12], box {16, 1}, no swizzle.
static int make_geom_map(const Geom* geom, int P, CUtensorMap* tm) {
  static std::atomic<PFN_cuTensorMapEncodeTiled_v12000> encode{nullptr};
  PFN_cuTensorMapEncodeTiled_v12000 fn = encode.load(std::memory_order_acquire);
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* sym = nullptr;
    GSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
    if (!sym) return GSB_E_UNSUPPORTED;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    encode.store(fn, std::memory_order_release);
  }
  const cuuint64_t gdim[2] = {12, (cuuint64_t)(P > 0 ? P : 1)};
  const cuuint64_t gstr[1] = {sizeof(Geom)};
  const cuuint32_t box[2] = {16, 1};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<Geom*>(geom), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GSB_OK : GSB_E_CUDA;
}

// variant: 0 per-hit blend with cp.async gathers (default), 1 transposed two-phase blend, 2 per-hit blend with TMA
// gather4 row gathers, 3 per-hit blend behind a CTA-wide pre-cull
int launch_render_fwd(const View& v, int P, const Geom* geom, const uint32_t* point_list,
                      const uint2* ranges, const uint32_t* tile_order, float* color, float* depth, float* alpha,
                      uint32_t* n_contrib, float* final_T, uint2* hits, uint32_t* hit_count, int variant,
                      bool debug, cudaStream_t st) {
  const int T = v.gx * v.gy;
  if (T == 0) return GSB_OK;
  const bool transposed = variant == 1;
#ifndef GSB_FWD_SMEM_PAD
#define GSB_FWD_SMEM_PAD 0
#endif
  constexpr size_t smem = (size_t)WARPS * STAGES * 3 * 32 * sizeof(float4) + (size_t)WARPS * 2 * 32 * sizeof(uint32_t) + GSB_FWD_SMEM_PAD;
  constexpr size_t smem_g4 = (size_t)WARPS * STAGES * G4_STAGE_BYTES + (size_t)WARPS * 2 * 32 * sizeof(uint32_t) +
                             (size_t)WARPS * STAGES * sizeof(uint64_t) + GSB_FWD_SMEM_PAD;
  constexpr size_t smem_pc = PC_SMEM_BYTES + GSB_FWD_SMEM_PAD;
  static std::atomic<unsigned long long> configured{0};   // bit per device: the attribute is per device
  int dev = 0;
  GSB_CUDA(cudaGetDevice(&dev));
  if (!(configured.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g4));
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g4));
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_precull_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pc));
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_precull_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pc));
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_transposed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(WARPS * F2_WARP_BYTES + GSB_FWD_SMEM_PAD)));
    GSB_CUDA(cudaFuncSetAttribute(render_fwd_transposed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(WARPS * F2_WARP_BYTES + GSB_FWD_SMEM_PAD)));
    configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  if (transposed) {
    const size_t smem2 = WARPS * F2_WARP_BYTES + GSB_FWD_SMEM_PAD;
    if (hits && hit_count)
      render_fwd_transposed_kernel<true><<<T, WARPS * 32, smem2, st>>>(v, geom, point_list, ranges, tile_order, color,
                                                                       depth, alpha, n_contrib, final_T, hits, hit_count);
    else
      render_fwd_transposed_kernel<false><<<T, WARPS * 32, smem2, st>>>(v, geom, point_list, ranges, tile_order, color,
                                                                        depth, alpha, n_contrib, final_T, nullptr, nullptr);
    GSB_POST_LAUNCH(debug, st, "render_fwd_transposed_kernel");
    return GSB_OK;
  }
  if (variant == 3) {
    if (hits && hit_count)
      render_fwd_precull_kernel<true><<<T, WARPS * 32, smem_pc, st>>>(v, geom, point_list, ranges, tile_order, color,
                                                                      depth, alpha, n_contrib, final_T, hits, hit_count);
    else
      render_fwd_precull_kernel<false><<<T, WARPS * 32, smem_pc, st>>>(v, geom, point_list, ranges, tile_order, color,
                                                                       depth, alpha, n_contrib, final_T, nullptr, nullptr);
    GSB_POST_LAUNCH(debug, st, "render_fwd_precull_kernel");
    return GSB_OK;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (variant == 2) {
    const int rc = make_geom_map(geom, P, &tm);
    if (rc) return rc;
    if (hits && hit_count)
      render_fwd_kernel<true, true><<<T, WARPS * 32, smem_g4, st>>>(v, tm, geom, point_list, ranges, tile_order, color,
                                                                    depth, alpha, n_contrib, final_T, hits, hit_count);
    else
      render_fwd_kernel<false, true><<<T, WARPS * 32, smem_g4, st>>>(v, tm, geom, point_list, ranges, tile_order, color,
                                                                     depth, alpha, n_contrib, final_T, nullptr, nullptr);
    GSB_POST_LAUNCH(debug, st, "render_fwd_kernel<gather4>");
    return GSB_OK;
  }
  if (hits && hit_count)
    render_fwd_kernel<true, false><<<T, WARPS * 32, smem, st>>>(v, tm, geom, point_list, ranges, tile_order, color,
                                                                depth, alpha, n_contrib, final_T, hits, hit_count);
  else
    render_fwd_kernel<false, false><<<T, WARPS * 32, smem, st>>>(v, tm, geom, point_list, ranges, tile_order, color,
                                                                 depth, alpha, n_contrib, final_T, nullptr, nullptr);
  GSB_POST_LAUNCH(debug, st, "render_fwd_kernel");
  return GSB_OK;
}

}  // namespace gsb
