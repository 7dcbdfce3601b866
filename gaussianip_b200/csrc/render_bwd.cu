// Backward stage 1: per-pixel reverse (back-to-front) blending.  Replaces renderCUDA
// (backward) of the external operator ("Backward blend").
//
// Two kernels:
//  * render_bwd_replay_kernel (default): walks, back to front, the HIT RECORDS the forward blend wrote per warp
//    ({Gaussian id, mask of the lanes that blended it}, render_fwd.cu).  Every record is a Gaussian at least one
//    pixel of the warp really blended, the mask says which, so nothing is culled, rect-tested or alpha-tested
//    again: of the round-1 kernel's ~276 M warp instructions per view, the list walk (7.7 M candidates), the
//    rectangle tests and the 1.5 M whole-warp alpha evaluations that only decide validity are gone.
//  * render_bwd_kernel (gsb_set_blend_variant(3/4), "rescan"): the round-1 kernel, which needs no records and
//    re-walks the tile list with per-warp culling.  Kept as the cross-check of the replay kernel.
//
// Rescan kernel:
// Same organisation as render_fwd.cu: one CTA per 16x16 tile, warp = 8x4 pixels, warps fully
// independent (no __syncthreads), per-warp cp.async ring of 32-instance chunks, per-warp
// conservative culling.  Differences that matter:
//  * each warp starts at the LAST instance any of ITS pixels blended (warp max of n_contrib),
//    not at the end of the tile's list, and walks towards the front;
//  * the ten per-Gaussian partial gradients of the 32 pixels of a warp are summed with a
//    13-shuffle multi-value butterfly and leave the SM as ONE red.global per component per
//    (warp, Gaussian) — 10 lanes hitting one 48 B GGrad record — instead of ~10 atomicAdd
//    per (pixel, Gaussian).
#include <atomic>
#include <cstdio>

#include "gsb_common.cuh"

namespace gsb {

namespace {

#ifdef GSB_BWD_STATS
// measurement build only (GSB_NVCC_EXTRA=-DGSB_BWD_STATS): where do the warp-hits go?
__device__ unsigned long long g_bwd_stats[12];
#endif

constexpr int WARPS = 8;
#ifndef GSB_BWD_STAGES
#define GSB_BWD_STAGES 2
#endif
constexpr int STAGES = GSB_BWD_STAGES;
#ifndef GSB_BWD_HB
#define GSB_BWD_HB 1
#endif
constexpr int BH = GSB_BWD_HB;       // hits evaluated together (rescan kernel)
#ifndef GSB_BWD_RB
#define GSB_BWD_RB 1
#endif
constexpr int RB = GSB_BWD_RB;       // records replayed together (replay kernel)

// EXPERIMENTAL variant (gsb_set_blend_variant(2), not the default; passed gradient parity on one scene, not yet
// benchmarked — DESIGN.md §8.1):
// instead of reducing every hit's ten partial sums over the 32 lanes with the butterfly (51 instructions for, on
// average, 7.4 useful lanes), the lanes that actually contribute append their ten values to a packed per-warp slab
// in shared memory, one segment per hit; when the slab fills up (and at the end) lanes (g, j) = (lane / 10, lane % 10)
// sum component j of segment 3 r + g sequentially and issue the red.  Segments of different hits are independent,
// so three hits are reduced per pass.
#ifndef GSB_BWD_SLAB
#define GSB_BWD_SLAB 40
#endif
constexpr int SLAB = GSB_BWD_SLAB;     // packed (pixel, hit) pairs per warp between two reductions (>= 32)
constexpr int SLAB_STRIDE = 11;        // floats per pair: 10 used, odd stride keeps the stores conflict-free
constexpr int SEG_MAX = 30;            // hits per reduction (three at a time)
static_assert(SLAB >= 32, "one hit can contribute 32 pairs");
constexpr size_t PACKED_WARP_BYTES = (size_t)SLAB * SLAB_STRIDE * 4 + (SEG_MAX + 2) * 4 + SEG_MAX * 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Sum 12 per-lane values over the 32 lanes of a warp.  On return lane L (any L) holds in the
// return value the total of slot  (L&16 ? 6:0) + (L&8 ? 3:0) + (L&4 ? 2:0) + (L&2 ? 1:0);
// the combination (L&4 && L&2) is the padding slot and must be ignored.
__device__ __forceinline__ float butterfly12(const float (&v)[12], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  float u[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float keep = b4 ? v[i + 6] : v[i];
    const float send = b4 ? v[i] : v[i + 6];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float t[4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float keep = b3 ? u[i + 3] : u[i];
    const float send = b3 ? u[i] : u[i + 3];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  t[3] = 0.0f;
  float s[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = b2 ? t[i + 2] : t[i];
    const float send = b2 ? t[i] : t[i + 2];
    s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = b1 ? s[1] : s[0];
  const float send = b1 ? s[0] : s[1];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

template <bool PACKED>
__global__ void __launch_bounds__(WARPS * 32)
render_bwd_kernel(View v, const Geom* __restrict__ geom, const uint32_t* __restrict__ point_list,
                  const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ final_T,
                  const float* __restrict__ dL_dcolor,
                  const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha,
                  GGrad* __restrict__ ggrad) {
  extern __shared__ float4 smem_dyn[];
  float4 (*s_rec)[STAGES][3][32] = reinterpret_cast<float4 (*)[STAGES][3][32]>(smem_dyn);
  uint32_t (*s_gid)[STAGES][32] =
      reinterpret_cast<uint32_t (*)[STAGES][32]>(smem_dyn + WARPS * STAGES * 3 * 32);
  // PACKED only: per-warp slab + segment table behind the rings
  char* packed_base = reinterpret_cast<char*>(smem_dyn) + (size_t)WARPS * STAGES * (3 * 32 * sizeof(float4) + 32 * sizeof(uint32_t));

  const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles are launched first
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  float* slab = reinterpret_cast<float*>(packed_base + (size_t)warp * PACKED_WARP_BYTES);
  int* seg_start = reinterpret_cast<int*>(slab + SLAB * SLAB_STRIDE);        // [SEG_MAX + 2]
  uint32_t* seg_gid = reinterpret_cast<uint32_t*>(seg_start + SEG_MAX + 2);  // [SEG_MAX]
  int n_pairs = 0, n_seg = 0;                                                // warp-uniform
  const int pix_x = tx * TILE_X + wx + lane_px(lane);
  const int pix_y = ty * TILE_Y + wy + lane_py(lane);
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  // Cull rectangle of the warp = bounding box of the pixels that have started their walk (a
  // pixel joins when the position drops to its n_contrib); it grows towards the full 8x4 block.
  float cxw = 0.f, cyw = 0.f, hwx = -1.f, hwy = -1.f;
  uint32_t active_prev = 0u;
  const size_t hw = (size_t)v.H * v.W;
  const size_t pix = (size_t)pix_y * v.W + pix_x;

  const uint2 range = ranges[tile];
  const uint32_t my_last = inside ? n_contrib[pix] : 0u;
  const int n = (int)__reduce_max_sync(0xffffffffu, my_last);   // instances this warp must revisit
  if (n == 0) return;
  const int chunks = (n + 31) >> 5;
  const uint32_t* pl = point_list + range.x;

  const float T_final = inside ? final_T[pix] : 0.0f;
  float T = T_final;
  const float gC0 = inside ? dL_dcolor[pix] : 0.f, gC1 = inside ? dL_dcolor[hw + pix] : 0.f,
              gC2 = inside ? dL_dcolor[2 * hw + pix] : 0.f;
  const float gD = inside ? dL_ddepth[pix] : 0.f, gA = inside ? dL_dalpha[pix] : 0.f;
  const float bg_dot = v.bg[0] * gC0 + v.bg[1] * gC1 + v.bg[2] * gC2;
  float Bdot = T_final * bg_dot;     // see the derivation at its use

  float4 (*ring)[3][32] = s_rec[warp];
  uint32_t (*gring)[32] = s_gid[warp];
  // chunk index c counts from the BACK: it covers list positions [(chunks-1-c)*32, +32)
  auto issue = [&](int c, uint32_t gid) {
    if (c < chunks) {
      const int e = (chunks - 1 - c) * 32 + lane;
      if (e < n) {
        const float4* src = reinterpret_cast<const float4*>(geom + gid);
        float4 (*st)[32] = ring[c & (STAGES - 1)];
        cp_async16(&st[0][lane], src);
        cp_async16(&st[1][lane], src + 1);
        cp_async16(&st[2][lane], src + 2);
        gring[c & (STAGES - 1)][lane] = gid;
      }
    }
    cp_async_commit();
  };
  auto fetch_gid = [&](int c) -> uint32_t {
    if (c >= chunks) return 0u;
    const int e = (chunks - 1 - c) * 32 + lane;
    return e < n ? pl[e] : 0u;
  };
  // PACKED: reduce the slab's segments (one per hit) and send them out; three hits per pass
  auto flush = [&]() {
    if (lane == 0) seg_start[n_seg] = n_pairs;          // end sentinel
    __syncwarp();
    const int grp = lane / 10, comp = lane - grp * 10;    // lanes 30, 31 (grp 3) idle
    for (int h0 = 0; h0 < n_seg; h0 += 3) {
      const int h = h0 + grp;
      if (grp < 3 && h < n_seg) {
        const int a = seg_start[h], b = seg_start[h + 1];
        float sum = 0.0f;
        for (int sl = a; sl < b; ++sl) sum += slab[sl * SLAB_STRIDE + comp];
        if (sum != 0.0f) atomicAdd(reinterpret_cast<float*>(ggrad + seg_gid[h]) + comp, sum);
      }
    }
    __syncwarp();
    n_pairs = 0; n_seg = 0;
  };
#pragma unroll
  for (int c = 0; c < STAGES - 1; ++c) issue(c, fetch_gid(c));
  uint32_t gid_next = fetch_gid(STAGES - 1);

#ifdef GSB_BWD_STATS
  unsigned long long st_cand = 0, st_rect = 0, st_any = 0, st_lanes = 0, st_hist[5] = {0, 0, 0, 0, 0};
#endif
  for (int c = 0; c < chunks; ++c) {
    issue(c + STAGES - 1, gid_next);
    gid_next = fetch_gid(c + STAGES);
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    float4 (*st)[32] = ring[c & (STAGES - 1)];
    const int base = (chunks - 1 - c) * 32;
    const int e = base + lane;
    const bool started = my_last > (uint32_t)base;          // has a position inside or behind this chunk
    const uint32_t active = __ballot_sync(0xffffffffu, started);
    if (active != active_prev) {
      active_prev = active;
      const int lx = lane_px(lane), ly = lane_py(lane);
      const int x0 = __reduce_min_sync(0xffffffffu, started ? lx : 64), x1 = __reduce_max_sync(0xffffffffu, started ? lx : -1);
      const int y0 = __reduce_min_sync(0xffffffffu, started ? ly : 64), y1 = __reduce_max_sync(0xffffffffu, started ? ly : -1);
      hwx = 0.5f * (float)(x1 - x0); hwy = 0.5f * (float)(y1 - y0);
      cxw = (float)(tx * TILE_X + wx + x0) + hwx; cyw = (float)(ty * TILE_Y + wy + y0) + hwy;
    }
    bool hit = false;
    if (e < n) {
      const float4 a = st[0][lane];
      hit = (fabsf(a.x - cxw) <= a.z + hwx) && (fabsf(a.y - cyw) <= a.w + hwy);
    }
    uint32_t mask = __ballot_sync(0xffffffffu, hit);
#ifdef GSB_BWD_STATS
    st_cand += __popc(__ballot_sync(0xffffffffu, e < n));
    st_rect += __popc(mask);
#endif
    // Hits are taken BH at a time (back to front): their geometry/alpha are independent and the
    // 13-shuffle reductions can interleave.
    while (mask) {
      int k[BH];
#pragma unroll
      for (int i = 0; i < BH; ++i) {
        k[i] = mask ? 31 - __clz(mask) : -1;
        if (k[i] >= 0) mask &= ~(1u << k[i]);
      }
      float dx[BH], dy[BH], G[BH], alpha[BH];
      float4 q[BH], f[BH];
      bool valid[BH];
#pragma unroll
      for (int i = 0; i < BH; ++i) {
        valid[i] = false;
        if (k[i] >= 0) {
          const float4 a = st[0][k[i]];
          q[i] = st[1][k[i]];
          f[i] = st[2][k[i]];
          dx[i] = a.x - pxf; dy[i] = a.y - pyf;
          const float power = gauss_exponent2(q[i].x, q[i].y, q[i].z, dx[i], dy[i]);   // log2 of the weight
          G[i] = exp2_blend(power);
          alpha[i] = fminf(ALPHA_CAP, q[i].w * G[i]);
          valid[i] = ((uint32_t)(base + k[i] + 1) <= my_last) && power <= 0.0f && alpha[i] >= ALPHA_MIN;
        }
      }
      float g[BH][12];
#pragma unroll
      for (int i = 0; i < BH; ++i) {
#pragma unroll
        for (int j = 0; j < 12; ++j) g[i][j] = 0.0f;
        if (valid[i]) {
          // With phi_j = <c_j, dL/dC> + depth_j dL/dD + dL/dA and the suffix sum
          //   Bdot_i = T_final <bg, dL/dC> + sum_{j behind i} phi_j alpha_j T_j,
          // dL/dalpha_i = T_i phi_i - Bdot_i / (1 - alpha_i): the reference's five "colour behind"
          // recurrences collapse into ONE scalar (identical in exact arithmetic).
          const float inv1ma = rcp_blend(1.0f - alpha[i]);
          T *= inv1ma;
          const float w = alpha[i] * T;
          const float phi = f[i].y * gC0 + f[i].z * gC1 + f[i].w * gC2 + f[i].x * gD + gA;
          const float dL_da = T * phi - Bdot * inv1ma;
          Bdot += phi * w;
          // raw moments of s = dL/dG * G about the splat centre; the linear maps to
          // d(ndc xy) and d(conic) are applied once per Gaussian in preprocess_bwd
          const float s_ = q[i].w * dL_da * G[i];
          g[i][0] = s_ * dx[i];
          g[i][1] = s_ * dy[i];
          g[i][2] = g[i][0] * dx[i];
          g[i][3] = g[i][0] * dy[i];
          g[i][4] = g[i][1] * dy[i];
          g[i][5] = G[i] * dL_da;
          g[i][6] = w * gD;
          g[i][7] = w * gC0;
          g[i][8] = w * gC1;
          g[i][9] = w * gC2;
        }
      }
      if (PACKED) {
#pragma unroll
        for (int i = 0; i < BH; ++i) {
          const uint32_t vb = __ballot_sync(0xffffffffu, valid[i]);
          if (!vb) continue;
          const int nv = __popc(vb);
          if (n_pairs + nv > SLAB || n_seg == SEG_MAX) flush();
          if (lane == 0) { seg_start[n_seg] = n_pairs; seg_gid[n_seg] = gring[c & (STAGES - 1)][k[i]]; }
          if (valid[i]) {
            float* dst = slab + (n_pairs + __popc(vb & ((1u << lane) - 1u))) * SLAB_STRIDE;
#pragma unroll
            for (int j = 0; j < 10; ++j) dst[j] = g[i][j];
          }
          n_pairs += nv; ++n_seg;
        }
        continue;
      }
      const int slot = ((lane & 16) ? 6 : 0) + ((lane & 8) ? 3 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
      const bool writer = !(lane & 1) && !((lane & 4) && (lane & 2)) && slot < 10;
      // (direct per-lane atomics for splats that only graze the warp were measured slower:
      //  521 -> 560 us at FRINGE=3, 569 us at FRINGE=1; the butterfly is kept for every hit)
#pragma unroll
      for (int i = 0; i < BH; ++i) {
#ifdef GSB_BWD_STATS
        {
          const int nv = __popc(__ballot_sync(0xffffffffu, valid[i]));
          if (nv) { st_any++; st_lanes += nv; st_hist[nv <= 2 ? 0 : nv <= 4 ? 1 : nv <= 8 ? 2 : nv <= 16 ? 3 : 4]++; }
        }
#endif
        if (!__any_sync(0xffffffffu, valid[i])) continue;
        const float tot = butterfly12(g[i], lane);
        if (writer && tot != 0.0f)
          atomicAdd(reinterpret_cast<float*>(ggrad + gring[c & (STAGES - 1)][k[i]]) + slot, tot);
      }
    }
    __syncwarp();
  }
  if (PACKED && n_seg) flush();
  cp_async_wait<0>();
#ifdef GSB_BWD_STATS
  if (lane == 0) {
    atomicAdd(&g_bwd_stats[0], (unsigned long long)chunks);
    atomicAdd(&g_bwd_stats[1], st_cand); atomicAdd(&g_bwd_stats[2], st_rect);
    atomicAdd(&g_bwd_stats[3], st_any); atomicAdd(&g_bwd_stats[4], st_lanes);
    for (int i = 0; i < 5; ++i) atomicAdd(&g_bwd_stats[5 + i], st_hist[i]);
  }
#endif
}


// ---- replay kernel --------------------------------------------------------------------------------------------
// Same per-pixel state and arithmetic as the rescan kernel (T rebuilt backwards from final_T, one suffix scalar
// Bdot), same raw-moment outputs; the iteration space is the warp's record list instead of the tile's instance list.
// Records are fetched 32 at a time (one coalesced 256 B load, lane l <- record l), their Geom rows gathered with
// cp.async into the same 2-stage ring, one block of records ahead of the arithmetic.
// The two HALVES of the warp (lanes 0-15 / 16-31 = left / right 4x4 pixels, bits 0-15 / 16-31 of a record's mask)
// replay their own subsequences of a block: an iteration handles the next record of the left half and the next
// record of the right half — usually two different Gaussians — and reduces the ten partial sums of each over its
// 16 lanes with a 12-shuffle butterfly, so a block of 32 records costs max(|A|, |B|) iterations instead of 32.

// Sum 12 per-lane values over the 16 lanes of a half-warp.  On return lane L holds the total of slot
// (L&8 ? 6:0) + (L&4 ? 3:0) + (L&2 ? 2:0) + (L&1 ? 1:0); the combination (L&2 && L&1) is the padding slot.
__device__ __forceinline__ float butterfly12_half(const float (&v)[12], int lane) {
  const bool b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
  float u[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float keep = b3 ? v[i + 6] : v[i];
    const float send = b3 ? v[i] : v[i + 6];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float t[4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float keep = b2 ? u[i + 3] : u[i];
    const float send = b2 ? u[i] : u[i + 3];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  t[3] = 0.0f;
  float s[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = b1 ? t[i + 2] : t[i];
    const float send = b1 ? t[i] : t[i + 2];
    s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float keep = b0 ? s[1] : s[0];
  const float send = b0 ? s[0] : s[1];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

template <bool PACKED>
__global__ void __launch_bounds__(WARPS * 32)
render_bwd_replay_kernel(View v, const Geom* __restrict__ geom, const uint2* __restrict__ ranges,
                         const uint32_t* __restrict__ tile_order, const uint2* __restrict__ hits,
                         const uint32_t* __restrict__ hit_count, const float* __restrict__ final_T,
                         const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                         const float* __restrict__ dL_dalpha, GGrad* __restrict__ ggrad) {
  extern __shared__ float4 smem_dyn[];
  float4 (*s_rec)[STAGES][3][32] = reinterpret_cast<float4 (*)[STAGES][3][32]>(smem_dyn);
  uint2 (*s_meta)[STAGES][32] = reinterpret_cast<uint2 (*)[STAGES][32]>(smem_dyn + WARPS * STAGES * 3 * 32);
  char* packed_base = reinterpret_cast<char*>(smem_dyn) + (size_t)WARPS * STAGES * (3 * 32 * sizeof(float4) + 32 * sizeof(uint2));

  const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles are launched first
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4;
  const int n_rec = (int)hit_count[tile * WARPS + warp];
  if (n_rec == 0) return;
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  float* slab = reinterpret_cast<float*>(packed_base + (size_t)warp * PACKED_WARP_BYTES);
  int* seg_start = reinterpret_cast<int*>(slab + SLAB * SLAB_STRIDE);        // [SEG_MAX + 2]
  uint32_t* seg_gid = reinterpret_cast<uint32_t*>(seg_start + SEG_MAX + 2);  // [SEG_MAX]
  int n_pairs = 0, n_seg = 0;                                                // warp-uniform
  const int pix_x = tx * TILE_X + wx + lane_px(lane);
  const int pix_y = ty * TILE_Y + wy + lane_py(lane);
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const size_t hw = (size_t)v.H * v.W;
  const size_t pix = (size_t)pix_y * v.W + pix_x;
  const uint2 range = ranges[tile];
  const uint2* recs = hits + ((size_t)range.x * WARPS + (size_t)warp * (size_t)(range.y - range.x));
  const int blocks = (n_rec + 31) >> 5;

  const float T_final = inside ? final_T[pix] : 0.0f;
  float T = T_final;
  const float gC0 = inside ? dL_dcolor[pix] : 0.f, gC1 = inside ? dL_dcolor[hw + pix] : 0.f,
              gC2 = inside ? dL_dcolor[2 * hw + pix] : 0.f;
  const float gD = inside ? dL_ddepth[pix] : 0.f, gA = inside ? dL_dalpha[pix] : 0.f;
  const float bg_dot = v.bg[0] * gC0 + v.bg[1] * gC1 + v.bg[2] * gC2;
  float Bdot = T_final * bg_dot;

  float4 (*ring)[3][32] = s_rec[warp];
  uint2 (*mring)[32] = s_meta[warp];
  // block index c counts from the BACK: it covers records [(blocks-1-c)*32, +32)
  auto fetch_rec = [&](int c) -> uint2 {
    if (c >= blocks) return make_uint2(0u, 0u);
    const int e = (blocks - 1 - c) * 32 + lane;
    return e < n_rec ? recs[e] : make_uint2(0u, 0u);
  };
  auto issue = [&](int c, uint2 rec) {
    if (c < blocks) {
      const int e = (blocks - 1 - c) * 32 + lane;
      if (e < n_rec) {
        const float4* src = reinterpret_cast<const float4*>(geom + rec.x);
        float4 (*st)[32] = ring[c & (STAGES - 1)];
        cp_async16(&st[0][lane], src);
        cp_async16(&st[1][lane], src + 1);
        cp_async16(&st[2][lane], src + 2);
      }
      mring[c & (STAGES - 1)][lane] = rec;        // (0, 0) beyond the end: belongs to neither half
    }
    cp_async_commit();
  };
  auto flush = [&]() {
    if (lane == 0) seg_start[n_seg] = n_pairs;          // end sentinel
    __syncwarp();
    const int grp = lane / 10, comp = lane - grp * 10;    // lanes 30, 31 (grp 3) idle
    for (int h0 = 0; h0 < n_seg; h0 += 3) {
      const int h = h0 + grp;
      if (grp < 3 && h < n_seg) {
        const int a = seg_start[h], b = seg_start[h + 1];
        float sum = 0.0f;
        for (int sl = a; sl < b; ++sl) sum += slab[sl * SLAB_STRIDE + comp];
        if (sum != 0.0f) atomicAdd(reinterpret_cast<float*>(ggrad + seg_gid[h]) + comp, sum);
      }
    }
    __syncwarp();
    n_pairs = 0; n_seg = 0;
  };
#pragma unroll
  for (int c = 0; c < STAGES - 1; ++c) issue(c, fetch_rec(c));
  uint2 rec_next = fetch_rec(STAGES - 1);

  const int slot = ((lane & 8) ? 6 : 0) + ((lane & 4) ? 3 : 0) + ((lane & 2) ? 2 : 0) + ((lane & 1) ? 1 : 0);
  const bool writer = !((lane & 2) && (lane & 1)) && slot < 10;
  for (int c = 0; c < blocks; ++c) {
    issue(c + STAGES - 1, rec_next);
    rec_next = fetch_rec(c + STAGES);
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    float4 (*st)[32] = ring[c & (STAGES - 1)];
    const uint2* meta = mring[c & (STAGES - 1)];
    // which records of this block concern the left / the right half
    const uint32_t mine = meta[lane].y;
    const uint32_t recsA = __ballot_sync(0xffffffffu, (mine & 0x0000ffffu) != 0u);
    const uint32_t recsB = __ballot_sync(0xffffffffu, (mine & 0xffff0000u) != 0u);
    uint32_t pend = half ? recsB : recsA;         // every lane keeps the pending records of ITS half
    while (__any_sync(0xffffffffu, pend != 0u)) {
      // back to front: the highest remaining record of this lane's half (-1: none left)
      const int r = 31 - __clz(pend);
      pend &= ~(r >= 0 ? (1u << r) : 0u);
      uint2 mr = make_uint2(0u, 0u);
      bool valid = false;
      float g[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) g[j] = 0.0f;
      if (r >= 0) {
        mr = meta[r];                                     // {Gaussian id, lanes that blended it}: uniform per half
        valid = (mr.y >> lane) & 1u;
        if (valid) {
          const float4 a = st[0][r];
          const float4 q = st[1][r];
          const float4 f = st[2][r];
          const float dx = a.x - pxf, dy = a.y - pyf;
          const float G = exp2_blend(gauss_exponent2(q.x, q.y, q.z, dx, dy));
          const float alpha = fminf(ALPHA_CAP, q.w * G);
          // With phi_j = <c_j, dL/dC> + depth_j dL/dD + dL/dA and the suffix sum
          //   Bdot_i = T_final <bg, dL/dC> + sum_{j behind i} phi_j alpha_j T_j,
          // dL/dalpha_i = T_i phi_i - Bdot_i / (1 - alpha_i)
          const float inv1ma = rcp_blend(1.0f - alpha);
          T *= inv1ma;
          const float w = alpha * T;
          const float phi = f.y * gC0 + f.z * gC1 + f.w * gC2 + f.x * gD + gA;
          const float dL_da = T * phi - Bdot * inv1ma;
          Bdot += phi * w;
          // raw moments of s = dL/dG * G about the splat centre; the linear maps to d(ndc xy) and d(conic)
          // are applied once per Gaussian in preprocess_bwd
          const float s_ = q.w * dL_da * G;
          g[0] = s_ * dx;
          g[1] = s_ * dy;
          g[2] = g[0] * dx;
          g[3] = g[0] * dy;
          g[4] = g[1] * dy;
          g[5] = G * dL_da;
          g[6] = w * gD;
          g[7] = w * gC0;
          g[8] = w * gC1;
          g[9] = w * gC2;
        }
      }
      if (PACKED) {
        // one segment per (half, record): the left half's first, then the right half's
        const int rA = __shfl_sync(0xffffffffu, r, 0), rB = __shfl_sync(0xffffffffu, r, 16);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int rr = hh ? rB : rA;
          if (rr < 0) continue;
          const uint2 m2 = meta[rr];
          const uint32_t bits = m2.y & (hh ? 0xffff0000u : 0x0000ffffu);
          const int nv = __popc(bits);
          if (n_pairs + nv > SLAB || n_seg == SEG_MAX) flush();
          if (lane == 0) { seg_start[n_seg] = n_pairs; seg_gid[n_seg] = m2.x; }
          if (half == hh && valid) {
            float* dst = slab + (n_pairs + __popc(bits & ((1u << lane) - 1u))) * SLAB_STRIDE;
#pragma unroll
            for (int j = 0; j < 10; ++j) dst[j] = g[j];
          }
          n_pairs += nv; ++n_seg;
        }
      } else {
        const float tot = butterfly12_half(g, lane);
#ifdef GSB_BWD_NO_RED      // diagnostic build only (timing without the global reductions; results are wrong)
        if (r >= 0 && writer && tot == 123.456f) atomicAdd(reinterpret_cast<float*>(ggrad + mr.x) + slot, tot);
#else
        if (r >= 0 && writer && tot != 0.0f) atomicAdd(reinterpret_cast<float*>(ggrad + mr.x) + slot, tot);
#endif
      }
    }
    __syncwarp();
  }
  if (PACKED && n_seg) flush();
  cp_async_wait<0>();
}


// ---- transposed replay kernel (default) ---------------------------------------------------------------------------
// The replay kernel above still evaluates one (half, Gaussian) hit per warp iteration, with ~9 of 32 lanes holding a
// pixel the Gaussian really touched: 110 warp instructions per record (ncu r2d/r2e), most of them issued for idle
// lanes, 56 of them the shuffle butterfly.  This kernel splits a block of 32 records into two phases that both run
// with (nearly) every lane busy:
//   phase 1, lane = PIXEL: each lane walks, back to front, only the records of the block whose mask names it (the
//     32x32 bit matrix {record, lane} is transposed with five shuffles), runs the sequential part of the backward —
//     weight, alpha, transmittance, suffix sum — and leaves the two numbers the rest needs, u = G dL/dalpha and
//     w = alpha T, in a packed shared-memory slab (slot = prefix count of the record + rank of the lane in its mask);
//   phase 2, lane = RECORD: each lane owns one record of the block, loops over that record's pixels reading (u, w)
//     from the slab and the pixel's constants from two small tables, accumulates the ten sums in registers and
//     sends them to the Gaussian's gradient record with three 16-byte reductions.
// No butterfly, no per-hit atomics by idle lanes; a block costs max_lane(#records of a pixel) + max_record(#pixels)
// short iterations instead of 32 (or max(|A|,|B|)) long ones.
#ifndef GSB_BWD_T2_CAP
#define GSB_BWD_T2_CAP 512
#endif
constexpr int T2_CAP = GSB_BWD_T2_CAP;  // (pixel, record) pairs per pass; a record has at most 32
static_assert(T2_CAP >= 32, "one record must fit");
constexpr size_t T2_WARP_BYTES = (size_t)STAGES * (3 * 32 * sizeof(float4) + 32 * sizeof(uint2)) +
                                 T2_CAP * sizeof(float2) + 32 * sizeof(float4) + 32 * sizeof(float2) + 32 * sizeof(uint32_t);

// 32x32 bit-matrix transpose across the warp: bit r of the result in lane l = bit l of x in lane r
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int sft = 16; sft >= 1; sft >>= 1) {
    const uint32_t mlo = sft == 16 ? 0x0000ffffu : sft == 8 ? 0x00ff00ffu : sft == 4 ? 0x0f0f0f0fu
                         : sft == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, sft);
    x = (lane & sft) ? ((x & ~mlo) | ((y >> sft) & mlo)) : ((x & mlo) | ((y << sft) & ~mlo));
  }
  return x;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

#ifndef GSB_BWD_T2_MINB
#define GSB_BWD_T2_MINB 3
#endif
__global__ void __launch_bounds__(WARPS * 32, GSB_BWD_T2_MINB)
render_bwd_transposed_kernel(View v, const Geom* __restrict__ geom, const uint2* __restrict__ ranges,
                             const uint32_t* __restrict__ tile_order, const uint2* __restrict__ hits,
                             const uint32_t* __restrict__ hit_count, const float* __restrict__ final_T,
                             const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                             const float* __restrict__ dL_dalpha, GGrad* __restrict__ ggrad) {
  extern __shared__ float4 smem_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* wbase = reinterpret_cast<char*>(smem_dyn) + (size_t)warp * T2_WARP_BYTES;
  float4 (*ring)[3][32] = reinterpret_cast<float4 (*)[3][32]>(wbase);                          // [STAGES][3][32]
  uint2 (*mring)[32] = reinterpret_cast<uint2 (*)[32]>(wbase + STAGES * 3 * 32 * sizeof(float4));   // [STAGES][32]
  float2* slab = reinterpret_cast<float2*>(wbase + STAGES * (3 * 32 * sizeof(float4) + 32 * sizeof(uint2)));
  float4* s_pg = reinterpret_cast<float4*>(slab + T2_CAP);        // per pixel: dL/dD, dL/dC0, dL/dC1, dL/dC2
  float2* s_pxy = reinterpret_cast<float2*>(s_pg + 32);           // per pixel: x, y
  uint32_t* s_off = reinterpret_cast<uint32_t*>(s_pxy + 32);      // per record of the block: first slab slot

  const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles are launched first
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int n_rec = (int)hit_count[tile * WARPS + warp];
  if (n_rec == 0) return;
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int pix_x = tx * TILE_X + wx + lane_px(lane);
  const int pix_y = ty * TILE_Y + wy + lane_py(lane);
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const size_t hw = (size_t)v.H * v.W;
  const size_t pix = (size_t)pix_y * v.W + pix_x;
  const uint2 range = ranges[tile];
  const uint2* recs = hits + ((size_t)range.x * WARPS + (size_t)warp * (size_t)(range.y - range.x));
  const int blocks = (n_rec + 31) >> 5;

  const float T_final = inside ? final_T[pix] : 0.0f;
  float T = T_final;
  const float gC0 = inside ? dL_dcolor[pix] : 0.f, gC1 = inside ? dL_dcolor[hw + pix] : 0.f,
              gC2 = inside ? dL_dcolor[2 * hw + pix] : 0.f;
  const float gD = inside ? dL_ddepth[pix] : 0.f, gA = inside ? dL_dalpha[pix] : 0.f;
  const float bg_dot = v.bg[0] * gC0 + v.bg[1] * gC1 + v.bg[2] * gC2;
  float Bdot = T_final * bg_dot;     // suffix sum, see render_bwd_kernel
  s_pg[lane] = make_float4(gD, gC0, gC1, gC2);
  s_pxy[lane] = make_float2(pxf, pyf);

  // block index c counts from the BACK: it covers records [(blocks-1-c)*32, +32)
  auto fetch_rec = [&](int c) -> uint2 {
    if (c >= blocks) return make_uint2(0u, 0u);
    const int e = (blocks - 1 - c) * 32 + lane;
    return e < n_rec ? recs[e] : make_uint2(0u, 0u);
  };
  auto issue = [&](int c, uint2 rec) {
    if (c < blocks) {
      const int e = (blocks - 1 - c) * 32 + lane;
      if (e < n_rec) {
        const float4* src = reinterpret_cast<const float4*>(geom + rec.x);
        float4 (*st)[32] = ring[c & (STAGES - 1)];
        cp_async16(&st[0][lane], src);
        cp_async16(&st[1][lane], src + 1);
        cp_async16(&st[2][lane], src + 2);
      }
      mring[c & (STAGES - 1)][lane] = rec;        // (0, 0) beyond the end: no pixels
    }
    cp_async_commit();
  };
#pragma unroll
  for (int c = 0; c < STAGES - 1; ++c) issue(c, fetch_rec(c));
  uint2 rec_next = fetch_rec(STAGES - 1);

  const uint32_t lt = (1u << lane) - 1u;
  for (int c = 0; c < blocks; ++c) {
    issue(c + STAGES - 1, rec_next);
    rec_next = fetch_rec(c + STAGES);
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    float4 (*st)[32] = ring[c & (STAGES - 1)];
    const uint2* meta = mring[c & (STAGES - 1)];
    const uint2 my_rec = meta[lane];                       // this lane's record of the block (phase 2)
    const int cnt = __popc(my_rec.y);
    // exclusive prefix of the records' pixel counts = first slab slot of each record
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const int my_off = incl - cnt;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    s_off[lane] = (uint32_t)my_off;
    const uint32_t mine = transpose32(my_rec.y, lane);     // records of the block that name this lane's pixel
    __syncwarp();
    // passes over runs of records whose pairs fit the slab, from the BACK of the block (one pass in most blocks)
    (void)total;
    int hi = 32;
    while (hi > 0) {
      // the longest run [lo, hi) with incl[hi-1] - off[lo] <= T2_CAP: lanes lo..hi-1 satisfy it (monotone in lo)
      const int end = __shfl_sync(0xffffffffu, incl, hi - 1);
      const uint32_t fm = __ballot_sync(0xffffffffu, lane < hi && end - my_off <= T2_CAP);
      const int lo = hi - __popc(fm);
      const uint32_t sel = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
      const int base = __shfl_sync(0xffffffffu, my_off, lo);
      // ---- phase 1: lane = pixel -------------------------------------------------------------------------
      uint32_t pend = mine & sel;
      while (__any_sync(0xffffffffu, pend != 0u)) {
        const int r = 31 - __clz(pend);                    // back to front; -1: this pixel has nothing left
        if (r >= 0) {
          pend &= ~(1u << r);
          const uint32_t m = meta[r].y;
          const int slot = (int)s_off[r] - base + __popc(m & lt);
          const float4 a = st[0][r];
          const float4 q = st[1][r];
          const float4 f = st[2][r];
          const float dx = a.x - pxf, dy = a.y - pyf;
          const float G = exp2_blend(gauss_exponent2(q.x, q.y, q.z, dx, dy));
          const float alpha = fminf(ALPHA_CAP, q.w * G);
          const float inv1ma = rcp_blend(1.0f - alpha);
          T *= inv1ma;
          const float w = alpha * T;
          const float phi = f.y * gC0 + f.z * gC1 + f.w * gC2 + f.x * gD + gA;
          const float dL_da = T * phi - Bdot * inv1ma;
          Bdot += phi * w;
          slab[slot] = make_float2(G * dL_da, w);
        }
      }
      __syncwarp();
      // ---- phase 2: lane = record ------------------------------------------------------------------------
      if (lane >= lo && lane < hi && cnt > 0) {
        const float4 a = st[0][lane];
        const float o = st[1][lane].w;
        uint32_t m = my_rec.y;
        const float2* src = slab + (my_off - base);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f, s6 = 0.f, s7 = 0.f, s8 = 0.f, s9 = 0.f;
#ifndef GSB_BWD_T2_PIPE
#define GSB_BWD_T2_PIPE 0
#endif
#if GSB_BWD_T2_PIPE
        // software-pipelined: the next pixel's three shared-memory loads are in flight while this one is summed
        int p = __ffs(m) - 1;
        m &= m - 1;
        float2 uw = *src++;
        float2 xy = s_pxy[p];
        float4 pg = s_pg[p];
        while (true) {
          const bool more = m != 0u;
          float2 uwn = uw, xyn = xy;
          float4 pgn = pg;
          if (more) {
            const int pn = __ffs(m) - 1;
            m &= m - 1;
            uwn = *src++;
            xyn = s_pxy[pn];
            pgn = s_pg[pn];
          }
          const float dx = a.x - xy.x, dy = a.y - xy.y;
          const float sv = o * uw.x;
          const float sx = sv * dx, sy = sv * dy;
          s0 += sx; s1 += sy; s2 += sx * dx; s3 += sx * dy; s4 += sy * dy;
          s5 += uw.x;
          s6 += uw.y * pg.x; s7 += uw.y * pg.y; s8 += uw.y * pg.z; s9 += uw.y * pg.w;
          if (!more) break;
          uw = uwn; xy = xyn; pg = pgn;
        }
#else
        while (m) {
          const int p = __ffs(m) - 1;
          m &= m - 1;
          const float2 uw = *src++;
          const float2 xy = s_pxy[p];
          const float4 pg = s_pg[p];
          const float dx = a.x - xy.x, dy = a.y - xy.y;
          // raw moments of s = dL/dG * G about the splat centre (mapped to d(ndc xy), d(conic) in preprocess_bwd)
          const float sv = o * uw.x;
          const float sx = sv * dx, sy = sv * dy;
          s0 += sx; s1 += sy; s2 += sx * dx; s3 += sx * dy; s4 += sy * dy;
          s5 += uw.x;
          s6 += uw.y * pg.x; s7 += uw.y * pg.y; s8 += uw.y * pg.z; s9 += uw.y * pg.w;
        }
#endif
        float* dst = reinterpret_cast<float*>(ggrad + my_rec.x);
        red_add_v4(dst, s0, s1, s2, s3);
        red_add_v4(dst + 4, s4, s5, s6, s7);
        red_add_v4(dst + 8, s8, s9, 0.0f, 0.0f);
      }
      __syncwarp();                                        // the slab is rewritten by the next pass / block
      hi = lo;
    }
  }
  cp_async_wait<0>();
}

}  // namespace

int launch_render_bwd(const View& v, int P, const Geom* geom, const uint32_t* point_list,
                      const uint2* ranges, const uint32_t* tile_order, const uint32_t* n_contrib,
                      const float* final_T, const uint2* hits, const uint32_t* hit_count, const float* dL_dcolor,
                      const float* dL_ddepth, const float* dL_dalpha, GGrad* ggrad, int kind, bool packed,
                      bool debug, cudaStream_t st) {
  // kind 0: transposed replay (default), 1: per-hit replay (butterfly), 2: rescan (record-free)
  const bool replay = kind == 1;
  GSB_CUDA(cudaMemsetAsync(ggrad, 0, (size_t)P * sizeof(GGrad), st));
  const int T = v.gx * v.gy;
  if (T == 0 || P == 0) return GSB_OK;
#ifndef GSB_BWD_SMEM_PAD
#define GSB_BWD_SMEM_PAD 0
#endif
  // (padding = an occupancy cap for tuning sweeps: shared memory not used by CTAs stays L1)
  constexpr size_t ring_rescan = (size_t)WARPS * STAGES * (3 * 32 * sizeof(float4) + 32 * sizeof(uint32_t));
  constexpr size_t ring_replay = (size_t)WARPS * STAGES * (3 * 32 * sizeof(float4) + 32 * sizeof(uint2));
  constexpr size_t slab_bytes = (size_t)WARPS * PACKED_WARP_BYTES;
  static std::atomic<unsigned long long> configured{0};   // bit per device: the attribute is per device
  int dev = 0;
  GSB_CUDA(cudaGetDevice(&dev));
  if (!(configured.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
    GSB_CUDA(cudaFuncSetAttribute(render_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(ring_rescan + GSB_BWD_SMEM_PAD)));
    GSB_CUDA(cudaFuncSetAttribute(render_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(ring_rescan + slab_bytes + GSB_BWD_SMEM_PAD)));
    GSB_CUDA(cudaFuncSetAttribute(render_bwd_replay_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(ring_replay + GSB_BWD_SMEM_PAD)));
    GSB_CUDA(cudaFuncSetAttribute(render_bwd_replay_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(ring_replay + slab_bytes + GSB_BWD_SMEM_PAD)));
    GSB_CUDA(cudaFuncSetAttribute(render_bwd_transposed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(WARPS * T2_WARP_BYTES + GSB_BWD_SMEM_PAD)));
    configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  if (kind == 0) {
    if (!hits || !hit_count) return GSB_E_INVALID;
    render_bwd_transposed_kernel<<<T, WARPS * 32, WARPS * T2_WARP_BYTES + GSB_BWD_SMEM_PAD, st>>>(
        v, geom, ranges, tile_order, hits, hit_count, final_T, dL_dcolor, dL_ddepth, dL_dalpha, ggrad);
    GSB_POST_LAUNCH(debug, st, "render_bwd_transposed_kernel");
    return GSB_OK;
  }
  if (replay) {
    if (!hits || !hit_count) return GSB_E_INVALID;
    const size_t smem = ring_replay + (packed ? slab_bytes : 0) + GSB_BWD_SMEM_PAD;
    if (packed)
      render_bwd_replay_kernel<true><<<T, WARPS * 32, smem, st>>>(v, geom, ranges, tile_order, hits, hit_count, final_T,
                                                                  dL_dcolor, dL_ddepth, dL_dalpha, ggrad);
    else
      render_bwd_replay_kernel<false><<<T, WARPS * 32, smem, st>>>(v, geom, ranges, tile_order, hits, hit_count, final_T,
                                                                   dL_dcolor, dL_ddepth, dL_dalpha, ggrad);
    GSB_POST_LAUNCH(debug, st, "render_bwd_replay_kernel");
    return GSB_OK;
  }
  const size_t smem = ring_rescan + (packed ? slab_bytes : 0) + GSB_BWD_SMEM_PAD;
  if (packed)
    render_bwd_kernel<true><<<T, WARPS * 32, smem, st>>>(v, geom, point_list, ranges, tile_order, n_contrib, final_T,
                                                         dL_dcolor, dL_ddepth, dL_dalpha, ggrad);
  else
    render_bwd_kernel<false><<<T, WARPS * 32, smem, st>>>(v, geom, point_list, ranges, tile_order, n_contrib, final_T,
                                                          dL_dcolor, dL_ddepth, dL_dalpha, ggrad);
  GSB_POST_LAUNCH(debug, st, "render_bwd_kernel");
#ifdef GSB_BWD_STATS
  {
    unsigned long long h[12];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_bwd_stats, sizeof(h));
    fprintf(stderr, "[bwd stats, cumulative] warp-chunks %llu  candidates %llu  rect-hits %llu  hits-with-valid-lane %llu  "
            "valid lanes %llu  hist(1-2,3-4,5-8,9-16,17-32) %llu %llu %llu %llu %llu\n", h[0], h[1], h[2], h[3], h[4],
            h[5], h[6], h[7], h[8], h[9]);
  }
#endif
  return GSB_OK;
}

}  // namespace gsb
