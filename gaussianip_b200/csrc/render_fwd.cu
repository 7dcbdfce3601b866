// Stage 3: per-tile front-to-back alpha blending.  Replaces renderCUDA (forward) of the
// external operator ("Forward blend").
//
// One CTA per 16x16 tile, 8 warps; warp w owns an 8x4 pixel block so that a small splat
// overlaps few warps.  Instances are gathered (point_list -> 48 B Geom record, L2-resident)
// straight into shared memory with cp.async, double-buffered in batches of 256 so the gather
// of batch b+1 overlaps the blend of batch b.  Each warp first votes which of the 256 staged
// instances can reach any of ITS 32 pixels (conservative extent test, see preprocess.cu) and
// only evaluates those, in list order, so results are identical to evaluating all of them.
// A warp stops when all of its pixels are saturated; the CTA stops when all warps have.
#include "gsb_common.cuh"

namespace gsb {

namespace {

constexpr int BATCH = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__global__ void __launch_bounds__(256)
render_fwd_kernel(View v, const Geom* __restrict__ geom, const uint32_t* __restrict__ point_list,
                  const uint2* __restrict__ ranges, float* __restrict__ out_color,
                  float* __restrict__ out_depth, float* __restrict__ out_alpha,
                  uint32_t* __restrict__ n_contrib, float* __restrict__ final_T) {
  __shared__ float4 s_a[2][BATCH];  // x, y, conA, conB
  __shared__ float4 s_b[2][BATCH];  // conC, opacity, depth, r
  __shared__ float4 s_c[2][BATCH];  // g, b, extx, exty

  const int tile = blockIdx.x;
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int pix_x = tx * TILE_X + wx + (lane & 7);
  const int pix_y = ty * TILE_Y + wy + (lane >> 3);
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const float cxw = (float)(tx * TILE_X + wx) + 3.5f, cyw = (float)(ty * TILE_Y + wy) + 1.5f;

  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const int rounds = (n + BATCH - 1) / BATCH;

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dz = 0.f, A = 0.f;
  uint32_t last = 0;
  bool done = !inside;
  bool warp_done = __all_sync(0xffffffffu, done);

  auto issue = [&](int b, uint32_t gid) {
    const int e = b * BATCH + tid;
    if (e < n) {
      const float4* src = reinterpret_cast<const float4*>(geom + gid);
      const int buf = b & 1;
      cp_async16(&s_a[buf][tid], src);
      cp_async16(&s_b[buf][tid], src + 1);
      cp_async16(&s_c[buf][tid], src + 2);
    }
    cp_async_commit();
  };
  auto fetch_gid = [&](int b) -> uint32_t {
    const int e = b * BATCH + tid;
    return (b < rounds && e < n) ? point_list[range.x + e] : 0u;
  };

  if (rounds > 0) {
    issue(0, fetch_gid(0));
    uint32_t gid_next = fetch_gid(1);
    for (int b = 0; b < rounds; ++b) {
      if (b + 1 < rounds) {
        issue(b + 1, gid_next);
        gid_next = fetch_gid(b + 2);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const int buf = b & 1;
      const int cnt = min(BATCH, n - b * BATCH);
      if (!warp_done) {
        for (int j = 0; j < BATCH / 32; ++j) {
          if (j * 32 >= cnt) break;
          const int e = j * 32 + lane;
          bool hit = false;
          if (e < cnt) {
            const float4 a = s_a[buf][e];
            const float4 c = s_c[buf][e];
            hit = (fabsf(a.x - cxw) <= c.z + 3.5f) && (fabsf(a.y - cyw) <= c.w + 1.5f);
          }
          uint32_t mask = __ballot_sync(0xffffffffu, hit);
          while (mask) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            const int e2 = j * 32 + k;
            const float4 a = s_a[buf][e2];
            const float4 q = s_b[buf][e2];
            const float4 c = s_c[buf][e2];
            if (!done) {
              const float dx = a.x - pxf, dy = a.y - pyf;
              const float power = -0.5f * (a.z * dx * dx + q.x * dy * dy) - a.w * dx * dy;
              if (power <= 0.0f) {
                const float alpha = fminf(ALPHA_CAP, q.y * __expf(power));
                if (alpha >= ALPHA_MIN) {
                  const float test_T = T * (1.0f - alpha);
                  if (test_T < T_MIN) {
                    done = true;
                  } else {
                    const float w = alpha * T;
                    C0 += q.w * w; C1 += c.x * w; C2 += c.y * w;
                    Dz += q.z * w; A += w;
                    T = test_T;
                    last = (uint32_t)(b * BATCH + e2 + 1);
                  }
                }
              }
            }
          }
          if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
        }
      }
      if (__syncthreads_and(warp_done)) break;
    }
    cp_async_wait<0>();
  }

  if (inside) {
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
                      const uint2* ranges, float* color, float* depth, float* alpha,
                      uint32_t* n_contrib, float* final_T, bool debug, cudaStream_t st) {
  const int T = v.gx * v.gy;
  if (T == 0) return GSB_OK;
  render_fwd_kernel<<<T, 256, 0, st>>>(v, geom, point_list, ranges, color, depth, alpha, n_contrib, final_T);
  GSB_POST_LAUNCH(debug, st, "render_fwd_kernel");
  return GSB_OK;
}

}  // namespace gsb
