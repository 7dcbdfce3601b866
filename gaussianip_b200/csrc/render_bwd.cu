// Backward stage 1: per-pixel reverse (back-to-front) blending.  Replaces renderCUDA
// (backward) of the external operator ("Backward blend").
//
// Same tiling as render_fwd.cu (one CTA per 16x16 tile, warp = 8x4 pixels, cp.async
// double-buffered gathers, per-warp conservative culling).  Differences that matter:
//  * the walk starts at the LAST instance any pixel of the tile actually blended
//    (block max of n_contrib), not at the end of the tile's list;
//  * the ten per-Gaussian partial gradients of the 32 pixels of a warp are summed with a
//    13-shuffle multi-value butterfly and leave the SM as ONE red.global per component per
//    (warp, Gaussian) — 10 lanes hitting one 48 B GGrad record — instead of ~10 atomicAdd
//    per (pixel, Gaussian).
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

__global__ void __launch_bounds__(256)
render_bwd_kernel(View v, const Geom* __restrict__ geom, const uint32_t* __restrict__ point_list,
                  const uint2* __restrict__ ranges, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ final_T, const float* __restrict__ dL_dcolor,
                  const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha,
                  GGrad* __restrict__ ggrad) {
  __shared__ float4 s_a[2][BATCH];
  __shared__ float4 s_b[2][BATCH];
  __shared__ float4 s_c[2][BATCH];
  __shared__ uint32_t s_gid[2][BATCH];
  __shared__ uint32_t s_max[8];

  const int tile = blockIdx.x;
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int pix_x = tx * TILE_X + wx + (lane & 7);
  const int pix_y = ty * TILE_Y + wy + (lane >> 3);
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const float cxw = (float)(tx * TILE_X + wx) + 3.5f, cyw = (float)(ty * TILE_Y + wy) + 1.5f;
  const size_t hw = (size_t)v.H * v.W;
  const size_t pix = (size_t)pix_y * v.W + pix_x;

  const uint2 range = ranges[tile];
  const uint32_t my_last = inside ? n_contrib[pix] : 0u;
  const uint32_t warp_last = __reduce_max_sync(0xffffffffu, my_last);
  if (lane == 0) s_max[warp] = warp_last;
  __syncthreads();
  uint32_t n_eff = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) n_eff = max(n_eff, s_max[w]);
  if (n_eff == 0) return;
  const int n = (int)n_eff;  // <= range.y - range.x
  const int rounds = (n + BATCH - 1) / BATCH;

  const float T_final = inside ? final_T[pix] : 0.0f;
  float T = T_final;
  const float gC0 = inside ? dL_dcolor[pix] : 0.f, gC1 = inside ? dL_dcolor[hw + pix] : 0.f,
              gC2 = inside ? dL_dcolor[2 * hw + pix] : 0.f;
  const float gD = inside ? dL_ddepth[pix] : 0.f, gA = inside ? dL_dalpha[pix] : 0.f;
  const float bg_dot = v.bg[0] * gC0 + v.bg[1] * gC1 + v.bg[2] * gC2;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accD = 0.f, accA = 0.f;
  float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, lD = 0.f;
  const float half_W = 0.5f * (float)v.W, half_H = 0.5f * (float)v.H;

  auto issue = [&](int b, uint32_t gid) {
    const int e = b * BATCH + tid;
    if (e < n) {
      const float4* src = reinterpret_cast<const float4*>(geom + gid);
      const int buf = b & 1;
      cp_async16(&s_a[buf][tid], src);
      cp_async16(&s_b[buf][tid], src + 1);
      cp_async16(&s_c[buf][tid], src + 2);
      s_gid[buf][tid] = gid;
    }
    cp_async_commit();
  };
  auto fetch_gid = [&](int b) -> uint32_t {
    const int e = b * BATCH + tid;
    return (b >= 0 && e < n) ? point_list[range.x + e] : 0u;
  };

  issue(rounds - 1, fetch_gid(rounds - 1));
  uint32_t gid_next = fetch_gid(rounds - 2);
  for (int b = rounds - 1; b >= 0; --b) {
    if (b > 0) {
      issue(b - 1, gid_next);
      gid_next = fetch_gid(b - 2);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int buf = b & 1;
    const int cnt = min(BATCH, n - b * BATCH);
    if ((uint32_t)(b * BATCH) < warp_last) {
      for (int j = (cnt - 1) / 32; j >= 0; --j) {
        if ((uint32_t)(b * BATCH + j * 32) >= warp_last) continue;
        const int e = j * 32 + lane;
        bool hit = false;
        if (e < cnt) {
          const float4 a = s_a[buf][e];
          const float4 c = s_c[buf][e];
          hit = (fabsf(a.x - cxw) <= c.z + 3.5f) && (fabsf(a.y - cyw) <= c.w + 1.5f);
        }
        uint32_t mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = 31 - __clz(mask);
          mask &= ~(1u << k);
          const int e2 = j * 32 + k;
          const uint32_t pos = (uint32_t)(b * BATCH + e2 + 1);
          const float4 a = s_a[buf][e2];
          const float4 q = s_b[buf][e2];
          const float4 c = s_c[buf][e2];
          float g[12];
#pragma unroll
          for (int i = 0; i < 12; ++i) g[i] = 0.0f;
          bool contributes = false;
          if (pos <= my_last) {
            const float dx = a.x - pxf, dy = a.y - pyf;
            const float power = -0.5f * (a.z * dx * dx + q.x * dy * dy) - a.w * dx * dy;
            if (power <= 0.0f) {
              const float G = __expf(power);
              const float alpha = fminf(ALPHA_CAP, q.y * G);
              if (alpha >= ALPHA_MIN) {
                contributes = true;
                T = T / (1.0f - alpha);
                const float w = alpha * T;
                float dL_da = 0.0f;
                acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0; lc0 = q.w;
                acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1; lc1 = c.x;
                acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2; lc2 = c.y;
                accD = last_alpha * lD + (1.0f - last_alpha) * accD; lD = q.z;
                accA = last_alpha + (1.0f - last_alpha) * accA;
                dL_da += (q.w - acc0) * gC0 + (c.x - acc1) * gC1 + (c.y - acc2) * gC2;
                dL_da += (q.z - accD) * gD;
                dL_da += (1.0f - accA) * gA;
                dL_da *= T;
                last_alpha = alpha;
                dL_da += (-T_final / (1.0f - alpha)) * bg_dot;
                const float dL_dG = q.y * dL_da;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_dx = -gdx * a.z - gdy * a.w;
                const float dG_dy = -gdy * q.x - gdx * a.w;
                g[0] = dL_dG * dG_dx * half_W;
                g[1] = dL_dG * dG_dy * half_H;
                g[2] = -0.5f * gdx * dx * dL_dG;
                g[3] = -gdx * dy * dL_dG;
                g[4] = -0.5f * gdy * dy * dL_dG;
                g[5] = G * dL_da;
                g[6] = w * gD;
                g[7] = w * gC0;
                g[8] = w * gC1;
                g[9] = w * gC2;
              }
            }
          }
          if (!__any_sync(0xffffffffu, contributes)) continue;
          const float total = butterfly12(g, lane);
          const int slot = ((lane & 16) ? 6 : 0) + ((lane & 8) ? 3 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
          const bool pad = (lane & 4) && (lane & 2);
          if (!(lane & 1) && !pad && slot < 10 && total != 0.0f) {
            float* dst = reinterpret_cast<float*>(ggrad + s_gid[buf][e2]) + slot;
            atomicAdd(dst, total);
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

int launch_render_bwd(const View& v, int P, const Geom* geom, const uint32_t* point_list,
                      const uint2* ranges, const uint32_t* n_contrib, const float* final_T,
                      const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                      GGrad* ggrad, bool debug, cudaStream_t st) {
  GSB_CUDA(cudaMemsetAsync(ggrad, 0, (size_t)P * sizeof(GGrad), st));
  const int T = v.gx * v.gy;
  if (T == 0 || P == 0) return GSB_OK;
  render_bwd_kernel<<<T, 256, 0, st>>>(v, geom, point_list, ranges, n_contrib, final_T, dL_dcolor, dL_ddepth,
                                       dL_dalpha, ggrad);
  GSB_POST_LAUNCH(debug, st, "render_bwd_kernel");
  return GSB_OK;
}

}  // namespace gsb
