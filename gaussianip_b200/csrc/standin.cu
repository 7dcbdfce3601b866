// REFERENCE-STRUCTURE STAND-IN (not the reference, not the product path).
//
// The rasterizer GaussianIP calls is an external CUDA package whose source is not available here,
// so its speed on a B200 cannot be measured.  These two kernels re-create, from the published
// algorithm and written from scratch, the STRUCTURE of that package's blend
// kernels: one thread per pixel of a 16x16 tile in row-major order, the tile's list staged in
// shared memory 256 instances at a time behind CTA barriers, every instance evaluated by every
// pixel (no culling), CTA-wide early exit, and a backward pass that issues one global atomicAdd
// per gradient component per (pixel, Gaussian).  They read the same Geom records / point_list /
// ranges as the native kernels and exist for two purposes only:
//   * bench.py --variant standin: context for north_star's ">= 2x the reference CUDA rasterizer"
//     target, always labelled as a stand-in;
//   * tests: an independent GPU evaluation with the same expf — contributor counts must match the
//     native kernels EXACTLY, which proves the per-warp culling is lossless.
#include "gsb_common.cuh"

namespace gsb {

namespace {

constexpr int BATCH = 256;

__global__ void __launch_bounds__(256)
standin_fwd_kernel(View v, const Geom* __restrict__ geom, const uint32_t* __restrict__ point_list,
                   const uint2* __restrict__ ranges, float* __restrict__ out_color, float* __restrict__ out_depth,
                   float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib, float* __restrict__ final_T) {
  __shared__ float4 s_a[BATCH], s_q[BATCH], s_f[BATCH];
  const int tile = blockIdx.x;
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int pix_x = tx * TILE_X + (threadIdx.x & 15), pix_y = ty * TILE_Y + (threadIdx.x >> 4);
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const int rounds = (n + BATCH - 1) / BATCH;
  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dz = 0.f, A = 0.f;
  uint32_t last = 0;
  bool done = !inside;
  for (int b = 0; b < rounds; ++b) {
    if (__syncthreads_count(done) == 256) break;
    const int e = b * BATCH + threadIdx.x;
    if (e < n) {
      const float4* src = reinterpret_cast<const float4*>(geom + point_list[range.x + e]);
      s_a[threadIdx.x] = src[0]; s_q[threadIdx.x] = src[1]; s_f[threadIdx.x] = src[2];
    }
    __syncthreads();
    const int cnt = min(BATCH, n - b * BATCH);
    for (int j = 0; !done && j < cnt; ++j) {
      const float4 a = s_a[j], q = s_q[j], f = s_f[j];
      const float dx = a.x - pxf, dy = a.y - pyf;
      const float power = gauss_exponent2(q.x, q.y, q.z, dx, dy);   // log2 of the weight (pre-scaled conic)
      if (power > 0.0f) continue;
      const float alpha = fminf(ALPHA_CAP, q.w * exp2_blend(power));
      if (alpha < ALPHA_MIN) continue;
      const float test_T = T * (1.0f - alpha);
      if (test_T < T_MIN) { done = true; continue; }
      const float w = alpha * T;
      C0 += f.y * w; C1 += f.z * w; C2 += f.w * w; Dz += f.x * w; A += w;
      T = test_T;
      last = (uint32_t)(b * BATCH + j + 1);
    }
  }
  if (inside) {
    const size_t hw = (size_t)v.H * v.W, pix = (size_t)pix_y * v.W + pix_x;
    out_color[pix] = C0 + T * v.bg[0];
    out_color[hw + pix] = C1 + T * v.bg[1];
    out_color[2 * hw + pix] = C2 + T * v.bg[2];
    out_depth[pix] = Dz; out_alpha[pix] = A; n_contrib[pix] = last; final_T[pix] = T;
  }
}

__global__ void __launch_bounds__(256)
standin_bwd_kernel(View v, const Geom* __restrict__ geom, const uint32_t* __restrict__ point_list,
                   const uint2* __restrict__ ranges, const uint32_t* __restrict__ n_contrib,
                   const float* __restrict__ final_T, const float* __restrict__ dL_dcolor,
                   const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha,
                   GGrad* __restrict__ ggrad) {
  __shared__ float4 s_a[BATCH], s_q[BATCH], s_f[BATCH];
  __shared__ uint32_t s_id[BATCH];
  const int tile = blockIdx.x;
  const int tx = tile % v.gx, ty = tile / v.gx;
  const int pix_x = tx * TILE_X + (threadIdx.x & 15), pix_y = ty * TILE_Y + (threadIdx.x >> 4);
  const bool inside = pix_x < v.W && pix_y < v.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const size_t hw = (size_t)v.H * v.W, pix = (size_t)pix_y * v.W + pix_x;
  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const int rounds = (n + BATCH - 1) / BATCH;
  const float T_final = inside ? final_T[pix] : 0.f;
  float T = T_final;
  const uint32_t my_last = inside ? n_contrib[pix] : 0u;
  const float gC0 = inside ? dL_dcolor[pix] : 0.f, gC1 = inside ? dL_dcolor[hw + pix] : 0.f,
              gC2 = inside ? dL_dcolor[2 * hw + pix] : 0.f;
  const float gD = inside ? dL_ddepth[pix] : 0.f, gA = inside ? dL_dalpha[pix] : 0.f;
  const float bg_dot = v.bg[0] * gC0 + v.bg[1] * gC1 + v.bg[2] * gC2;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accD = 0.f, accA = 0.f;
  float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, lD = 0.f;
  // the whole list is walked back to front, as the published backward does
  for (int b = rounds - 1; b >= 0; --b) {
    __syncthreads();
    const int e = b * BATCH + threadIdx.x;
    if (e < n) {
      const uint32_t id = point_list[range.x + e];
      const float4* src = reinterpret_cast<const float4*>(geom + id);
      s_a[threadIdx.x] = src[0]; s_q[threadIdx.x] = src[1]; s_f[threadIdx.x] = src[2]; s_id[threadIdx.x] = id;
    }
    __syncthreads();
    const int cnt = min(BATCH, n - b * BATCH);
    for (int j = cnt - 1; j >= 0; --j) {
      const uint32_t pos = (uint32_t)(b * BATCH + j + 1);
      if (pos > my_last) continue;
      const float4 a = s_a[j], q = s_q[j], f = s_f[j];
      const float dx = a.x - pxf, dy = a.y - pyf;
      const float power = gauss_exponent2(q.x, q.y, q.z, dx, dy);   // log2 of the weight (pre-scaled conic)
      if (power > 0.0f) continue;
      const float G = exp2_blend(power);
      const float alpha = fminf(ALPHA_CAP, q.w * G);
      if (alpha < ALPHA_MIN) continue;
      T = T / (1.0f - alpha);
      const float w = alpha * T;
      acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0; lc0 = f.y;
      acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1; lc1 = f.z;
      acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2; lc2 = f.w;
      accD = last_alpha * lD + (1.0f - last_alpha) * accD; lD = f.x;
      accA = last_alpha + (1.0f - last_alpha) * accA;
      float dL_da = (f.y - acc0) * gC0 + (f.z - acc1) * gC1 + (f.w - acc2) * gC2 + (f.x - accD) * gD +
                    (1.0f - accA) * gA;
      dL_da *= T;
      last_alpha = alpha;
      dL_da += (-T_final / (1.0f - alpha)) * bg_dot;
      const float s_ = q.w * dL_da * G;
      float* dst = reinterpret_cast<float*>(ggrad + s_id[j]);   // same moment layout as the native kernel
      atomicAdd(dst + 0, s_ * dx);
      atomicAdd(dst + 1, s_ * dy);
      atomicAdd(dst + 2, s_ * dx * dx);
      atomicAdd(dst + 3, s_ * dx * dy);
      atomicAdd(dst + 4, s_ * dy * dy);
      atomicAdd(dst + 5, G * dL_da);
      atomicAdd(dst + 6, w * gD);
      atomicAdd(dst + 7, w * gC0);
      atomicAdd(dst + 8, w * gC1);
      atomicAdd(dst + 9, w * gC2);
    }
  }
}

}  // namespace

int launch_standin_fwd(const View& v, const Geom* geom, const uint32_t* point_list, const uint2* ranges,
                       float* color, float* depth, float* alpha, uint32_t* n_contrib, float* final_T, bool debug,
                       cudaStream_t st) {
  const int T = v.gx * v.gy;
  if (T == 0) return GSB_OK;
  standin_fwd_kernel<<<T, 256, 0, st>>>(v, geom, point_list, ranges, color, depth, alpha, n_contrib, final_T);
  GSB_POST_LAUNCH(debug, st, "standin_fwd_kernel");
  return GSB_OK;
}

int launch_standin_bwd(const View& v, int P, const Geom* geom, const uint32_t* point_list, const uint2* ranges,
                       const uint32_t* n_contrib, const float* final_T, const float* dL_dcolor,
                       const float* dL_ddepth, const float* dL_dalpha, GGrad* ggrad, bool debug, cudaStream_t st) {
  GSB_CUDA(cudaMemsetAsync(ggrad, 0, (size_t)P * sizeof(GGrad), st));
  const int T = v.gx * v.gy;
  if (T == 0 || P == 0) return GSB_OK;
  standin_bwd_kernel<<<T, 256, 0, st>>>(v, geom, point_list, ranges, n_contrib, final_T, dL_dcolor, dL_ddepth,
                                        dL_dalpha, ggrad);
  GSB_POST_LAUNCH(debug, st, "standin_bwd_kernel");
  return GSB_OK;
}

}  // namespace gsb
