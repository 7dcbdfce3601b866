// Row (f1) of SURVEY.md §8: the step that follows the hot path every iteration — densification
// statistics + Adam for the six parameter groups — as ONE kernel over all groups.
// Reference: threestudio/systems/GaussianIP.py:446-475 (statistics), gaussiansplatting/scene/
// gaussian_model.py:138-159 (torch.optim.Adam, six groups with their own lr, eps = 1e-15) and :420-422
// (add_densification_stats).  Element-wise math and rounding order follow torch.optim.Adam
// (lerp for exp_avg, mul+addcmul for exp_avg_sq, sqrt / bias_correction2_sqrt + eps, addcdiv).
// HBM-bound: 28 B per parameter element (read p, g, m, v; write p, m, v).
#include "gsb_common.cuh"

namespace gsb {

namespace {

struct AdamBatch {
  int G;
  float* p[GSB_ADAM_MAX_GROUPS];
  const float* g[GSB_ADAM_MAX_GROUPS];
  float* m[GSB_ADAM_MAX_GROUPS];
  float* v[GSB_ADAM_MAX_GROUPS];
  long long n[GSB_ADAM_MAX_GROUPS];
  long long first_block[GSB_ADAM_MAX_GROUPS + 1];   // blocks of group k: [first_block[k], first_block[k+1])
  float step_size[GSB_ADAM_MAX_GROUPS];              // lr / (1 - beta1^t)
  float bc2_sqrt[GSB_ADAM_MAX_GROUPS];               // sqrt(1 - beta2^t), t = the group's own step count
  float beta1, beta2, eps, grad_scale;               // grad_scale multiplies g first (AMP unscale), 1 = off
  float omb1, omb2;                                  // 1 - beta, rounded from double as torch does
  // densification statistics (optional, stats_n = 0 disables): one extra range of blocks
  long long stats_n;
  const float* viewspace_grad;   // [P,3] summed over views (and ranks)
  const int32_t* radii;          // [P] max over views (and ranks)
  float* xyz_gradient_accum;     // [P,1]
  float* denom;                  // [P,1]
  float* max_radii2D;            // [P]
};

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC = 4;
constexpr int ADAM_PER_BLOCK = ADAM_THREADS * ADAM_VEC;

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float step_size, float bc2_sqrt,
                                         const AdamBatch& B) {
  g *= B.grad_scale;
  m = m + (g - m) * B.omb1;                                 // exp_avg.lerp_(grad, 1 - beta1)
  v = v * B.beta2 + B.omb2 * g * g;                         // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + B.eps;
  p = p - step_size * (m / denom);                          // param.addcdiv_(exp_avg, denom, -step_size)
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_stats_kernel(AdamBatch B) {
  const long long blk = blockIdx.x;
  if (blk >= B.first_block[B.G]) {
    // ---- densification statistics (GaussianIP.py:456-457, gaussian_model.py:420-422) ----
    const long long i = (blk - B.first_block[B.G]) * ADAM_THREADS + threadIdx.x;
    if (i >= B.stats_n) return;
    const int r = B.radii[i];
    if (r > 0) {
      const float gx = B.viewspace_grad[3 * i], gy = B.viewspace_grad[3 * i + 1];
      B.xyz_gradient_accum[i] += sqrtf(gx * gx + gy * gy);
      B.denom[i] += 1.0f;
      B.max_radii2D[i] = fmaxf(B.max_radii2D[i], (float)r);
    }
    return;
  }
  int k = 0;
#pragma unroll
  for (int j = 1; j < GSB_ADAM_MAX_GROUPS; ++j)
    if (j < B.G && blk >= B.first_block[j]) k = j;
  const long long base = (blk - B.first_block[k]) * ADAM_PER_BLOCK + (long long)threadIdx.x * ADAM_VEC;
  const long long n = B.n[k];
  if (base >= n) return;
  float* p = B.p[k] + base;
  const float* g = B.g[k] + base;
  float* m = B.m[k] + base;
  float* v = B.v[k] + base;
  const float ss = B.step_size[k], bq = B.bc2_sqrt[k];
  const bool vec = base + ADAM_VEC <= n &&
                   (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  if (vec) {
    float4 P4 = *reinterpret_cast<float4*>(p), M4 = *reinterpret_cast<float4*>(m), V4 = *reinterpret_cast<float4*>(v);
    const float4 G4 = *reinterpret_cast<const float4*>(g);
    adam_one(P4.x, G4.x, M4.x, V4.x, ss, bq, B); adam_one(P4.y, G4.y, M4.y, V4.y, ss, bq, B);
    adam_one(P4.z, G4.z, M4.z, V4.z, ss, bq, B); adam_one(P4.w, G4.w, M4.w, V4.w, ss, bq, B);
    *reinterpret_cast<float4*>(p) = P4; *reinterpret_cast<float4*>(m) = M4; *reinterpret_cast<float4*>(v) = V4;
  } else {
    for (int j = 0; j < ADAM_VEC && base + j < n; ++j) adam_one(p[j], g[j], m[j], v[j], ss, bq, B);
  }
}

}  // namespace

int launch_adam_stats(int G, float* const* p, const float* const* g, float* const* m, float* const* v,
                      const long long* n, const float* lr, double beta1, double beta2, double eps,
                      const long long* steps, float grad_scale, long long stats_n, const float* viewspace_grad, const int32_t* radii,
                      float* xyz_gradient_accum, float* denom, float* max_radii2D, cudaStream_t st) {
  if (G < 0 || G > GSB_ADAM_MAX_GROUPS || (G > 0 && !steps)) return GSB_E_INVALID;
  AdamBatch B;
  B.G = G;
  long long blocks = 0;
  for (int k = 0; k < G; ++k) {
    if (n[k] < 0 || (n[k] > 0 && (!p[k] || !g[k] || !m[k] || !v[k]))) return GSB_E_INVALID;
    B.p[k] = p[k]; B.g[k] = g[k]; B.m[k] = m[k]; B.v[k] = v[k]; B.n[k] = n[k];
    B.first_block[k] = blocks;
    blocks += (n[k] + ADAM_PER_BLOCK - 1) / ADAM_PER_BLOCK;
    if (steps[k] < 1) return GSB_E_INVALID;
    const double bc1 = 1.0 - pow(beta1, (double)steps[k]);
    const double bc2 = 1.0 - pow(beta2, (double)steps[k]);
    B.step_size[k] = (float)((double)lr[k] / bc1);
    B.bc2_sqrt[k] = (float)sqrt(bc2);
  }
  for (int k = G; k <= GSB_ADAM_MAX_GROUPS; ++k) B.first_block[k] = blocks;
  B.omb1 = (float)(1.0 - beta1); B.omb2 = (float)(1.0 - beta2);
  B.beta1 = (float)beta1; B.beta2 = (float)beta2; B.eps = (float)eps; B.grad_scale = grad_scale;
  B.stats_n = stats_n;
  B.viewspace_grad = viewspace_grad; B.radii = radii; B.xyz_gradient_accum = xyz_gradient_accum;
  B.denom = denom; B.max_radii2D = max_radii2D;
  if (stats_n > 0) {
    if (!viewspace_grad || !radii || !xyz_gradient_accum || !denom || !max_radii2D) return GSB_E_INVALID;
    blocks += (stats_n + ADAM_THREADS - 1) / ADAM_THREADS;
  }
  if (blocks == 0) return GSB_OK;
  if (blocks > 0x7fffffffll) return GSB_E_UNSUPPORTED;
  adam_stats_kernel<<<(unsigned)blocks, ADAM_THREADS, 0, st>>>(B);
  GSB_POST_LAUNCH(false, st, "adam_stats_kernel");
  return GSB_OK;
}

}  // namespace gsb
