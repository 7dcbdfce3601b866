// Micro-benchmark: scalar FFMA vs packed FFMA2 issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 av = make_float2(a, a * 1.0001f), bv = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { x[i].x = fmaf(x[i].x, av.x, bv.x); x[i].y = fmaf(x[i].y, av.y, bv.y); }
      else x[i] = __ffma2_rn(x[i], av, bv);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); else k<1><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 8 * 256 * (double)iters * 16;
      if (rep) printf("%s: %.3f ms  %.1f TFLOP/s (fp32 FMA = 2 flop)\n", mode ? "FFMA2 packed" : "FFMA scalar", ms, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
