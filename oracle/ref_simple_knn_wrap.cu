// TEST INFRASTRUCTURE ONLY.  C-ABI shim around the reference's own SimpleKNN::knn
// (gaussiansplatting/submodules/simple-knn/simple_knn.cu:186-221), compiled together with the
// reference source WHERE IT LIES under /root/reference into oracle/_ref/ (see oracle/Makefile).
// Nothing of the reference is copied into this repository; the built .so is git-ignored and only
// the GPU parity tests load it, as the checker for gsb_knn_dist2.
#include <cuda_runtime.h>
#include "simple_knn.h"

extern "C" int ref_simple_knn_dist2(int P, const float* points_dev, float* mean_dist2_dev) {
  if (P <= 0) return 0;
  SimpleKNN::knn(P, reinterpret_cast<float3*>(const_cast<float*>(points_dev)), mean_dist2_dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  return (int)e;
}
