// Probe for the Blackwell row-gather TMA (cp.async.bulk.tensor.2d ... tile::gather4) on the rasterizer's 48-byte
// Geom rows: which tensor-map box shape the instruction accepts, whether the gathered rows are
// right, and what a warp-private 32-row stage costs against 3 x LDGSTS (cp.async 16 B) per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_gather4_probe tma_gather4_probe.cu && ./tma_gather4_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ROW_FLOATS = 12;      // Geom: 48 B
constexpr int WARPS = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a transaction count that never completes (wrong expect_tx) traps instead of hanging the GPU
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
    if (spins > (1u << 24)) asm volatile("trap;");
  }
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}

// Every warp gathers `chunks` x 32 rows (ids from `idx`) into its own stage and accumulates a checksum; with
// chunks == 1 the rows themselves are written out for the correctness check.
template <bool TMA, int PITCH>
__global__ void __launch_bounds__(WARPS * 32)
gather_kernel(const __grid_constant__ CUtensorMap tm, const float* __restrict__ table, const uint32_t* __restrict__ idx,
              int chunks, float* __restrict__ rows_out, float* __restrict__ sums) {
  // a gather4 destination must be 128-byte aligned: the 4 rows of a group (192 B) sit in a 256-byte slot
  __shared__ __align__(128) float stage[WARPS][2][8 * 64];
  __shared__ __align__(8) uint64_t bars[WARPS][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t* my_idx = idx + ((size_t)blockIdx.x * WARPS + warp) * (size_t)chunks * 32;
  if (TMA) {
    if (lane == 0) { mbar_init(&bars[warp][0], 1); mbar_init(&bars[warp][1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  auto issue = [&](int c) {
    if (c >= chunks) { if (!TMA) asm volatile("cp.async.commit_group;\n" ::); return; }
    const uint32_t id = my_idx[c * 32 + lane];
    float* dst = stage[warp][c & 1];
    if (TMA) {
      const int id1 = __shfl_down_sync(0xffffffffu, id, 1), id2 = __shfl_down_sync(0xffffffffu, id, 2),
                id3 = __shfl_down_sync(0xffffffffu, id, 3);
      if (lane == 0) mbar_expect_tx(&bars[warp][c & 1], 32 * PITCH * 4);
      __syncwarp();
      if ((lane & 3) == 0) tma_gather4(dst + (lane >> 2) * 64, &tm, 0, (int)id, id1, id2, id3, &bars[warp][c & 1]);
    } else {
      const float4* src = reinterpret_cast<const float4*>(table + (size_t)id * ROW_FLOATS);
      float4* d4 = reinterpret_cast<float4*>(dst + (lane >> 2) * 64 + (lane & 3) * PITCH);
      cp_async16(d4, src); cp_async16(d4 + 1, src + 1); cp_async16(d4 + 2, src + 2);
      asm volatile("cp.async.commit_group;\n" ::);
    }
  };
  float acc = 0.f;
  issue(0);
  for (int c = 0; c < chunks; ++c) {
    issue(c + 1);
    if (TMA) mbar_wait(&bars[warp][c & 1], (uint32_t)((c >> 1) & 1));
    else { asm volatile("cp.async.wait_group 1;\n" ::); }
    __syncwarp();
    const float* st = stage[warp][c & 1];
    // consume like the blend kernels: broadcast reads of every row's first float4
    for (int k = 0; k < 32; k += 4) {
      const float4 a = *reinterpret_cast<const float4*>(st + (k >> 2) * 64 + (k & 3) * PITCH);
      acc += a.x + a.y * 0.5f + a.z * 0.25f + a.w;
    }
    if (chunks == 1 && rows_out)
      for (int j = 0; j < ROW_FLOATS; ++j)
        rows_out[(((size_t)blockIdx.x * WARPS + warp) * 32 + lane) * ROW_FLOATS + j] = st[(lane >> 2) * 64 + (lane & 3) * PITCH + j];
    __syncwarp();
  }
  if (!TMA) asm volatile("cp.async.wait_group 0;\n" ::);
  if (lane == 0) sums[blockIdx.x * WARPS + warp] = acc;
}

int main() {
  const int P = 1 << 20;
  std::vector<float> h((size_t)P * ROW_FLOATS);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 1000) * 0.001f + (float)(i / ROW_FLOATS);
  float* d_table; CK(cudaMalloc(&d_table, h.size() * 4)); CK(cudaMemcpy(d_table, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
  const int grid = 148 * 4, chunks_timed = 256;
  const size_t n_idx = (size_t)grid * WARPS * chunks_timed * 32;
  std::vector<uint32_t> hidx(n_idx);
  uint32_t s = 12345u;
  for (auto& v : hidx) { s = s * 1664525u + 1013904223u; v = (s >> 8) % P; }
  uint32_t* d_idx; CK(cudaMalloc(&d_idx, n_idx * 4)); CK(cudaMemcpy(d_idx, hidx.data(), n_idx * 4, cudaMemcpyHostToDevice));
  float *d_rows, *d_sums;
  CK(cudaMalloc(&d_rows, (size_t)grid * WARPS * 32 * ROW_FLOATS * 4)); CK(cudaMalloc(&d_sums, (size_t)grid * WARPS * 4));
  for (int boxw : {12, 16}) {
    const int boxh = 1;
    CUtensorMap tm;
    cuuint64_t gdim[2] = {ROW_FLOATS, (cuuint64_t)P};
    cuuint64_t gstr[1] = {ROW_FLOATS * 4};
    cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_table, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {%d,%d} on 12-float rows: encode -> %d\n", boxw, boxh, (int)r);
    if (r != CUDA_SUCCESS) continue;
    // correctness: one chunk per warp
    CK(cudaMemset(d_rows, 0, (size_t)grid * WARPS * 32 * ROW_FLOATS * 4));
    if (boxw == 12) gather_kernel<true, 12><<<grid, WARPS * 32>>>(tm, d_table, d_idx, 1, d_rows, d_sums);
    else gather_kernel<true, 16><<<grid, WARPS * 32>>>(tm, d_table, d_idx, 1, d_rows, d_sums);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  gather4 kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> rows((size_t)grid * WARPS * 32 * ROW_FLOATS);
    CK(cudaMemcpy(rows.data(), d_rows, rows.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t w = 0; w < (size_t)grid * WARPS; ++w)
      for (int l = 0; l < 32; ++l)
        for (int j = 0; j < ROW_FLOATS; ++j)
          if (rows[(w * 32 + l) * ROW_FLOATS + j] != h[(size_t)hidx[w * 32 + l] * ROW_FLOATS + j]) ++bad;   // chunks == 1 layout
    printf("  rows gathered: %zu mismatching floats of %zu\n", bad, rows.size());
    if (bad) continue;
    // timing: TMA gather4 vs LDGSTS, same consumption
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      float ms_t = 0.f, ms_l = 0.f;
      auto run_t = [&]() {
        if (boxw == 12) gather_kernel<true, 12><<<grid, WARPS * 32>>>(tm, d_table, d_idx, chunks_timed, nullptr, d_sums);
        else gather_kernel<true, 16><<<grid, WARPS * 32>>>(tm, d_table, d_idx, chunks_timed, nullptr, d_sums);
      };
      run_t();
      CK(cudaEventRecord(a)); run_t(); CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms_t, a, b));
      gather_kernel<false, 12><<<grid, WARPS * 32>>>(tm, d_table, d_idx, chunks_timed, nullptr, d_sums);
      CK(cudaEventRecord(a)); gather_kernel<false, 12><<<grid, WARPS * 32>>>(tm, d_table, d_idx, chunks_timed, nullptr, d_sums); CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms_l, a, b));
      const double rows_total = (double)grid * WARPS * chunks_timed * 32;
      printf("  %.0f M rows of 48 B: gather4 %.3f ms (%.1f G rows/s), LDGSTS %.3f ms (%.1f G rows/s)\n", rows_total * 1e-6, ms_t,
             rows_total / ms_t * 1e-6, ms_l, rows_total / ms_l * 1e-6);
    }
  }
  return 0;
}
