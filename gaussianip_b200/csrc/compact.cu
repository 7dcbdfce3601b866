// the data movement of densify / clone / split / prune
// (gaussiansplatting/scene/gaussian_model.py:281-418).  The reference moves every parameter tensor,
// both Adam moments of each and the three statistics tensors one at a time with boolean-mask
// indexing (`t[mask]` = nonzero + gather + a host synchronisation each, ~21 per prune) and
// `torch.cat` (two full copies per densify).  Here the mask becomes an index list once
// (stable stream compaction = torch.nonzero order) and ONE launch gathers the rows of all
// tensors; appended rows that must start at zero (Adam moments of new points) are zero-filled
// by the same launch.  Pure HBM-bound byte movement: coalesced along rows, grid sized by the
// tensor with the most elements, no host round trip except the single row count the caller
// needs to size the outputs.
#include "gsb_common.cuh"

namespace gsb {

constexpr int CP_THREADS = 256;
constexpr int CP_TILE = 2048;        // mask bytes per block in the index passes

struct GatherBatch {
  const float* src[GSB_GATHER_MAX_TENSORS];
  float* dst[GSB_GATHER_MAX_TENSORS];
  int width[GSB_GATHER_MAX_TENSORS];
};

__device__ __forceinline__ int block_count(const uint8_t* __restrict__ mask, long long n, long long base, int* warp_tot,
                                           int (&mine)[CP_TILE / CP_THREADS], int& before) {
  // element e of the tile is owned by thread e / 8 (8 consecutive bytes per thread keeps the order)
  constexpr int PER = CP_TILE / CP_THREADS;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const long long i = base + (long long)threadIdx.x * PER + j;
    mine[j] = (i < n && mask[i]) ? 1 : 0;
    cnt += mine[j];
  }
  // exclusive scan of cnt over the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < CP_THREADS / 32; ++w) {
    const int t = warp_tot[w];
    if (w < warp) woff += t;
    total += t;
  }
  before = woff + incl - cnt;
  return total;
}

__global__ void __launch_bounds__(CP_THREADS) mask_count_kernel(const uint8_t* __restrict__ mask, long long n,
                                                                uint32_t* __restrict__ block_sums) {
  __shared__ int warp_tot[CP_THREADS / 32];
  int mine[CP_TILE / CP_THREADS], before;
  const int total = block_count(mask, n, (long long)blockIdx.x * CP_TILE, warp_tot, mine, before);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = (uint32_t)total;
}

// single block: exclusive scan of the block sums in place, total -> count[0]
__global__ void __launch_bounds__(1024) mask_scan_kernel(uint32_t* __restrict__ block_sums, int nb,
                                                         uint32_t* __restrict__ count) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nb ? block_sums[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      const uint32_t t = warp_tot[w];
      if (w < warp) woff += t;
      total += t;
    }
    const uint32_t carry = carry_s;
    if (i < nb) block_sums[i] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) count[0] = carry_s;
}

__global__ void __launch_bounds__(CP_THREADS) mask_write_kernel(const uint8_t* __restrict__ mask, long long n,
                                                                const uint32_t* __restrict__ block_offs,
                                                                int64_t* __restrict__ index) {
  __shared__ int warp_tot[CP_THREADS / 32];
  constexpr int PER = CP_TILE / CP_THREADS;
  int mine[PER], before;
  const long long base = (long long)blockIdx.x * CP_TILE;
  block_count(mask, n, base, warp_tot, mine, before);
  long long o = (long long)block_offs[blockIdx.x] + before;
#pragma unroll
  for (int j = 0; j < PER; ++j)
    if (mine[j]) index[o++] = base + (long long)threadIdx.x * PER + j;
}

// dst[t][(dst_row0 + r) * w + c] = src[t] ? src[t][index ? index[r] : r][c] : 0     for r < n_rows
__global__ void __launch_bounds__(CP_THREADS) gather_rows_kernel(const GatherBatch b, long long n_rows,
                                                                 const int64_t* __restrict__ index,
                                                                 long long dst_row0) {
  const int t = blockIdx.y;
  const int w = b.width[t];
  const float* __restrict__ src = b.src[t];
  float* __restrict__ dst = b.dst[t] + dst_row0 * w;
  const long long total = n_rows * w;
  for (long long e = (long long)blockIdx.x * CP_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * CP_THREADS) {
    float v = 0.0f;
    if (src) {
      const long long r = e / w;
      const int c = (int)(e - r * w);
      const long long sr = index ? index[r] : r;
      v = src[sr * w + c];
    }
    dst[e] = v;
  }
}

size_t mask_index_tmp_bytes(long long n) {
  const long long nb = (n + CP_TILE - 1) / CP_TILE;
  return (size_t)(nb > 0 ? nb : 1) * sizeof(uint32_t);
}

int launch_mask_to_index(long long n, const uint8_t* mask, int64_t* index, uint32_t* count, void* tmp, cudaStream_t st) {
  if (n == 0) {
    GSB_CUDA(cudaMemsetAsync(count, 0, sizeof(uint32_t), st));
    return GSB_OK;
  }
  const long long nb = (n + CP_TILE - 1) / CP_TILE;
  if (nb > 0x7fffffffLL) return GSB_E_INVALID;
  uint32_t* block_sums = static_cast<uint32_t*>(tmp);
  mask_count_kernel<<<(int)nb, CP_THREADS, 0, st>>>(mask, n, block_sums);
  GSB_POST_LAUNCH(false, st, "mask_count_kernel");
  mask_scan_kernel<<<1, 1024, 0, st>>>(block_sums, (int)nb, count);
  GSB_POST_LAUNCH(false, st, "mask_scan_kernel");
  mask_write_kernel<<<(int)nb, CP_THREADS, 0, st>>>(mask, n, block_sums, index);
  GSB_POST_LAUNCH(false, st, "mask_write_kernel");
  return GSB_OK;
}

int launch_gather_rows(int n_tensors, const float* const* src, float* const* dst, const int* widths, long long n_rows,
                       const int64_t* index, long long dst_row0, cudaStream_t st) {
  if (n_tensors <= 0 || n_rows <= 0) return GSB_OK;
  if (n_tensors > GSB_GATHER_MAX_TENSORS) return GSB_E_INVALID;
  GatherBatch b;
  int wmax = 0;
  for (int t = 0; t < GSB_GATHER_MAX_TENSORS; ++t) {
    const bool used = t < n_tensors;
    b.src[t] = used ? src[t] : nullptr;
    b.dst[t] = used ? dst[t] : nullptr;
    b.width[t] = used ? widths[t] : 1;
    if (used) {
      if (widths[t] <= 0 || !dst[t]) return GSB_E_INVALID;
      wmax = widths[t] > wmax ? widths[t] : wmax;
    }
  }
  const long long most = n_rows * wmax;
  long long blocks = (most + CP_THREADS * 4 - 1) / (CP_THREADS * 4);   // ~4 elements per thread of the widest tensor
  const long long cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)n_tensors);
  gather_rows_kernel<<<grid, CP_THREADS, 0, st>>>(b, n_rows, index, dst_row0);
  GSB_POST_LAUNCH(false, st, "gather_rows_kernel");
  return GSB_OK;
}

}  // namespace gsb
