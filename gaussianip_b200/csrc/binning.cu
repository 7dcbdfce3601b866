// Stage 2: binning.  Replaces InclusiveSum + duplicateWithKeys + cub::DeviceRadixSort +
// identifyTileRanges of the external operator ("Binning") with
// hand-written kernels: a stable onesweep LSD radix sort (8-bit digits; global histograms
// once, one kernel per pass with decoupled look-back, match-based stable in-block ranking), a
// gathered scan + instance emission, and a tile-boundary pass.
//
// Two modes produce the IDENTICAL instance order (tile, then depth bits, then Gaussian id):
//   GSB_BIN_TWO_LEVEL  sort the P Gaussians once by 32-bit depth key (4 passes over P), emit
//                      instances in that order, then stably partition the D instances by tile
//                      id (ceil(log2 T / 8) passes over D) — ~2.6x less traffic at D = 2P;
//   GSB_BIN_FLAT64     emit (tile<<32|depth) keys in Gaussian order and sort all D 64-bit
//                      keys (the reference's structure; kept for cross-checks).
// Everything that depends on D reads it from device memory (counts[CNT_D]); grids are sized
// by the capacity D_cap, so the host never synchronises to learn D.
#include <atomic>

#include "gsb_common.cuh"

namespace gsb {

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef GSB_RADIX_IPT
#define GSB_RADIX_IPT 8
#endif
constexpr int RS_IPT = GSB_RADIX_IPT;
constexpr int RS_TILE = RS_THREADS * RS_IPT;  // keys per block (2048 at 8 per thread)

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ long long load_n(const uint32_t* d_n, long long n_cap) {
  if (!d_n) return n_cap;
  long long n = (long long)*d_n;
  return n < n_cap ? n : n_cap;
}

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, int shift) {
  return (uint32_t)(k >> shift) & 0xFFu;
}

// Lanes of the warp holding the same 8-bit digit (and valid).  Eight ballots, fixed cost:
// match.any.sync iterates once per DISTINCT value, ~30x slower on random digits (measured r1:
// 27 us per 1 M-key pass with match.any).
__device__ __forceinline__ uint32_t match_digit(uint32_t d, bool valid) {
  uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

// exclusive scan of one value per thread over a 256-thread block; returns exclusive prefix,
// *total receives the block sum.  s_warp must hold 8 uint32.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t wsum = (lane < RS_WARPS) ? s_warp[lane] : 0;
#pragma unroll
  for (int o = 1; o < RS_WARPS; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, wsum, o);
    if (lane >= o) wsum += t;
  }
  const uint32_t wprefix = __shfl_sync(0xffffffffu, wsum, warp) - s_warp[warp];
  if (total) *total = __shfl_sync(0xffffffffu, wsum, RS_WARPS - 1);
  __syncthreads();
  return wprefix + inc - v;
}

// ---- radix sort -----------------------------------------------------------------------------
// "Onesweep" organisation: ONE pass over the keys builds the global digit histograms of every
// pass (they do not depend on the arrangement), then each pass is a single kernel that reads
// its keys once and writes them once.  A block learns how many keys of each digit precede it
// from the blocks before it by decoupled look-back over one 32-bit status word per (block,
// digit): [31:30] = flag (0 not ready, 1 block aggregate, 2 inclusive prefix), [29:0] = count.
// Block ids are handed out by an atomic counter in arrival order, so a block only ever waits
// for blocks that are already running.  Stability: the in-block rank follows the original
// order (warps own consecutive segments, rounds and lanes ascend), blocks are ordered by id.

constexpr uint32_t ST_AGG = 1u << 30, ST_PREFIX = 2u << 30, ST_MASK = (1u << 30) - 1;
constexpr int MAX_PASSES = 8;
// All blocks of a 1-2 M key pass are co-resident (one wave), so the inclusive prefix travels
// down the chain of blocks one look-back round trip at a time: read LB predecessors per trip.
#ifndef GSB_RADIX_LB
#define GSB_RADIX_LB 4
#endif
constexpr int LB = GSB_RADIX_LB;

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
radix_global_hist_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ d_n, long long n_cap,
                         int passes, uint32_t* __restrict__ ghist /*[passes][256]*/) {
  __shared__ uint32_t s_h[MAX_PASSES][256];
  const long long n = load_n(d_n, n_cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long block_base = (long long)blockIdx.x * RS_TILE;
  if (block_base >= n) return;
  for (int p = 0; p < passes; ++p) s_h[p][tid] = 0;
  __syncthreads();
  const long long seg = block_base + (long long)warp * (RS_IPT * 32);
  KeyT kk[RS_IPT];
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {          // all loads in flight before any is consumed
    const long long idx = seg + r * 32 + lane;
    kk[r] = idx < n ? keys[idx] : (KeyT)0;
  }
#pragma unroll 4
  for (int r = 0; r < RS_IPT; ++r) {
    const long long idx = seg + r * 32 + lane;
    const bool valid = idx < n;
    const KeyT k = kk[r];
    for (int p = 0; p < passes; ++p) {
      const uint32_t d = digit_of(k, 8 * p);
      const uint32_t peers = match_digit(d, valid);
      if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_h[p][d], (uint32_t)__popc(peers));
    }
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const uint32_t c = s_h[p][tid];
    if (c) atomicAdd(&ghist[p * 256 + tid], c);
  }
}

template <typename KeyT>
constexpr size_t onesweep_smem_bytes() {
  return (size_t)RS_TILE * (sizeof(KeyT) + sizeof(uint32_t)) + (size_t)(RS_WARPS * 256 + 256 + RS_WARPS + 4 + 256) * sizeof(uint32_t);
}

template <typename KeyT, bool IOTA>
__global__ void __launch_bounds__(RS_THREADS)
radix_onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                      KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                      const uint32_t* __restrict__ d_n, long long n_cap, int shift,
                      const uint32_t* __restrict__ ghist /*[256] of this pass*/,
                      uint32_t* __restrict__ ghist_next /*[256] of the next pass, or null*/,
                      volatile uint32_t* __restrict__ status /*[num_blocks][256] of this pass*/,
                      uint32_t* __restrict__ block_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KeyT* s_keys = reinterpret_cast<KeyT*>(smem_raw);                         // block-sorted keys
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + RS_TILE);
  uint32_t (*s_cnt)[256] = reinterpret_cast<uint32_t (*)[256]>(s_vals + RS_TILE);
  uint32_t* s_goff = reinterpret_cast<uint32_t*>(s_cnt + RS_WARPS);         // global slot - local slot, per digit
  uint32_t* s_warp = s_goff + 256;
  uint32_t* s_bid = s_warp + RS_WARPS;
  uint32_t* s_hn = s_bid + 4;                                               // next pass's digit counts

  const long long n = load_n(d_n, n_cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *s_bid = atomicAdd(block_counter, 1u);
  s_hn[tid] = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) s_cnt[w][tid] = 0;
  // exclusive scan of the global histogram = first output slot of every digit
  const uint32_t dbase = block_exclusive_scan_256(ghist[tid], s_warp, nullptr);  // syncs inside
  const uint32_t bid = *s_bid;
  const long long block_base = (long long)bid * RS_TILE;
  if (block_base >= n) return;
  const int count = (int)((n - block_base) < (long long)RS_TILE ? (n - block_base) : (long long)RS_TILE);

  KeyT key[RS_IPT];
  uint32_t val[RS_IPT];
  uint32_t rank[RS_IPT];
  const long long seg = block_base + (long long)warp * (RS_IPT * 32);
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {          // all loads in flight before the ranking rounds
    const long long idx = seg + r * 32 + lane;
    const bool valid = idx < n;
    key[r] = valid ? keys_in[idx] : (KeyT)0;
    val[r] = valid ? (IOTA ? (uint32_t)idx : vals_in[idx]) : 0u;
  }
  if (ghist_next) {
    // the histogram of the NEXT digit does not depend on the arrangement: count it here, while
    // the keys are in registers, instead of in a separate pass over the keys
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r)
      if (seg + r * 32 + lane < n) atomicAdd(&s_hn[digit_of(key[r], shift + 8)], 1u);
  }
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    const long long idx = seg + r * 32 + lane;
    const bool valid = idx < n;
    const uint32_t d = digit_of(key[r], shift);
    const uint32_t peers = match_digit(d, valid);
    const int leader = valid ? __ffs(peers) - 1 : lane;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = s_cnt[warp][d];
      s_cnt[warp][d] = old + (uint32_t)__popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + (uint32_t)__popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  {
    // thread tid owns digit tid: exclusive offsets of the warps inside the block, block count
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = s_cnt[w][tid];
      s_cnt[w][tid] = run;
      run += c;
    }
    volatile uint32_t* mine = status + (size_t)bid * 256 + tid;
    uint32_t excl = 0;
    if (bid == 0) {
      *mine = ST_PREFIX | run;
    } else {
      *mine = ST_AGG | run;
      long long j = (long long)bid - 1;
      bool found = false;
      while (!found) {                           // LB predecessors per L2 round trip
        uint32_t v[LB];
#pragma unroll
        for (int k = 0; k < LB; ++k) {
          v[k] = 2u << 30;                       // virtual block -1: inclusive prefix 0
          if (j - k >= 0) v[k] = status[(size_t)(j - k) * 256 + tid];
        }
        int used = 0;
#pragma unroll
        for (int k = 0; k < LB; ++k) {
          if (found || used != k) continue;
          const uint32_t flag = v[k] & ~ST_MASK;
          if (flag == 0) continue;               // running but not published yet: retry from here
          excl += v[k] & ST_MASK;
          used = k + 1;
          if (flag == ST_PREFIX) found = true;
        }
        j -= used;
      }
      *mine = ST_PREFIX | (excl + run);
    }
    // position of this digit's run inside the block-sorted tile
    const uint32_t local = block_exclusive_scan_256(run, s_warp, nullptr);   // syncs inside
    s_goff[tid] = dbase + excl - local;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) s_cnt[w][tid] += local;
  }
  __syncthreads();
  // stage the tile in shared memory in sorted order, then write runs of equal digits with
  // consecutive threads -> consecutive addresses (coalesced instead of one sector per key)
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    const long long idx = seg + r * 32 + lane;
    if (idx < n) {
      const uint32_t lpos = s_cnt[warp][digit_of(key[r], shift)] + rank[r];
      s_keys[lpos] = key[r];
      s_vals[lpos] = val[r];
    }
  }
  __syncthreads();
  for (int i = tid; i < count; i += RS_THREADS) {
    const KeyT k = s_keys[i];
    const uint32_t pos = s_goff[digit_of(k, shift)] + (uint32_t)i;
    keys_out[pos] = k;
    vals_out[pos] = s_vals[i];
  }
  if (ghist_next && s_hn[tid]) atomicAdd(&ghist_next[tid], s_hn[tid]);   // ordered by the barrier above
}

// ---- scan of tiles_touched (gathered through an order) + instance emission -------------------

constexpr int SC_IPT = 8;
constexpr int SC_TILE = RS_THREADS * SC_IPT;  // 2048 Gaussians per block

__global__ void __launch_bounds__(RS_THREADS)
tiles_partial_kernel(const uint32_t* __restrict__ tiles, const uint32_t* __restrict__ order, int P,
                     uint32_t* __restrict__ blocksums) {
  __shared__ uint32_t s_warp[RS_WARPS];
  const int base = blockIdx.x * SC_TILE;
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    const int j = base + k * RS_THREADS + threadIdx.x;
    if (j < P) sum += tiles[order ? order[j] : j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < RS_WARPS; ++w) t += s_warp[w];
    blocksums[blockIdx.x] = t;
  }
}

// single block: exclusive scan of the per-block sums; publishes D and the overflow flag
__global__ void __launch_bounds__(RS_THREADS)
scan_blocksums_kernel(uint32_t* __restrict__ blocksums, int num_blocks, uint32_t* __restrict__ counts,
                      long long D_cap, int mode, const uint32_t* __restrict__ flag_word) {
  __shared__ uint32_t s_warp[RS_WARPS];
  unsigned long long carry = 0;
  for (int base = 0; base < num_blocks; base += RS_THREADS) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < num_blocks ? blocksums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan_256(v, s_warp, &total);
    if (i < num_blocks) blocksums[i] = (uint32_t)carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    const unsigned long long D = carry;
    counts[CNT_D] = D > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)D;
    counts[CNT_OVERFLOW] = D > (unsigned long long)D_cap ? 1u : 0u;
    counts[CNT_VISIBLE] = 0; counts[CNT_MAXTILES] = 0;   // reserved; the whole block is copied to the host
    counts[4] = (uint32_t)mode;
    counts[CNT_PREFILTER] = *flag_word;     // points culled although the caller declared them prefiltered
    counts[6] = 0; counts[7] = 0;
  }
}

// MODE 0: write (tile id, Gaussian id) ; MODE 1: write (tile<<32|depth, Gaussian id)
template <int MODE>
__global__ void __launch_bounds__(RS_THREADS)
emit_kernel(const uint32_t* __restrict__ tiles, const uint32_t* __restrict__ order, int P,
            const uint32_t* __restrict__ blocksums, const ushort4* __restrict__ rect,
            const uint32_t* __restrict__ dkeys, int gx, long long D_cap, uint32_t* __restrict__ tkeys,
            uint64_t* __restrict__ keys64, uint32_t* __restrict__ vals, uint32_t* __restrict__ ghist0) {
  __shared__ uint32_t s_warp[RS_WARPS];
  __shared__ uint32_t s_h0[256];     // digit-0 histogram of the emitted keys (first pass of the sort)
  s_h0[threadIdx.x] = 0;
  __syncthreads();
  // blocked arrangement: thread t owns SC_IPT consecutive Gaussians of the emission order
  const int first = blockIdx.x * SC_TILE + threadIdx.x * SC_IPT;
  uint32_t gid[SC_IPT], cnt[SC_IPT];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    const int j = first + k;
    gid[k] = 0; cnt[k] = 0;
    if (j < P) {
      gid[k] = order ? order[j] : (uint32_t)j;
      cnt[k] = tiles[gid[k]];
    }
    mine += cnt[k];
  }
  uint32_t off = blocksums[blockIdx.x] + block_exclusive_scan_256(mine, s_warp, nullptr);
  constexpr uint32_t BIG = 64;   // splats touching more tiles than this are emitted by the whole warp
  uint32_t offk[SC_IPT];
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) { offk[k] = off; off += cnt[k]; }
  auto put = [&](long long o, uint32_t tile, uint32_t dk, uint32_t g) {
    if (o < D_cap) {
      if (MODE == 0) tkeys[o] = tile;
      else keys64[o] = ((uint64_t)tile << 32) | dk;
      vals[o] = g;
      atomicAdd(&s_h0[(MODE == 0 ? tile : dk) & 0xFFu], 1u);
    }
  };
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    if (cnt[k] == 0 || cnt[k] > BIG) continue;
    const ushort4 rc = rect[gid[k]];
    const uint32_t dk = (MODE == 1) ? dkeys[gid[k]] : 0u;
    long long o = offk[k];
    for (int y = rc.y; y < rc.w; ++y)
      for (int x = rc.x; x < rc.z; ++x) put(o++, (uint32_t)(y * gx + x), dk, gid[k]);
  }
  // large splats (a Gaussian in front of a zoomed-in camera can touch every tile): one lane looping
  // over thousands of tiles would serialise the warp, so the 32 lanes share each of them
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    uint32_t big = __ballot_sync(0xffffffffu, cnt[k] > BIG);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const uint32_t g = __shfl_sync(0xffffffffu, gid[k], src);
      const uint32_t c = __shfl_sync(0xffffffffu, cnt[k], src);
      const uint32_t o0 = __shfl_sync(0xffffffffu, offk[k], src);
      const ushort4 rc = rect[g];
      const uint32_t dk = (MODE == 1) ? dkeys[g] : 0u;
      const uint32_t wx = (uint32_t)(rc.z - rc.x);
      for (uint32_t i = lane; i < c; i += 32) {
        const uint32_t y = rc.y + i / wx, x = rc.x + i % wx;
        put((long long)o0 + i, y * (uint32_t)gx + x, dk, g);
      }
    }
  }
  __syncthreads();
  if (s_h0[threadIdx.x]) atomicAdd(&ghist0[threadIdx.x], s_h0[threadIdx.x]);
}

template <typename KeyT>
__global__ void tile_ranges_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ d_n,
                                   long long n_cap, uint2* __restrict__ ranges) {
  const long long n = load_n(d_n, n_cap);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int sh = sizeof(KeyT) == 8 ? 32 : 0;
  const uint32_t t = (uint32_t)(keys[i] >> sh);
  if (i == 0) ranges[t].x = 0;
  else {
    const uint32_t p = (uint32_t)(keys[i - 1] >> sh);
    if (p != t) { ranges[p].y = (uint32_t)i; ranges[t].x = (uint32_t)i; }
  }
  if (i == n - 1) ranges[t].y = (uint32_t)n;
}

// Heaviest-first CTA order for the blend kernels: tiles bucketed by floor(log2(list length)),
// longest bucket first (order inside a bucket is irrelevant: it only affects scheduling).
__global__ void __launch_bounds__(1024)
tile_order_kernel(const uint2* __restrict__ ranges, int T, uint32_t* __restrict__ order) {
  __shared__ uint32_t s_cnt[33], s_cur[33];
  if (threadIdx.x < 33) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const uint2 r = ranges[t];
    atomicAdd(&s_cnt[32 - __clz(r.y - r.x)], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int b = 32; b >= 0; --b) { s_cur[b] = run; run += s_cnt[b]; }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const uint2 r = ranges[t];
    order[atomicAdd(&s_cur[32 - __clz(r.y - r.x)], 1u)] = (uint32_t)t;
  }
}

__global__ void compose_keys_kernel(const uint32_t* __restrict__ tkeys, const uint32_t* __restrict__ point_list,
                                    const uint32_t* __restrict__ dkeys, const uint32_t* __restrict__ d_n,
                                    long long n_cap, uint64_t* __restrict__ out) {
  const long long n = load_n(d_n, n_cap);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = ((uint64_t)tkeys[i] << 32) | dkeys[point_list[i]];
}

inline int bits_for(unsigned int max_value) {  // bits needed to represent values in [0, max_value]
  int b = 0;
  while (max_value) { ++b; max_value >>= 1; }
  return b;
}

}  // namespace

int radix_num_passes(int end_bit) { return (end_bit + 7) / 8; }
bool radix_result_in_A(int passes) { return (passes & 1) != 0; }

// tmp layout: [MAX_PASSES][256] global histograms | [MAX_PASSES] block counters (padded to 256) |
//             [MAX_PASSES][num_blocks][256] look-back status words
size_t radix_tmp_bytes(long long n_cap) {
  const long long nb = (n_cap + RS_TILE - 1) / RS_TILE;
  return (size_t)(MAX_PASSES * 256 + 256 + (size_t)MAX_PASSES * (nb > 0 ? nb : 1) * 256) * sizeof(uint32_t);
}

static size_t radix_used_bytes(long long n_cap, int passes) {
  const long long nb = (n_cap + RS_TILE - 1) / RS_TILE;
  return (size_t)(MAX_PASSES * 256 + 256 + (size_t)passes * (nb > 0 ? nb : 1) * 256) * sizeof(uint32_t);
}

// Zero the histograms / counters / look-back words of one sort.  Must run BEFORE the kernel that
// produces the digit-0 histogram (preprocess for the depth keys, emit for the tile keys).
int radix_prepare(long long n_cap, int end_bit, void* tmp, cudaStream_t st) {
  if (n_cap <= 0) return GSB_OK;
  GSB_CUDA(cudaMemsetAsync(tmp, 0, radix_used_bytes(n_cap, radix_num_passes(end_bit)), st));
  return GSB_OK;
}

uint32_t* radix_hist0(void* tmp) { return static_cast<uint32_t*>(tmp); }
// A spare word of the (zeroed) counter row: preprocess_fwd raises it when GsbSettings.prefiltered is set and a point
// fails the near-plane test; scan_blocksums publishes it as counts[5].
uint32_t* radix_flag_word(void* tmp) { return static_cast<uint32_t*>(tmp) + MAX_PASSES * 256 + 255; }

// Stable LSD sort on bits [0,end_bit).  Pass 0 reads (src_keys, src_vals); pass p writes
// buffer A when p is even and buffer B when p is odd.  src may alias B (never A).  The result
// is in A when the number of passes is odd, in B when it is even (see radix_result_in_A).
// hist0_ready: the caller ran radix_prepare and the producer of the keys already accumulated the
// digit-0 histogram into radix_hist0(tmp); otherwise both happen here.  Every pass counts the next
// pass's digits while it holds the keys, so there is no separate histogram pass.
template <typename KeyT>
int radix_sort_pairs(long long n_cap, const uint32_t* d_n, const KeyT* src_keys, const uint32_t* src_vals,
                     KeyT* keysA, uint32_t* valsA, KeyT* keysB, uint32_t* valsB, int end_bit,
                     bool iota_vals, bool hist0_ready, void* tmp, bool debug, cudaStream_t st) {
  if (n_cap <= 0) return GSB_OK;
  const int passes = radix_num_passes(end_bit);
  if (passes == 0) return GSB_OK;
  if (passes > MAX_PASSES) return GSB_E_INVALID;
  const int nb = (int)((n_cap + RS_TILE - 1) / RS_TILE);
  uint32_t* ghist = static_cast<uint32_t*>(tmp);
  uint32_t* counters = ghist + MAX_PASSES * 256;
  uint32_t* status = counters + 256;
  constexpr size_t smem = onesweep_smem_bytes<KeyT>();
  {
    static std::atomic<unsigned long long> configured{0};   // bit per device (the attribute is per device and instantiation)
    int dev = 0;
    GSB_CUDA(cudaGetDevice(&dev));
    if (!(configured.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
      GSB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<KeyT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      GSB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<KeyT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  if (!hist0_ready) {
    int rc = radix_prepare(n_cap, end_bit, tmp, st);
    if (rc) return rc;
    radix_global_hist_kernel<KeyT><<<nb, RS_THREADS, 0, st>>>(src_keys, d_n, n_cap, 1, ghist);
    GSB_POST_LAUNCH(debug, st, "radix_global_hist_kernel");
  }
  for (int p = 0; p < passes; ++p) {
    const KeyT* kin = p == 0 ? src_keys : ((p & 1) ? keysA : keysB);
    const uint32_t* vin = p == 0 ? src_vals : ((p & 1) ? valsA : valsB);
    KeyT* kout = (p & 1) ? keysB : keysA;
    uint32_t* vout = (p & 1) ? valsB : valsA;
    uint32_t* stp = status + (size_t)p * nb * 256;
    uint32_t* gnext = p + 1 < passes ? ghist + (p + 1) * 256 : nullptr;
    if (p == 0 && iota_vals)
      radix_onesweep_kernel<KeyT, true><<<nb, RS_THREADS, smem, st>>>(kin, vin, kout, vout, d_n, n_cap, 8 * p,
                                                                      ghist + p * 256, gnext, stp, counters + p);
    else
      radix_onesweep_kernel<KeyT, false><<<nb, RS_THREADS, smem, st>>>(kin, vin, kout, vout, d_n, n_cap, 8 * p,
                                                                       ghist + p * 256, gnext, stp, counters + p);
    GSB_POST_LAUNCH(debug, st, "radix_onesweep_kernel");
  }
  return GSB_OK;
}

template int radix_sort_pairs<uint32_t>(long long, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*,
                                        uint32_t*, uint32_t*, uint32_t*, int, bool, bool, void*, bool, cudaStream_t);
template int radix_sort_pairs<uint64_t>(long long, const uint32_t*, const uint64_t*, const uint32_t*, uint64_t*,
                                        uint32_t*, uint64_t*, uint32_t*, int, bool, bool, void*, bool, cudaStream_t);

int launch_bin_sort(const View& v, int P, void* saved, void* scratch, const GsbLayout& L,
                    long long D_cap, int mode, uint32_t* host_counts, cudaEvent_t event, bool debug,
                    cudaStream_t st) {
  uint32_t* counts = at<uint32_t>(saved, L.off_counts);
  uint32_t* point_list = at<uint32_t>(saved, L.off_point_list);
  uint2* ranges = at<uint2>(saved, L.off_ranges);
  const ushort4* rect = at<ushort4>(scratch, L.off_rect);
  const uint32_t* tiles = at<uint32_t>(scratch, L.off_tiles);
  const uint32_t* dkeys = at<uint32_t>(scratch, L.off_dkeys0);   // written by preprocess, kept intact
  uint32_t* blocksums = at<uint32_t>(scratch, L.off_blocksums);
  void* hist = at<char>(scratch, L.off_hist);
  const int T = v.gx * v.gy;
  GSB_CUDA(cudaMemsetAsync(ranges, 0, (size_t)T * sizeof(uint2), st));
  auto publish = [&]() -> int {
    if (host_counts)
      GSB_CUDA(cudaMemcpyAsync(host_counts, counts, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (event) GSB_CUDA(cudaEventRecord(event, st));
    return GSB_OK;
  };
  if (P == 0) {
    GSB_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(uint32_t), st));
    tile_order_kernel<<<1, 1024, 0, st>>>(ranges, T, at<uint32_t>(saved, L.off_tile_order));
    GSB_POST_LAUNCH(debug, st, "tile_order_kernel");
    return publish();
  }
  const int sc_blocks = (P + SC_TILE - 1) / SC_TILE;
  const int tile_bits = bits_for((unsigned)(T - 1));
  const int cap_blocks = (int)((D_cap + 255) / 256);
  uint32_t* alt_vals = at<uint32_t>(scratch, L.off_tvals_alt);
  int rc;

  if (mode == GSB_BIN_TWO_LEVEL) {
    // (1) Gaussians by depth key (ties by index: the sort is stable and values start as iota)
    uint32_t* kA = at<uint32_t>(scratch, L.off_dkeys1);
    uint32_t* vA = at<uint32_t>(scratch, L.off_didx0);
    uint32_t* kB = at<uint32_t>(scratch, L.off_dkeys2);
    uint32_t* vB = at<uint32_t>(scratch, L.off_didx1);
    prof_begin(GSB_STAGE_DEPTH_SORT, st);
    // digit-0 histogram of the depth keys was accumulated by preprocess_fwd (after radix_prepare)
    rc = radix_sort_pairs<uint32_t>(P, nullptr, dkeys, nullptr, kA, vA, kB, vB, 32, true, true, hist, debug, st);
    prof_end(GSB_STAGE_DEPTH_SORT, st);
    if (rc) return rc;
    prof_begin(GSB_STAGE_SCAN_EMIT, st);
    const uint32_t* order = vB;  // 4 passes -> result in B
    // (2) offsets in that order, D, overflow flag
    tiles_partial_kernel<<<sc_blocks, RS_THREADS, 0, st>>>(tiles, order, P, blocksums);
    GSB_POST_LAUNCH(debug, st, "tiles_partial_kernel");
    scan_blocksums_kernel<<<1, RS_THREADS, 0, st>>>(blocksums, sc_blocks, counts, D_cap, mode, radix_flag_word(hist));
    GSB_POST_LAUNCH(debug, st, "scan_blocksums_kernel");
    if ((rc = publish())) return rc;
    // (3) emit (tile id, Gaussian id) in (depth, id, tile) order, then stable partition by tile
    const bool inA = radix_result_in_A(radix_num_passes(tile_bits));
    uint32_t* tkA = at<uint32_t>(scratch, L.off_tkeys0);
    uint32_t* tkB = at<uint32_t>(scratch, L.off_tkeys1);
    uint32_t* tvA = inA ? point_list : alt_vals;
    uint32_t* tvB = inA ? alt_vals : point_list;
    if ((rc = radix_prepare(D_cap, tile_bits, hist, st))) return rc;
    emit_kernel<0><<<sc_blocks, RS_THREADS, 0, st>>>(tiles, order, P, blocksums, rect, dkeys, v.gx, D_cap, tkB,
                                                     nullptr, tvB, radix_hist0(hist));
    GSB_POST_LAUNCH(debug, st, "emit_kernel");
    prof_end(GSB_STAGE_SCAN_EMIT, st);
    prof_begin(GSB_STAGE_TILE_SORT, st);
    rc = radix_sort_pairs<uint32_t>(D_cap, counts + CNT_D, tkB, tvB, tkA, tvA, tkB, tvB, tile_bits, false, true,
                                    hist, debug, st);
    prof_end(GSB_STAGE_TILE_SORT, st);
    if (rc) return rc;
    ProfScope pr(GSB_STAGE_RANGES, st);
    // (a single-tile image needs 0 passes: the emitted order in B is already final)
    tile_ranges_kernel<uint32_t><<<cap_blocks, 256, 0, st>>>(inA ? tkA : tkB, counts + CNT_D, D_cap, ranges);
    GSB_POST_LAUNCH(debug, st, "tile_ranges_kernel");
    tile_order_kernel<<<1, 1024, 0, st>>>(ranges, T, at<uint32_t>(saved, L.off_tile_order));
    GSB_POST_LAUNCH(debug, st, "tile_order_kernel");
    return GSB_OK;
  }
  if (mode == GSB_BIN_FLAT64) {
    prof_begin(GSB_STAGE_SCAN_EMIT, st);
    tiles_partial_kernel<<<sc_blocks, RS_THREADS, 0, st>>>(tiles, nullptr, P, blocksums);
    GSB_POST_LAUNCH(debug, st, "tiles_partial_kernel");
    scan_blocksums_kernel<<<1, RS_THREADS, 0, st>>>(blocksums, sc_blocks, counts, D_cap, mode, radix_flag_word(hist));
    GSB_POST_LAUNCH(debug, st, "scan_blocksums_kernel");
    if ((rc = publish())) return rc;
    const int end_bit = 32 + tile_bits;
    const bool inA = radix_result_in_A(radix_num_passes(end_bit));
    uint64_t* kA = at<uint64_t>(scratch, L.off_keys64_0);
    uint64_t* kB = at<uint64_t>(scratch, L.off_keys64_1);
    uint32_t* tvA = inA ? point_list : alt_vals;
    uint32_t* tvB = inA ? alt_vals : point_list;
    if ((rc = radix_prepare(D_cap, end_bit, hist, st))) return rc;
    emit_kernel<1><<<sc_blocks, RS_THREADS, 0, st>>>(tiles, nullptr, P, blocksums, rect, dkeys, v.gx, D_cap,
                                                     nullptr, kB, tvB, radix_hist0(hist));
    GSB_POST_LAUNCH(debug, st, "emit_kernel");
    prof_end(GSB_STAGE_SCAN_EMIT, st);
    prof_begin(GSB_STAGE_TILE_SORT, st);
    rc = radix_sort_pairs<uint64_t>(D_cap, counts + CNT_D, kB, tvB, kA, tvA, kB, tvB, end_bit, false, true, hist,
                                    debug, st);
    prof_end(GSB_STAGE_TILE_SORT, st);
    if (rc) return rc;
    ProfScope pr(GSB_STAGE_RANGES, st);
    tile_ranges_kernel<uint64_t><<<cap_blocks, 256, 0, st>>>(inA ? kA : kB, counts + CNT_D, D_cap, ranges);
    GSB_POST_LAUNCH(debug, st, "tile_ranges_kernel");
    tile_order_kernel<<<1, 1024, 0, st>>>(ranges, T, at<uint32_t>(saved, L.off_tile_order));
    GSB_POST_LAUNCH(debug, st, "tile_order_kernel");
    return GSB_OK;
  }
  return GSB_E_INVALID;
}

int launch_debug_sorted_keys(const View& v, int P, const void* saved, const void* scratch, const GsbLayout& L,
                             long long D_cap, uint64_t* keys_out, cudaStream_t st) {
  (void)P;
  const uint32_t* counts = at<uint32_t>(saved, L.off_counts);
  uint32_t h_counts[8];
  GSB_CUDA(cudaMemcpyAsync(h_counts, counts, sizeof(h_counts), cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaStreamSynchronize(st));
  const int mode = (int)h_counts[4];
  const long long n = h_counts[CNT_D] < (unsigned long long)D_cap ? h_counts[CNT_D] : D_cap;
  if (n == 0) return GSB_OK;
  const int T = v.gx * v.gy;
  const int tile_bits = bits_for((unsigned)(T - 1));
  if (mode == GSB_BIN_FLAT64) {
    const bool inA = radix_result_in_A(radix_num_passes(32 + tile_bits));
    const uint64_t* src = at<uint64_t>(scratch, inA ? L.off_keys64_0 : L.off_keys64_1);
    GSB_CUDA(cudaMemcpyAsync(keys_out, src, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    return GSB_OK;
  }
  const bool inA = radix_result_in_A(radix_num_passes(tile_bits));
  const uint32_t* tk = at<uint32_t>(scratch, inA ? L.off_tkeys0 : L.off_tkeys1);
  compose_keys_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(tk, at<uint32_t>(saved, L.off_point_list),
                                                              at<uint32_t>(scratch, L.off_dkeys0), counts + CNT_D,
                                                              D_cap, keys_out);
  GSB_POST_LAUNCH(false, st, "compose_keys_kernel");
  return GSB_OK;
}

}  // namespace gsb
