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
// "Onesweep" organisation: the global digit histogram of a pass does not depend on the arrangement of the keys,
// so it is counted by whoever holds the keys in registers before the pass (the producing kernel for digit 0, the
// previous pass for the others), and each pass is a single kernel that reads its keys once and writes them once.
// A block learns how many keys of each digit precede it from the blocks before it through one 32-bit status word
// per (block, digit): bit 30 = published, [29:0] = count.  Block ids are handed out by an atomic counter in
// arrival order, so a block only ever waits for blocks that are already running.  Stability: the in-block rank
// follows the original order (warps own consecutive segments, rounds and lanes ascend), blocks are ordered by id.
//
// TWO-LEVEL look-back.  At the sizes of this path (1-2 M keys) all blocks of a pass are co-resident and finish
// ranking at about the same time, so the classic chained look-back (read a few predecessors, stop at the first
// inclusive prefix) degenerates into a wavefront that needs ~sqrt(blocks/2) dependent L2 round trips — 8.4 trips
// per warp on average, 30 % of the executed instructions and half of the stall samples of a pass (ncu r2l).
// Here blocks form groups of LB_GROUP: a block sums the AGGREGATES of the predecessors inside its group (they are
// published right after ranking; up to LB_GROUP-1 independent loads), the last block of a group publishes the
// group's aggregate, and every block adds the aggregates of the groups before its own.  Two dependent waits, a
// bounded number of loads, nothing spins while it makes progress.

constexpr uint32_t ST_READY = 1u << 30, ST_MASK = (1u << 30) - 1;
constexpr int MAX_PASSES = 8;
constexpr int LB_GROUP = 32;        // blocks per look-back group
#ifndef GSB_RADIX_LB
#define GSB_RADIX_LB 16
#endif
constexpr int LB = GSB_RADIX_LB;    // status words requested per L2 round trip
#ifndef GSB_RADIX_POLL_NS
#define GSB_RADIX_POLL_NS 100
#endif

__host__ __device__ inline long long radix_blocks(long long n_cap) { return (n_cap + RS_TILE - 1) / RS_TILE; }
__host__ __device__ inline long long radix_groups(long long nb) { return (nb + LB_GROUP - 1) / LB_GROUP; }

// Status words are written and polled with gpu-scope relaxed accesses (flag and payload share the word, so no
// ordering between accesses is needed; `volatile` would compile to system-scope STRONG.SYS accesses).
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Sum of `count` published status words base[0], base[stride], ... (spins on the ones not published yet).
__device__ __forceinline__ uint32_t sum_published(const uint32_t* base, int count, size_t stride) {
  uint32_t sum = 0;
  for (int k0 = 0; k0 < count; k0 += LB) {
    uint32_t v[LB];
#pragma unroll
    for (int k = 0; k < LB; ++k) v[k] = (k0 + k < count) ? ld_status(base + (size_t)(k0 + k) * stride) : ST_READY;
#pragma unroll
    for (int k = 0; k < LB; ++k) {
      while (!(v[k] & ST_READY)) {         // not published yet: yield the issue slots to the blocks still ranking
        __nanosleep(GSB_RADIX_POLL_NS);
        v[k] = ld_status(base + (size_t)(k0 + k) * stride);
      }
      sum += v[k] & ST_MASK;
    }
  }
  return sum;
}

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

// Diagnostic build only (-DGSB_RADIX_TIMING): per-block clock64 stamps at the phase boundaries of a pass.
#ifdef GSB_RADIX_TIMING
__device__ long long g_radix_timing[8 * 8192];
#define RS_STAMP(k) do { if (tid == 0 && bid < 8192u) g_radix_timing[bid * 8 + (k)] = clock64(); } while (0)
#else
#define RS_STAMP(k) do { } while (0)
#endif

template <typename KeyT>
constexpr size_t onesweep_smem_bytes() {
  return (size_t)RS_TILE * (sizeof(KeyT) + sizeof(uint32_t)) + (size_t)(RS_WARPS * 256 + 256 + RS_WARPS + 4 + 256) * sizeof(uint32_t);
}

// RANGES (last pass of a tile sort): the pass that puts the instances into their final order also records where
// every tile's list starts and ends.  In the block-sorted tile equal tile ids are adjacent (the input of the last
// pass is ordered by all lower bits, and the partition by the top digit is stable), so the first / last key of
// every run of equal tiles sends its output position to ranges[tile] with one atomic: x accumulates the maximum
// of ~position (= the minimum position; the array starts zeroed), y the maximum of position + 1.
// tile_order_kernel decodes x and resets empty tiles to (0, 0).
#ifndef GSB_RADIX_MINB
#define GSB_RADIX_MINB 4
#endif
template <typename KeyT, bool IOTA, bool RANGES>
__global__ void __launch_bounds__(RS_THREADS, GSB_RADIX_MINB)
radix_onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                      KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                      const uint32_t* __restrict__ d_n, long long n_cap, int shift,
                      const uint32_t* __restrict__ ghist /*[256] of this pass*/,
                      uint32_t* __restrict__ ghist_next /*[256] of the next pass, or null*/,
                      uint32_t* __restrict__ status /*[num_blocks][256] of this pass*/,
                      uint32_t* __restrict__ gstatus /*[num_groups][256] of this pass*/,
                      uint32_t* __restrict__ block_counter, uint2* __restrict__ ranges) {
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
  RS_STAMP(0);
#ifdef GSB_RADIX_TIMING
  if (tid == 0 && bid < 8192u) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_radix_timing[bid * 8 + 7] = gt; }
#endif

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
  RS_STAMP(1);
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    const long long idx = seg + r * 32 + lane;
    const bool valid = idx < n;
    const uint32_t d = digit_of(key[r], shift);
    const uint32_t peers = match_digit(d, valid);
#ifndef GSB_RADIX_RANK_NOBRANCH
#define GSB_RADIX_RANK_NOBRANCH 1
#endif
#if GSB_RADIX_RANK_NOBRANCH
    // every lane reads its digit's counter (peers read the same word: a broadcast), the lowest peer alone updates
    // it: no divergent region and no shuffle of the old value in the hottest loop of the pass
    const uint32_t old = s_cnt[warp][d];
    __syncwarp();
    if (valid && (peers & lt) == 0u) s_cnt[warp][d] = old + (uint32_t)__popc(peers);
    rank[r] = old + (uint32_t)__popc(peers & lt);
    __syncwarp();
#else
    const int leader = valid ? __ffs(peers) - 1 : lane;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = s_cnt[warp][d];
      s_cnt[warp][d] = old + (uint32_t)__popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + (uint32_t)__popc(peers & lt);
    __syncwarp();
#endif
  }
  RS_STAMP(2);
  __syncthreads();
  RS_STAMP(3);
  // thread tid owns digit tid: exclusive offsets of the warps inside the block, block count
  uint32_t run = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    const uint32_t c = s_cnt[w][tid];
    s_cnt[w][tid] = run;
    run += c;
  }
  const uint32_t j = bid % LB_GROUP, g = bid / LB_GROUP;
  st_status(status + (size_t)bid * 256 + tid, ST_READY | run);   // this block's aggregate: nothing else is needed from it
  // position of this digit's run inside the block-sorted tile
  const uint32_t local = block_exclusive_scan_256(run, s_warp, nullptr);   // syncs inside
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) s_cnt[w][tid] += local;
  __syncthreads();
  // stage the tile in shared memory in sorted order (needs only block-local offsets), so that the wait for the
  // predecessors below overlaps with it; runs of equal digits are then written by consecutive threads ->
  // consecutive addresses (coalesced instead of one sector per key)
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    const long long idx = seg + r * 32 + lane;
    if (idx < n) {
      const uint32_t lpos = s_cnt[warp][digit_of(key[r], shift)] + rank[r];
      s_keys[lpos] = key[r];
      s_vals[lpos] = val[r];
    }
  }
  RS_STAMP(4);
  {
    // two-level look-back (see the header of this section): aggregates of the predecessors inside my group,
    // then aggregates of the groups before mine.  A digit this block holds no key of needs no offset; only the
    // last block of a group must sum every digit, because it publishes the group's aggregate.
    const bool leader = j == LB_GROUP - 1;
    uint32_t excl = 0;
    if (run != 0u || leader) {
      excl = sum_published(status + (size_t)(bid - j) * 256 + tid, (int)j, 256);
      if (leader) st_status(gstatus + (size_t)g * 256 + tid, ST_READY | (excl + run));
      if (run != 0u) excl += sum_published(gstatus + tid, (int)g, 256);
    }
    s_goff[tid] = dbase + excl - local;
  }
  __syncthreads();
  RS_STAMP(5);
  for (int i = tid; i < count; i += RS_THREADS) {
    const KeyT k = s_keys[i];
    const uint32_t pos = s_goff[digit_of(k, shift)] + (uint32_t)i;
    keys_out[pos] = k;
    vals_out[pos] = s_vals[i];
    if (RANGES) {
      constexpr int tsh = sizeof(KeyT) == 8 ? 32 : 0;
      const uint32_t t = (uint32_t)(k >> tsh);
      if (i == 0 || (uint32_t)(s_keys[i - 1] >> tsh) != t) atomicMax(&ranges[t].x, ~pos);
      if (i == count - 1 || (uint32_t)(s_keys[i + 1] >> tsh) != t) atomicMax(&ranges[t].y, pos + 1u);
    }
  }
  if (ghist_next && s_hn[tid]) atomicAdd(&ghist_next[tid], s_hn[tid]);   // ordered by the barrier above
  RS_STAMP(6);
}

// ---- scan of tiles_touched (gathered through an order) + instance emission -------------------

#ifndef GSB_EMIT_IPT
#define GSB_EMIT_IPT 4
#endif
constexpr int SC_IPT = GSB_EMIT_IPT;
constexpr int SC_TILE = RS_THREADS * SC_IPT;  // Gaussians per block
static_assert(SC_TILE >= 512, "GsbLayout.off_blocksums is sized for >= 512 Gaussians per block");

// Scan chain of the emission kernel (one 64-bit word per block and per group of LB_GROUP blocks, bit 63 =
// published; word 0 = the ticket counter, word 1 = the parked prefiltered flag).  Same two-level scheme as the radix look-back, one value per block.
constexpr unsigned long long CH_READY = 1ull << 63;
__device__ __forceinline__ unsigned long long ld_chain(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_chain(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__host__ __device__ inline size_t emit_chain_words(long long blocks) { return (size_t)(2 + blocks + radix_groups(blocks)); }

// Offsets + instance emission in ONE kernel.  A block gathers tiles_touched of its Gaussians through the
// emission order, scans them, learns the number of instances before it from the scan chain (blocks take tickets in
// arrival order; warp 0 sums the published totals of the predecessors in its group and of the earlier groups, 32
// words per round trip) and writes its instances.  The block holding the last ticket publishes D, the overflow
// flag and the rest of the counts row.  (Round 1 ran a partial-sum kernel, a one-block scan and this kernel.)
// MODE 0: write (tile id, Gaussian id) ; MODE 1: write (tile<<32|depth, Gaussian id)
#ifndef GSB_EMIT_MINB
#define GSB_EMIT_MINB 6
#endif
template <int MODE>
__global__ void __launch_bounds__(RS_THREADS, GSB_EMIT_MINB)
emit_kernel(const uint32_t* __restrict__ order, int P,
            unsigned long long* __restrict__ chain, int num_blocks, const ushort4* __restrict__ rect,
            const uint32_t* __restrict__ dkeys, int gx, long long D_cap, uint32_t* __restrict__ tkeys,
            uint64_t* __restrict__ keys64, uint32_t* __restrict__ vals, uint32_t* __restrict__ ghist0,
            uint32_t* __restrict__ counts, int mode, const uint32_t* __restrict__ flag_word) {
  __shared__ uint32_t s_warp[RS_WARPS];
  __shared__ uint32_t s_h0[256];     // digit-0 histogram of the emitted keys (first pass of the sort)
  __shared__ uint32_t s_bid, s_base;
  s_h0[threadIdx.x] = 0;
  // persistent CTAs (one wave): each takes tickets until the tiles of SC_TILE Gaussians are used up
  while (true) {
  __syncthreads();                   // s_bid / s_base of the previous round are no longer read
  if (threadIdx.x == 0) s_bid = (uint32_t)atomicAdd(chain, 1ull);
  __syncthreads();
  const int bid = (int)s_bid;
  if (bid >= num_blocks) break;
  // blocked arrangement: thread t owns SC_IPT consecutive Gaussians of the emission order
  const int first = bid * SC_TILE + threadIdx.x * SC_IPT;
  uint32_t gid[SC_IPT], cnt[SC_IPT], dk[SC_IPT];
  ushort4 rc[SC_IPT];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    const int j = first + k;
    gid[k] = 0;
    if (j < P) gid[k] = order ? order[j] : (uint32_t)j;
  }
  // one gather per Gaussian: the tile rectangle (all zero for a culled Gaussian) also gives the instance count
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    rc[k] = (first + k < P) ? rect[gid[k]] : make_ushort4(0, 0, 0, 0);
    dk[k] = (MODE == 1 && first + k < P) ? dkeys[gid[k]] : 0u;
    cnt[k] = (uint32_t)(rc[k].z - rc[k].x) * (uint32_t)(rc[k].w - rc[k].y);
    mine += cnt[k];
  }
  uint32_t total;
  const uint32_t ex = block_exclusive_scan_256(mine, s_warp, &total);     // syncs inside
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned long long* bst = chain + 2;
    unsigned long long* gst = bst + num_blocks;
    const int j = bid % LB_GROUP, g = bid / LB_GROUP;
    if (lane == 0) st_chain(bst + bid, CH_READY | (unsigned long long)total);
    unsigned long long sum = 0;
    if (lane < j) {
      unsigned long long v;
      do { v = ld_chain(bst + (bid - j + lane)); } while (!(v & CH_READY));
      sum = v & ~CH_READY;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (j == LB_GROUP - 1 && lane == 0) st_chain(gst + g, CH_READY | (sum + total));
    unsigned long long gsum = 0;
    for (int k0 = 0; k0 < g; k0 += 32) {
      if (k0 + lane < g) {
        unsigned long long v;
        do { v = ld_chain(gst + (k0 + lane)); } while (!(v & CH_READY));
        gsum += v & ~CH_READY;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
    const unsigned long long base = sum + gsum;
    if (lane == 0) {
      s_base = (uint32_t)base;
      if (bid == num_blocks - 1) {
        const unsigned long long D = base + total;
        counts[CNT_D] = D > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)D;
        counts[CNT_OVERFLOW] = D > (unsigned long long)D_cap ? 1u : 0u;
        counts[CNT_VISIBLE] = 0; counts[CNT_MAXTILES] = 0;   // reserved; the whole row is copied to the host
        counts[4] = (uint32_t)mode;
        counts[CNT_PREFILTER] = *flag_word;     // points culled although the caller declared them prefiltered
        counts[6] = 0; counts[7] = 0;
      }
    }
  }
  __syncthreads();
  uint32_t off = s_base + ex;
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
    long long o = offk[k];
    for (int y = rc[k].y; y < rc[k].w; ++y)
      for (int x = rc[k].x; x < rc[k].z; ++x) put(o++, (uint32_t)(y * gx + x), dk[k], gid[k]);
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
      const ushort4 rb = rect[g];
      const uint32_t dkb = (MODE == 1) ? dkeys[g] : 0u;
      const uint32_t wx = (uint32_t)(rb.z - rb.x);
      for (uint32_t i = lane; i < c; i += 32) {
        const uint32_t y = rb.y + i / wx, x = rb.x + i % wx;
        put((long long)o0 + i, y * (uint32_t)gx + x, dkb, g);
      }
    }
  }
  }  // ticket loop
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
// `encoded`: ranges come from the RANGES pass of the sort (x = ~start, untouched tiles all zero) and are decoded
// in place first, so that every consumer sees the reference's values (start, end; (0, 0) for an empty tile).
__global__ void __launch_bounds__(1024)
tile_order_kernel(uint2* __restrict__ ranges, int T, uint32_t* __restrict__ order, int encoded) {
  __shared__ uint32_t s_cnt[33], s_cur[33];
  if (threadIdx.x < 33) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    uint2 r = ranges[t];
    if (encoded) {
      r = r.y == 0u ? make_uint2(0u, 0u) : make_uint2(~r.x, r.y);
      ranges[t] = r;                      // re-read below by the same thread
    }
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
//             [MAX_PASSES][num_blocks + num_groups][256] look-back status words (blocks, then groups)
static size_t radix_status_words(long long n_cap) {
  const long long nb = radix_blocks(n_cap > 0 ? n_cap : 1);
  return (size_t)(nb + radix_groups(nb)) * 256;
}
size_t radix_tmp_bytes(long long n_cap) {
  return (size_t)(MAX_PASSES * 256 + 256 + (size_t)MAX_PASSES * radix_status_words(n_cap)) * sizeof(uint32_t);
}

static size_t radix_used_bytes(long long n_cap, int passes) {
  return (size_t)(MAX_PASSES * 256 + 256 + (size_t)passes * radix_status_words(n_cap)) * sizeof(uint32_t);
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
// fails the near-plane test; the emission kernel publishes it as counts[5].
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
                     bool iota_vals, bool hist0_ready, void* tmp, bool debug, cudaStream_t st, uint2* ranges) {
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
      GSB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<KeyT, true, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      GSB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<KeyT, false, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      GSB_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<KeyT, false, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    uint32_t* stp = status + (size_t)p * radix_status_words(n_cap);
    uint32_t* gstp = stp + (size_t)nb * 256;
    uint32_t* gnext = p + 1 < passes ? ghist + (p + 1) * 256 : nullptr;
    if (p == 0 && iota_vals)
      radix_onesweep_kernel<KeyT, true, false><<<nb, RS_THREADS, smem, st>>>(
          kin, vin, kout, vout, d_n, n_cap, 8 * p, ghist + p * 256, gnext, stp, gstp, counters + p, nullptr);
    else if (p == passes - 1 && ranges)
      radix_onesweep_kernel<KeyT, false, true><<<nb, RS_THREADS, smem, st>>>(
          kin, vin, kout, vout, d_n, n_cap, 8 * p, ghist + p * 256, gnext, stp, gstp, counters + p, ranges);
    else
      radix_onesweep_kernel<KeyT, false, false><<<nb, RS_THREADS, smem, st>>>(
          kin, vin, kout, vout, d_n, n_cap, 8 * p, ghist + p * 256, gnext, stp, gstp, counters + p, nullptr);
    GSB_POST_LAUNCH(debug, st, "radix_onesweep_kernel");
  }
  return GSB_OK;
}

template int radix_sort_pairs<uint32_t>(long long, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*,
                                        uint32_t*, uint32_t*, uint32_t*, int, bool, bool, void*, bool, cudaStream_t,
                                        uint2*);
template int radix_sort_pairs<uint64_t>(long long, const uint32_t*, const uint64_t*, const uint32_t*, uint64_t*,
                                        uint32_t*, uint64_t*, uint32_t*, int, bool, bool, void*, bool, cudaStream_t,
                                        uint2*);

int launch_bin_sort(const View& v, int P, void* saved, void* scratch, const GsbLayout& L,
                    long long D_cap, int mode, uint32_t* host_counts, cudaEvent_t event, bool debug,
                    cudaStream_t st) {
  uint32_t* counts = at<uint32_t>(saved, L.off_counts);
  uint32_t* point_list = at<uint32_t>(saved, L.off_point_list);
  uint2* ranges = at<uint2>(saved, L.off_ranges);
  const ushort4* rect = at<ushort4>(scratch, L.off_rect);
  const uint32_t* dkeys = at<uint32_t>(scratch, L.off_dkeys0);   // written by preprocess, kept intact
  unsigned long long* chain = at<unsigned long long>(scratch, L.off_blocksums);
  void* hist = at<char>(scratch, L.off_hist);
  const int T = v.gx * v.gy;
  if (mode != GSB_BIN_TWO_LEVEL && mode != GSB_BIN_FLAT64) return GSB_E_INVALID;
  GSB_CUDA(cudaMemsetAsync(ranges, 0, (size_t)T * sizeof(uint2), st));
  auto publish = [&]() -> int {
    if (host_counts)
      GSB_CUDA(cudaMemcpyAsync(host_counts, counts, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (event) GSB_CUDA(cudaEventRecord(event, st));
    return GSB_OK;
  };
  if (P == 0) {
    GSB_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(uint32_t), st));
    tile_order_kernel<<<1, 1024, 0, st>>>(ranges, T, at<uint32_t>(saved, L.off_tile_order), 0);
    GSB_POST_LAUNCH(debug, st, "tile_order_kernel");
    return publish();
  }
  const int sc_blocks = (P + SC_TILE - 1) / SC_TILE;
  const int tile_bits = bits_for((unsigned)(T - 1));
  const int cap_blocks = (int)((D_cap + 255) / 256);
  uint32_t* alt_vals = at<uint32_t>(scratch, L.off_tvals_alt);
  const bool two_level = mode == GSB_BIN_TWO_LEVEL;
  int rc;

  const uint32_t* order = nullptr;
  if (two_level) {
    // (1) Gaussians by depth key (ties by index: the sort is stable and values start as iota)
    uint32_t* kA = at<uint32_t>(scratch, L.off_dkeys1);
    uint32_t* vA = at<uint32_t>(scratch, L.off_didx0);
    uint32_t* kB = at<uint32_t>(scratch, L.off_dkeys2);
    uint32_t* vB = at<uint32_t>(scratch, L.off_didx1);
    prof_begin(GSB_STAGE_DEPTH_SORT, st);
    // digit-0 histogram of the depth keys was accumulated by preprocess_fwd (after radix_prepare)
    rc = radix_sort_pairs<uint32_t>(P, nullptr, dkeys, nullptr, kA, vA, kB, vB, 32, true, true, hist, debug, st,
                                    nullptr);
    prof_end(GSB_STAGE_DEPTH_SORT, st);
    if (rc) return rc;
    order = vB;  // 4 passes -> result in B
  }
  // (2) offsets in emission order, D, overflow flag and the instances themselves, one kernel.  Two-level mode
  // emits (tile id, Gaussian id) in (depth, id, tile) order and then partitions stably by tile; flat mode emits
  // (tile<<32|depth, Gaussian id) in Gaussian order and sorts the 64-bit keys (the reference's structure).
  prof_begin(GSB_STAGE_SCAN_EMIT, st);
  const int end_bit = two_level ? tile_bits : 32 + tile_bits;
  const int passes = radix_num_passes(end_bit);
  const bool inA = radix_result_in_A(passes);
  uint32_t* tkA = at<uint32_t>(scratch, L.off_tkeys0);
  uint32_t* tkB = at<uint32_t>(scratch, L.off_tkeys1);
  uint64_t* kA64 = at<uint64_t>(scratch, L.off_keys64_0);
  uint64_t* kB64 = at<uint64_t>(scratch, L.off_keys64_1);
  uint32_t* tvA = inA ? point_list : alt_vals;
  uint32_t* tvB = inA ? alt_vals : point_list;
  int sms = 0;
  { const int rc_sm = device_sm_count(&sms); if (rc_sm) return rc_sm; }
  const int emit_grid = sc_blocks < sms * GSB_EMIT_MINB ? sc_blocks : sms * GSB_EMIT_MINB;
  GSB_CUDA(cudaMemsetAsync(chain, 0, emit_chain_words(sc_blocks) * sizeof(unsigned long long), st));
  // the prefiltered-violation flag of preprocess_fwd sits in the sort's counter row, which the next line zeroes
  // for the tile sort: park it in the spare chain word first
  GSB_CUDA(cudaMemcpyAsync(chain + 1, radix_flag_word(hist), sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  if ((rc = radix_prepare(D_cap, end_bit, hist, st))) return rc;
  if (two_level)
    emit_kernel<0><<<emit_grid, RS_THREADS, 0, st>>>(order, P, chain, sc_blocks, rect, dkeys, v.gx, D_cap, tkB,
                                                     nullptr, tvB, radix_hist0(hist), counts, mode,
                                                     reinterpret_cast<const uint32_t*>(chain + 1));
  else
    emit_kernel<1><<<emit_grid, RS_THREADS, 0, st>>>(nullptr, P, chain, sc_blocks, rect, dkeys, v.gx, D_cap,
                                                     nullptr, kB64, tvB, radix_hist0(hist), counts, mode,
                                                     reinterpret_cast<const uint32_t*>(chain + 1));
  GSB_POST_LAUNCH(debug, st, "emit_kernel");
  if ((rc = publish())) return rc;
  prof_end(GSB_STAGE_SCAN_EMIT, st);
  // (3) the sort; its last pass also writes the tile ranges
  prof_begin(GSB_STAGE_TILE_SORT, st);
  if (two_level)
    rc = radix_sort_pairs<uint32_t>(D_cap, counts + CNT_D, tkB, tvB, tkA, tvA, tkB, tvB, end_bit, false, true, hist,
                                    debug, st, ranges);
  else
    rc = radix_sort_pairs<uint64_t>(D_cap, counts + CNT_D, kB64, tvB, kA64, tvA, kB64, tvB, end_bit, false, true,
                                    hist, debug, st, ranges);
  prof_end(GSB_STAGE_TILE_SORT, st);
  if (rc) return rc;
  ProfScope pr(GSB_STAGE_RANGES, st);
  if (passes == 0) {
    // a single-tile image in two-level mode needs no pass: the emitted order (in B) is already final
    tile_ranges_kernel<uint32_t><<<cap_blocks, 256, 0, st>>>(tkB, counts + CNT_D, D_cap, ranges);
    GSB_POST_LAUNCH(debug, st, "tile_ranges_kernel");
  }
  tile_order_kernel<<<1, 1024, 0, st>>>(ranges, T, at<uint32_t>(saved, L.off_tile_order), passes > 0 ? 1 : 0);
  GSB_POST_LAUNCH(debug, st, "tile_order_kernel");
  return GSB_OK;
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

#ifdef GSB_RADIX_TIMING
extern "C" int gsb_debug_radix_timing(long long* out, int words) {
  return (int)cudaMemcpyFromSymbol(out, gsb::g_radix_timing, (size_t)words * sizeof(long long));
}
#endif
