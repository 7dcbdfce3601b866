// mean squared distance to the 3 nearest neighbours of every point — the
// `simple_knn._C.distCUDA2` operator GaussianIP calls once per model in create_from_pcd
// (gaussiansplatting/scene/gaussian_model.py:123, gs_renderer.py:387; reference algorithm:
// gaussiansplatting/submodules/simple-knn/simple_knn.cu:119-221).
//
// The reference sorts by a 30-bit Morton code, builds ONE level of 1024-point boxes and lets every
// thread test all P/1024 boxes (O(P^2/1024) box tests plus whole-box scans).  Here:
//   1. bounding box by ordered-uint atomics, kept on the device (the reference reads it back twice);
//   2. Morton codes, sorted with this library's own onesweep radix sort (binning.cu);
//   3. a 32-ary implicit box hierarchy over the Morton order: 32-point leaves (one warp each), then
//      boxes of 32 children per level up to a root level of <= 32 boxes;
//   4. one WARP per leaf answers its 32 queries together: the 31 leaf mates seed the three best
//      distances, then the warp walks the hierarchy front to back with skip-by-subtree, descending
//      only where some lane's box distance is <= its current third-best, and scans admitted leaves
//      from a 512 B shared-memory slab.
// The result is the EXACT 3-NN mean in the reference's own arithmetic: distances are
// fma(dz,dz, fma(dx,dx, dy*dy)) with d = other - self (what nvcc emits for the reference, checked
// in its SASS), the box distance uses the same expression so pruning is conservative in floating
// point, and the mean is ((b0+b1)+b2)/3 with IEEE division.  The multiset of the three smallest
// distances does not depend on visiting order, so the output is bit-identical to the reference.
#include <float.h>
#include "gsb_common.cuh"

namespace gsb {

constexpr int KNN_MAX_LEVELS = 6;      // 32^6 leaves >> any P that fits in memory
constexpr int KNN_THREADS = 256;

struct KnnTree {
  long long P;
  int n_leaf;                          // ceil(P / 32)
  int n_levels;                        // upper levels above the leaves (0 when n_leaf <= 32)
  int n_box[KNN_MAX_LEVELS];           // boxes per upper level, level 1 = 32 leaves each
  const float4* pts;                   // Morton-sorted points, w = original index bits
  const float4* leaf_box;              // [n_leaf][2]  (min.xyz, max.xyz)
  const float4* box[KNN_MAX_LEVELS];   // [n_box[l]][2]
};

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void knn_bbox_init_kernel(uint32_t* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(KNN_THREADS) knn_bbox_kernel(const float* __restrict__ pts, long long P,
                                                               uint32_t* __restrict__ bbox) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = pts[3 * i + c];
      mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicMin(&bbox[c], f2ord(mn[c]));
      atomicMax(&bbox[3 + c], f2ord(mx[c]));
    }
  }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {   // 10 bits -> every third bit
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

__global__ void __launch_bounds__(KNN_THREADS) knn_morton_kernel(const float* __restrict__ pts, long long P,
                                                                 const uint32_t* __restrict__ bbox,
                                                                 uint32_t* __restrict__ codes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t q[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float lo = ord2f(bbox[c]), hi = ord2f(bbox[3 + c]);
    const float ext = hi - lo;
    float t = ext > 0.0f ? (pts[3 * i + c] - lo) / ext * 1023.0f : 0.0f;
    t = fminf(fmaxf(t, 0.0f), 1023.0f);          // also maps NaN to 0
    q[c] = (uint32_t)t;
  }
  codes[i] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
}

// One warp per leaf: gather its 32 points in Morton order (float4, w = original index) and reduce
// the leaf's box.
__global__ void __launch_bounds__(KNN_THREADS) knn_gather_kernel(const float* __restrict__ pts,
                                                                 const uint32_t* __restrict__ order, long long P,
                                                                 float4* __restrict__ sorted,
                                                                 float4* __restrict__ leaf_box) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < P;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    const uint32_t src = order[i];
    p = make_float4(pts[3 * (size_t)src], pts[3 * (size_t)src + 1], pts[3 * (size_t)src + 2], __uint_as_float(src));
    sorted[i] = p;
  }
  float mn[3] = {valid ? p.x : FLT_MAX, valid ? p.y : FLT_MAX, valid ? p.z : FLT_MAX};
  float mx[3] = {valid ? p.x : -FLT_MAX, valid ? p.y : -FLT_MAX, valid ? p.z : -FLT_MAX};
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  const long long leaf = i >> 5;
  if ((threadIdx.x & 31) == 0 && leaf * 32 < P) {
    leaf_box[2 * leaf] = make_float4(mn[0], mn[1], mn[2], 0.f);
    leaf_box[2 * leaf + 1] = make_float4(mx[0], mx[1], mx[2], 0.f);
  }
}

// One warp per parent box: union of up to 32 child boxes.
__global__ void __launch_bounds__(KNN_THREADS) knn_level_kernel(const float4* __restrict__ child, int n_child,
                                                                float4* __restrict__ parent, int n_parent) {
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= n_parent) return;
  const int c = w * 32 + lane;
  float4 lo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), hi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
  if (c < n_child) { lo = child[2 * (size_t)c]; hi = child[2 * (size_t)c + 1]; }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
    lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
    lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
    hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
    hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
    hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
  }
  if (lane == 0) { parent[2 * (size_t)w] = lo; parent[2 * (size_t)w + 1] = hi; }
}

// Squared distance self -> other in the reference's rounding: fma(dz,dz, fma(dx,dx, dy*dy)) — nvcc keeps the
// SECOND product of `dx*dx + dy*dy` as the plain multiply and fuses the first (seen in the reference's SASS).
__device__ __forceinline__ float dist2_ref(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Lower bound of dist2_ref over every point inside [lo, hi]: per axis the gap is rounded exactly
// like a point difference would be, and rounding is monotone, so bound <= distance in floats.
__device__ __forceinline__ float box_dist2(const float4& lo, const float4& hi, const float4& p) {
  const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, p.x), __fsub_rn(p.x, hi.x)), 0.0f);
  const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, p.y), __fsub_rn(p.y, hi.y)), 0.0f);
  const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, p.z), __fsub_rn(p.z, hi.z)), 0.0f);
  return dist2_ref(dx, dy, dz);
}

__device__ __forceinline__ void keep3(float& b0, float& b1, float& b2, float d) {
  const float t0 = fmaxf(b0, d); b0 = fminf(b0, d);
  const float t1 = fmaxf(b1, t0); b1 = fminf(b1, t0);
  b2 = fminf(b2, t1);
}

__global__ void __launch_bounds__(KNN_THREADS) knn_query_kernel(const KnnTree t, float* __restrict__ mean_dist2) {
  __shared__ float4 slab[KNN_THREADS / 32][32];
  const int wslot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int own = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (own >= t.n_leaf) return;                                   // warp-uniform
  const long long me_i = (long long)own * 32 + lane;
  const bool valid = me_i < t.P;
  const float4 me = valid ? t.pts[me_i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;

  // seed: the other points of the own leaf
  slab[wslot][lane] = me;
  __syncwarp();
  {
    const int cnt = (int)min((long long)32, t.P - (long long)own * 32);
    for (int j = 0; j < cnt; ++j) {
      const float4 q = slab[wslot][j];
      const float d = dist2_ref(__fsub_rn(q.x, me.x), __fsub_rn(q.y, me.y), __fsub_rn(q.z, me.z));
      if (j != lane) keep3(b0, b1, b2, d);
    }
  }
  // lanes past the end never admit a box (bound > -1 always holds)
  if (!valid) b2 = -1.0f;

  int i = 0;
  while (i < t.n_leaf) {
    bool skipped = false;
#pragma unroll
    for (int l = KNN_MAX_LEVELS; l >= 1; --l) {                  // unrolled: t.box[] stays in param space
      const int span_bits = 5 * l;
      if (l > t.n_levels) continue;
      if (i & ((1 << span_bits) - 1)) continue;                  // not the first leaf of a level-l box
      const int b = i >> span_bits;
      const float4 lo = t.box[l - 1][2 * (size_t)b], hi = t.box[l - 1][2 * (size_t)b + 1];
      const bool need = !(box_dist2(lo, hi, me) > b2);
      if (!__any_sync(0xffffffffu, need)) { i += 1 << span_bits; skipped = true; break; }
    }
    if (skipped) continue;
    if (i != own) {
      const float4 lo = t.leaf_box[2 * (size_t)i], hi = t.leaf_box[2 * (size_t)i + 1];
      const bool need = !(box_dist2(lo, hi, me) > b2);
      if (__any_sync(0xffffffffu, need)) {
        const long long base = (long long)i * 32;
        const int cnt = (int)min((long long)32, t.P - base);
        __syncwarp();
        if (lane < cnt) slab[wslot][lane] = t.pts[base + lane];
        __syncwarp();
        if (need) {
#pragma unroll 4
          for (int j = 0; j < cnt; ++j) {
            const float4 q = slab[wslot][j];
            keep3(b0, b1, b2, dist2_ref(__fsub_rn(q.x, me.x), __fsub_rn(q.y, me.y), __fsub_rn(q.z, me.z)));
          }
        }
      }
    }
    ++i;
  }
  if (valid)
    mean_dist2[__float_as_uint(me.w)] = __fdiv_rn(__fadd_rn(__fadd_rn(b0, b1), b2), 3.0f);
}

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

struct KnnLayout {
  size_t bbox, codes, kA, vA, kB, vB, sorted, leaf_box, box[KNN_MAX_LEVELS], radix, total;
  int n_leaf, n_levels, n_box[KNN_MAX_LEVELS];
};

static int knn_layout(long long P, KnnLayout& L) {
  if (P < 0 || P > 0x7fffffffLL) return GSB_E_INVALID;
  const size_t n = (size_t)(P > 0 ? P : 1);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = up256(o + bytes); return r; };
  L.bbox = take(8 * sizeof(uint32_t));
  L.codes = take(n * 4); L.kA = take(n * 4); L.vA = take(n * 4); L.kB = take(n * 4); L.vB = take(n * 4);
  L.sorted = take(n * sizeof(float4));
  L.n_leaf = (int)((n + 31) / 32);
  L.leaf_box = take((size_t)L.n_leaf * 2 * sizeof(float4));
  L.n_levels = 0;
  int c = L.n_leaf;
  for (int l = 0; l < KNN_MAX_LEVELS; ++l) { L.n_box[l] = 0; L.box[l] = 0; }
  while (c > 32) {
    if (L.n_levels >= KNN_MAX_LEVELS) return GSB_E_INVALID;
    c = (c + 31) / 32;
    L.n_box[L.n_levels] = c;
    L.box[L.n_levels] = take((size_t)c * 2 * sizeof(float4));
    ++L.n_levels;
  }
  L.radix = take(radix_tmp_bytes((long long)n));
  L.total = o;
  return GSB_OK;
}

size_t knn_scratch_bytes(long long P) {
  KnnLayout L;
  if (knn_layout(P, L)) return 0;
  return L.total;
}

int launch_knn_dist2(long long P, const float* points, float* mean_dist2, void* scratch, size_t scratch_bytes,
                     bool debug, cudaStream_t st) {
  if (P == 0) return GSB_OK;
  KnnLayout L;
  int rc = knn_layout(P, L);
  if (rc) return rc;
  if (!points || !mean_dist2 || !scratch || scratch_bytes < L.total) return GSB_E_INVALID;
  uint32_t* bbox = at<uint32_t>(scratch, L.bbox);
  uint32_t* codes = at<uint32_t>(scratch, L.codes);
  float4* sorted = at<float4>(scratch, L.sorted);
  float4* leaf_box = at<float4>(scratch, L.leaf_box);
  const int blocks = (int)((P + KNN_THREADS - 1) / KNN_THREADS);

  knn_bbox_init_kernel<<<1, 32, 0, st>>>(bbox);
  GSB_POST_LAUNCH(debug, st, "knn_bbox_init_kernel");
  knn_bbox_kernel<<<min(blocks, 148 * 8), KNN_THREADS, 0, st>>>(points, P, bbox);
  GSB_POST_LAUNCH(debug, st, "knn_bbox_kernel");
  knn_morton_kernel<<<blocks, KNN_THREADS, 0, st>>>(points, P, bbox, codes);
  GSB_POST_LAUNCH(debug, st, "knn_morton_kernel");
  uint32_t* kA = at<uint32_t>(scratch, L.kA); uint32_t* vA = at<uint32_t>(scratch, L.vA);
  uint32_t* kB = at<uint32_t>(scratch, L.kB); uint32_t* vB = at<uint32_t>(scratch, L.vB);
  rc = radix_sort_pairs<uint32_t>(P, nullptr, codes, nullptr, kA, vA, kB, vB, 30, true, false,
                                  at<char>(scratch, L.radix), debug, st);
  if (rc) return rc;
  const uint32_t* order = radix_result_in_A(radix_num_passes(30)) ? vA : vB;

  const int pad_blocks = (int)(((long long)L.n_leaf * 32 + KNN_THREADS - 1) / KNN_THREADS);
  knn_gather_kernel<<<pad_blocks, KNN_THREADS, 0, st>>>(points, order, P, sorted, leaf_box);
  GSB_POST_LAUNCH(debug, st, "knn_gather_kernel");

  KnnTree t;
  t.P = P; t.n_leaf = L.n_leaf; t.n_levels = L.n_levels; t.pts = sorted; t.leaf_box = leaf_box;
  const float4* child = leaf_box;
  int n_child = L.n_leaf;
  for (int l = 0; l < KNN_MAX_LEVELS; ++l) { t.n_box[l] = L.n_box[l]; t.box[l] = nullptr; }
  for (int l = 0; l < L.n_levels; ++l) {
    float4* parent = at<float4>(scratch, L.box[l]);
    const int nb = (int)(((long long)L.n_box[l] * 32 + KNN_THREADS - 1) / KNN_THREADS);
    knn_level_kernel<<<nb, KNN_THREADS, 0, st>>>(child, n_child, parent, L.n_box[l]);
    GSB_POST_LAUNCH(debug, st, "knn_level_kernel");
    t.box[l] = parent; child = parent; n_child = L.n_box[l];
  }
  knn_query_kernel<<<pad_blocks, KNN_THREADS, 0, st>>>(t, mean_dist2);
  GSB_POST_LAUNCH(debug, st, "knn_query_kernel");
  return GSB_OK;
}

}  // namespace gsb
