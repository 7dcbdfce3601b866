// Internal declarations shared by the sm_100a kernels behind include/gsb.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gsb.h"

namespace gsb {

constexpr int TILE_X = 16;
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr float NEAR_Z = 0.2f;
constexpr float LOWPASS = 0.3f;
constexpr float ALPHA_CAP = 0.99f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float T_MIN = 0.0001f;

// Per-Gaussian record gathered by the blend kernels: 48 B, 16 B aligned, three float4.
struct __align__(16) Geom {
  float x, y, extx, exty;       // pixel centre, conservative half-extent of {alpha >= 1/255} in px
  float ca, cb, cc, opacity;    // conic, pre-scaled: -0.5 log2e A, -log2e B, -0.5 log2e C (see gauss_exponent2); opacity
  float depth, r, g, b;         // view z, colour
};
static_assert(sizeof(Geom) == 48, "Geom must be 48 bytes");

// Per-Gaussian gradient record accumulated by render_bwd (atomics land in one 48 B span).
// With s = dL/dG * G per (pixel, Gaussian) and d = centre - pixel, the first five slots are the
// raw moments sum(s dx), sum(s dy), sum(s dx^2), sum(s dx dy), sum(s dy^2); preprocess_bwd maps
// them to dL/d(ndc xy) and dL/d(conic).
struct __align__(16) GGrad {
  float sx, sy, sxx, sxy;
  float syy, dop, ddepth, dr;  // ..., dL/dopacity, dL/ddepth, dL/dred
  float dg, db, pad0, pad1;
};
static_assert(sizeof(GGrad) == 48, "GGrad must be 48 bytes");

enum Counts { CNT_D = 0, CNT_OVERFLOW = 1, CNT_VISIBLE = 2, CNT_MAXTILES = 3, CNT_MODE = 4, CNT_PREFILTER = 5 };

struct View {           // settings with device pointers, passed by value to kernels
  int H, W, gx, gy;     // image size, tile grid
  float tanfovx, tanfovy, focal_x, focal_y, scale_mod;
  int sh_degree;
  int raw;              // GSB_RAW_* bits
  int prefiltered;      // the caller asserts that every point passes the frustum test
  const float* bg;
  const float* view;
  const float* proj;
  const float* campos;
  const float* tanfov_dev;   // optional device [2] = {tanfovx, tanfovy}: overrides the by-value intrinsics (see GsbSettings)
};

inline View make_view(const GsbSettings* s) {
  View v;
  v.H = s->image_height; v.W = s->image_width;
  v.gx = (v.W + TILE_X - 1) / TILE_X; v.gy = (v.H + TILE_Y - 1) / TILE_Y;
  v.tanfovx = s->tanfovx; v.tanfovy = s->tanfovy;
  v.focal_x = (float)v.W / (2.0f * s->tanfovx);
  v.focal_y = (float)v.H / (2.0f * s->tanfovy);
  v.scale_mod = s->scale_modifier; v.sh_degree = s->sh_degree; v.raw = s->raw_inputs;
  v.prefiltered = s->prefiltered;
  v.bg = s->bg; v.view = s->viewmatrix; v.proj = s->projmatrix; v.campos = s->campos;
  v.tanfov_dev = s->tanfov_dev;
  return v;
}

// Pixel of a lane inside its warp's 8x4 block (blend kernels, hit-record masks): lanes 0-15 are the LEFT 4x4 pixels,
// lanes 16-31 the RIGHT 4x4 pixels, row-major inside each half.  The two halves of a warp walk their own hit
// sequences (a splat of ~3.7 px touches 1.4 of the two 4x4 halves on average, but its evaluation costs a warp
// instruction either way), so bits 0-15 / 16-31 of a record's mask belong to different streams.
__device__ __forceinline__ int lane_px(int lane) { return (lane & 3) | ((lane >> 4) << 2); }   // 0..7
__device__ __forceinline__ int lane_py(int lane) { return (lane >> 2) & 3; }                   // 0..3

// Intrinsics of a view: k[0..3] = tanfovx, tanfovy, focal_x, focal_y.  With GsbSettings.tanfov_dev they are read from
// device memory (a captured CUDA graph can then be replayed with new cameras) and the focal lengths are formed
// with the same fp32 operations make_view uses on the host.
__device__ __forceinline__ void load_intrinsics(const View& v, float* k) {
  float tx = v.tanfovx, ty = v.tanfovy, fx = v.focal_x, fy = v.focal_y;
  if (v.tanfov_dev) {
    tx = v.tanfov_dev[0]; ty = v.tanfov_dev[1];
    fx = __fdiv_rn((float)v.W, __fmul_rn(2.0f, tx));
    fy = __fdiv_rn((float)v.H, __fmul_rn(2.0f, ty));
  }
  k[0] = tx; k[1] = ty; k[2] = fx; k[3] = fy;
}

// The Geom record stores the conic PRE-SCALED for the blend kernels: with qa = -0.5 log2(e) A, qb = -log2(e) B,
// qc = -0.5 log2(e) C the Gaussian weight is G = 2^(qa dx^2 + qb dx dy + qc dy^2): five instructions and a bare
// ex2 instead of eight plus the log2(e) multiply.  preprocess_bwd un-scales when it needs A, B, C.
constexpr float LOG2E = 1.4426950408889634f;
constexpr float CONIC_SCALE_AC = -0.5f * LOG2E;
constexpr float CONIC_SCALE_B = -LOG2E;
__device__ __forceinline__ float gauss_exponent2(float qa, float qb, float qc, float dx, float dy) {
  return fmaf(qc * dy, dy, fmaf(qa, dx, qb * dy) * dx);
}
__device__ __forceinline__ float exp2_blend(float e) {     // 2^e, flush-to-zero (see exp_blend)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(e));
  return y;
}

// exp() of the blend kernels: ex2.approx on x*log2(e), flush-to-zero.  __expf adds a range fix-up so that results
// below 2^-126 come out as denormals (4 extra instructions per evaluation); such a Gaussian weight is ~1e-38,
// its alpha is far below 1/255 and the pair is skipped either way, so the fix-up buys nothing here.
__device__ __forceinline__ float exp_blend(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

// 1/x for x in [0.01, 1] (x = 1 - alpha, alpha <= 0.99): the bare MUFU.RCP; __fdividef wraps the same instruction in
// a denormal-range fix-up (4 extra instructions) that cannot trigger here.
__device__ __forceinline__ float rcp_blend(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Activations of the raw model parameters (GSB_RAW_*), written the way torch evaluates them.
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float4 act_normalize(float4 q, float* norm_out = nullptr) {
  const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);   // F.normalize eps
  if (norm_out) *norm_out = n;
  return make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
}

template <typename T>
inline T* at(void* base, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(base) + off); }
template <typename T>
inline const T* at(const void* base, size_t off) {
  return reinterpret_cast<const T*>(static_cast<const char*>(base) + off);
}

// error plumbing (api.cu)
int cuda_fail(cudaError_t e, const char* what);
void count_launch();
// number of SMs of the CURRENT device (cached per device; api.cu): persistent kernels size their grids with it
int device_sm_count(int* sms);
// stage timing (api.cu): no-ops unless gsb_profile_enable(1)
void prof_begin(int stage, cudaStream_t st);
void prof_end(int stage, cudaStream_t st);
struct ProfScope {
  int stage; cudaStream_t st;
  ProfScope(int s, cudaStream_t t) : stage(s), st(t) { prof_begin(stage, st); }
  ~ProfScope() { prof_end(stage, st); }
};
#define GSB_CUDA(call)                                          \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return gsb::cuda_fail(e__, #call);  \
  } while (0)
// after a kernel launch: always catch launch errors; in debug mode also synchronise
#define GSB_POST_LAUNCH(dbg, st, name)                                              \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    gsb::count_launch();                                                            \
    if (e__ == cudaSuccess && (dbg)) e__ = cudaStreamSynchronize(st);               \
    if (e__ != cudaSuccess) return gsb::cuda_fail(e__, name);                       \
  } while (0)

// ---- stage launchers (one per .cu) -------------------------------------------------------
int launch_preprocess_fwd(const View& v, int P, int K, const float* means3D, const float* scales,
                          const float* rots, const float* opac, const float* shs,
                          const float* colors, const float* cov3D, int32_t* radii, Geom* geom,
                          uint8_t* clamped, ushort4* rect, uint32_t* tiles, uint32_t* dkeys,
                          void* radix_tmp, bool debug, cudaStream_t st);

int launch_bin_sort(const View& v, int P, void* saved, void* scratch, const GsbLayout& L,
                    long long D_cap, int mode, uint32_t* host_counts, cudaEvent_t event, bool debug,
                    cudaStream_t st);

int launch_render_fwd(const View& v, int P, const Geom* geom, const uint32_t* point_list,
                      const uint2* ranges, const uint32_t* tile_order, float* color, float* depth, float* alpha,
                      uint32_t* n_contrib, float* final_T, uint2* hits, uint32_t* hit_count, int variant,
                      bool debug, cudaStream_t st);

int launch_render_bwd(const View& v, int P, const Geom* geom, const uint32_t* point_list,
                      const uint2* ranges, const uint32_t* tile_order, const uint32_t* n_contrib,
                      const float* final_T, const uint2* hits, const uint32_t* hit_count, const float* dL_dcolor,
                      const float* dL_ddepth, const float* dL_dalpha, GGrad* ggrad, int kind, bool packed,
                      bool debug, cudaStream_t st);

// reference-structure stand-in blend kernels (standin.cu; measurement context and cross-check only)
int launch_standin_fwd(const View& v, const Geom* geom, const uint32_t* point_list, const uint2* ranges,
                       float* color, float* depth, float* alpha, uint32_t* n_contrib, float* final_T, bool debug,
                       cudaStream_t st);
int launch_standin_bwd(const View& v, int P, const Geom* geom, const uint32_t* point_list, const uint2* ranges,
                       const uint32_t* n_contrib, const float* final_T, const float* dL_dcolor,
                       const float* dL_ddepth, const float* dL_dalpha, GGrad* ggrad, bool debug, cudaStream_t st);

struct BwdView {          // one view of a (batched) preprocess backward
  View v;
  const Geom* geom;
  const uint8_t* clamped;
  const GGrad* ggrad;
  const int32_t* radii;
};
struct BwdBatch {
  int V;
  BwdView a[GSB_MAX_VIEWS];
};
int launch_preprocess_bwd(const BwdBatch& B, int P, int K, const float* means3D, const float* scales,
                          const float* rots, const float* shs, const float* colors, const float* cov3D,
                          float* dmeans3D, float* dmeans2D, float* dshs, float* dcolors, float* dopac,
                          float* dscales, float* drots, float* dcov3D, int accumulate, bool debug,
                          cudaStream_t st);

int launch_adam_stats(int G, float* const* p, const float* const* g, float* const* m, float* const* v,
                      const long long* n, const float* lr, double beta1, double beta2, double eps,
                      const long long* steps, float grad_scale, long long stats_n, const float* viewspace_grad, const int32_t* radii,
                      float* xyz_gradient_accum, float* denom, float* max_radii2D, cudaStream_t st);

// radix sort (binning.cu)
size_t radix_tmp_bytes(long long n_cap);
template <typename KeyT>
int radix_sort_pairs(long long n_cap, const uint32_t* d_n, const KeyT* src_keys, const uint32_t* src_vals,
                     KeyT* keysA, uint32_t* valsA, KeyT* keysB, uint32_t* valsB, int end_bit,
                     bool iota_vals, bool hist0_ready, void* tmp, bool debug, cudaStream_t st,
                     uint2* ranges = nullptr);   // ranges: the last pass also writes encoded tile ranges (binning.cu)
int radix_prepare(long long n_cap, int end_bit, void* tmp, cudaStream_t st);
uint32_t* radix_hist0(void* tmp);
uint32_t* radix_flag_word(void* tmp);
bool radix_result_in_A(int passes);
int launch_debug_sorted_keys(const View& v, int P, const void* saved, const void* scratch, const GsbLayout& L,
                             long long D_cap, uint64_t* keys_out, cudaStream_t st);
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t st);
int radix_num_passes(int end_bit);
// fused gradient exchange (preprocess_bwd.cu)
struct ExchangePeers {
  int world, rank;
  long long rows_per_rank;
  char* base[GSB_MAX_RANKS];        // base address of every rank's copy of the symmetric buffer, as mapped HERE
  // optional per-launch extras of the fused exchange (gsb_exchange_set_aux): the MAX over ranks of the per-Gaussian
  // radii and the SUM over ranks of a scalar (the step's loss) ride in the same kernel, so a step needs no NCCL call
  int32_t* radii_max;               // [P] in the symmetric buffer (multicast address in mode 2, local copy in mode 3)
  const float* scalar_in;           // this rank's scalar, device
  float* scalar_out;                // one float in the symmetric buffer (owned by rank 0 in mode 3)
};
struct ExchangeSegments {
  long long off[GSB_EXCHANGE_MAX_SEGMENTS], count[GSB_EXCHANGE_MAX_SEGMENTS];
};
int set_exchange_peers(int world, int rank, long long rows_per_rank, const void* const* bases);
int set_exchange_aux(int32_t* radii_max, const float* scalar_in, float* scalar_out);
int launch_exchange_gather(const float* local, float* mc, int n_seg, const long long* off, const long long* count,
                           cudaStream_t st);
// densify/prune data movement (compact.cu)
size_t mask_index_tmp_bytes(long long n);
int launch_mask_to_index(long long n, const uint8_t* mask, int64_t* index, uint32_t* count, void* tmp, cudaStream_t st);
int launch_gather_rows(int n_tensors, const float* const* src, float* const* dst, const int* widths, long long n_rows,
                       const int64_t* index, long long dst_row0, cudaStream_t st);
// 3-NN mean distance (knn.cu)
size_t knn_scratch_bytes(long long P);
int launch_knn_dist2(long long P, const float* points, float* mean_dist2, void* scratch, size_t scratch_bytes,
                     bool debug, cudaStream_t st);

}  // namespace gsb
