// extern "C" surface of libgsb.so (include/gsb.h): argument validation, workspace layout,
// stage chaining on the caller's stream.  No torch types, no allocation, no exceptions.
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "gsb_common.cuh"

namespace gsb {

static thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return GSB_E_CUDA;
}

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_blend_variant{0};
static std::atomic<int> g_fwd_variant{0};     // 0 per-hit forward blend (cp.async), 1 transposed (two-phase), 2 per-hit with TMA gather4
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- stage profiler ------------------------------------------------------------------------
namespace {
struct ProfRec { int stage; cudaEvent_t a, b; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;      // recorded, not yet resolved
std::vector<cudaEvent_t> g_prof_pool;  // reusable events
cudaEvent_t g_prof_open[GSB_NUM_STAGES];
double g_prof_ms[GSB_NUM_STAGES];
int g_prof_calls[GSB_NUM_STAGES];
cudaEvent_t prof_get_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

int device_sm_count(int* sms) {
  static std::atomic<int> cache[64];      // zero-initialised; index = device ordinal (mod 64)
  int dev = 0;
  GSB_CUDA(cudaGetDevice(&dev));
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    GSB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    cache[dev & 63].store(n, std::memory_order_relaxed);
  }
  *sms = n;
  return GSB_OK;
}

void prof_begin(int stage, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e = prof_get_event();
  cudaEventRecord(e, st);
  g_prof_open[stage] = e;
}
void prof_end(int stage, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_open[stage]) return;
  cudaEvent_t e = prof_get_event();
  cudaEventRecord(e, st);
  g_prof_recs.push_back({stage, g_prof_open[stage], e});
  g_prof_open[stage] = nullptr;
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static bool settings_ok(const GsbSettings* s) {
  return s && s->image_height > 0 && s->image_width > 0 && s->sh_degree >= 0 && s->sh_degree <= 3 && s->bg &&
         s->viewmatrix && s->projmatrix && s->campos && s->tanfovx > 0.f && s->tanfovy > 0.f &&
         s->image_width <= 65535 * TILE_X && s->image_height <= 65535 * TILE_Y &&
         (s->raw_inputs & ~(GSB_RAW_OPACITY | GSB_RAW_SCALE | GSB_RAW_ROTATION)) == 0;
}

static int layout(int P, int H, int W, long long D_cap, GsbLayout* L) {
  if (P < 0 || H <= 0 || W <= 0 || D_cap < 0 || !L) return GSB_E_INVALID;
  if (D_cap > 0xFFFFFFF0ll) return GSB_E_UNSUPPORTED;
  memset(L, 0, sizeof(*L));
  const size_t T = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
  const size_t HW = (size_t)H * W, Pz = (size_t)P, Dz = (size_t)D_cap;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  L->off_geom = take(Pz * sizeof(Geom));
  L->off_clamped = take(Pz);
  L->off_counts = take(8 * sizeof(uint32_t));
  L->off_point_list = take(Dz * 4);
  L->off_ranges = take(T * sizeof(uint2));
  L->off_n_contrib = take(HW * 4);
  L->off_final_T = take(HW * 4);
  L->off_tile_order = take(T * 4);
  L->off_hit_count = take(T * 8 * 4);
  L->saved_bytes_forward_only = o;
  L->off_hits = take(Dz * 8 * sizeof(uint2));
  L->saved_bytes = o;
  o = 0;
  L->off_rect = take(Pz * 8);
  L->off_tiles = take(Pz * 4);
  L->off_dkeys0 = take(Pz * 4);
  L->off_dkeys1 = take(Pz * 4);
  L->off_dkeys2 = take(Pz * 4);
  L->off_didx0 = take(Pz * 4);
  L->off_didx1 = take(Pz * 4);
  L->off_offsets = take(Pz * 4);
  L->off_blocksums = take((Pz / 512 + Pz / 16384 + 8) * 8);   // scan chain of the emission kernel (u64 words)
  const size_t nmax = Dz > Pz ? Dz : Pz;
  L->off_hist = take(radix_tmp_bytes((long long)nmax));
  L->off_tkeys0 = take(Dz * 4);
  L->off_tkeys1 = take(Dz * 4);
  L->off_tvals_alt = take(Dz * 4);
  L->off_keys64_0 = take(Dz * 8);
  L->off_keys64_1 = take(Dz * 8);
  L->off_ggrad = take(Pz * sizeof(GGrad));
  L->scratch_bytes = o;
  return GSB_OK;
}

}  // namespace gsb

using namespace gsb;

extern "C" {

int gsb_abi_version(void) { return GSB_ABI_VERSION; }

const char* gsb_strerror(int code) {
  switch (code) {
    case GSB_OK: return "ok";
    case GSB_E_INVALID: return "invalid argument";
    case GSB_E_CUDA: return "CUDA error";
    case GSB_E_CAPACITY: return "instance capacity exceeded";
    case GSB_E_UNSUPPORTED: return "unsupported size";
    default: return "unknown error";
  }
}

const char* gsb_last_cuda_error(void) { return g_cuda_err; }

int gsb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
  g_prof_recs.clear();
  for (int i = 0; i < GSB_NUM_STAGES; ++i) { g_prof_ms[i] = 0.0; g_prof_calls[i] = 0; g_prof_open[i] = nullptr; }
  g_prof_on = on != 0;
  return GSB_OK;
}

int gsb_profile_read(float* ms_out, int* calls_out) {
  if (!ms_out || !calls_out) return GSB_E_INVALID;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) {
    GSB_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    GSB_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    g_prof_ms[r.stage] += ms;
    g_prof_calls[r.stage] += 1;
    g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
  }
  g_prof_recs.clear();
  for (int i = 0; i < GSB_NUM_STAGES; ++i) { ms_out[i] = (float)g_prof_ms[i]; calls_out[i] = g_prof_calls[i]; }
  return GSB_OK;
}

long long gsb_launch_count(void) { return g_launches.load(); }

int gsb_adam_step_groups(int n_groups, float* const* params, const float* const* grads, float* const* exp_avg,
                         float* const* exp_avg_sq, const long long* counts, const float* lrs, double beta1,
                         double beta2, double eps, const long long* steps, float grad_scale, long long stats_n,
                         const float* viewspace_grad, const int32_t* radii, float* xyz_gradient_accum, float* denom,
                         float* max_radii2D, void* stream) {
  if (n_groups > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !counts || !lrs || !steps)) return GSB_E_INVALID;
  return launch_adam_stats(n_groups, params, grads, exp_avg, exp_avg_sq, counts, lrs, beta1, beta2, eps, steps,
                           grad_scale, stats_n, viewspace_grad, radii, xyz_gradient_accum, denom, max_radii2D,
                           (cudaStream_t)stream);
}

int gsb_adam_step(int n_groups, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const long long* counts, const float* lrs, double beta1, double beta2,
                  double eps, long long step, float grad_scale, long long stats_n, const float* viewspace_grad,
                  const int32_t* radii, float* xyz_gradient_accum, float* denom, float* max_radii2D,
                  void* stream) {
  if (n_groups < 0 || n_groups > GSB_ADAM_MAX_GROUPS || step < 1) return GSB_E_INVALID;
  long long steps[GSB_ADAM_MAX_GROUPS];
  for (int k = 0; k < GSB_ADAM_MAX_GROUPS; ++k) steps[k] = step;
  return gsb_adam_step_groups(n_groups, params, grads, exp_avg, exp_avg_sq, counts, lrs, beta1, beta2, eps, steps,
                              grad_scale, stats_n, viewspace_grad, radii, xyz_gradient_accum, denom, max_radii2D,
                              stream);
}

int gsb_set_blend_variant(int variant) {
  // 10 / 11 / 12 select the forward blend kernel (per-hit with cp.async gathers / transposed / per-hit with TMA
  // gather4 row gathers) and leave the backward choice alone
  if (variant >= 10 && variant <= 13) { g_fwd_variant.store(variant - 10); return GSB_OK; }
  if (variant < 0 || variant > 4) return GSB_E_INVALID;
  g_blend_variant.store(variant);
  return GSB_OK;
}

int gsb_layout(int P, int H, int W, long long D_cap, GsbLayout* out) { return layout(P, H, W, D_cap, out); }

int gsb_preprocess_fwd(const GsbSettings* s, int P, int K, const float* means3D, const float* scales,
                       const float* rotations, const float* opacities, const float* shs,
                       const float* colors_precomp, const float* cov3D_precomp, int32_t* radii_out,
                       void* saved, void* scratch, long long D_cap, void* stream) {
  if (!settings_ok(s) || P < 0) return GSB_E_INVALID;
  if (P == 0) return GSB_OK;
  if (!means3D || !opacities || !radii_out || !saved || !scratch) return GSB_E_INVALID;
  if ((shs != nullptr) == (colors_precomp != nullptr)) return GSB_E_INVALID;
  if ((cov3D_precomp != nullptr) == (scales != nullptr && rotations != nullptr)) return GSB_E_INVALID;
  if (shs && K < (s->sh_degree + 1) * (s->sh_degree + 1)) return GSB_E_INVALID;
  GsbLayout L;
  int rc = layout(P, s->image_height, s->image_width, D_cap, &L);
  if (rc) return rc;
  const View v = make_view(s);
  ProfScope ps(GSB_STAGE_PREPROCESS_FWD, (cudaStream_t)stream);
  return launch_preprocess_fwd(v, P, K, means3D, scales, rotations, opacities, shs, colors_precomp,
                               cov3D_precomp, radii_out, at<Geom>(saved, L.off_geom),
                               at<uint8_t>(saved, L.off_clamped), at<ushort4>(scratch, L.off_rect),
                               at<uint32_t>(scratch, L.off_tiles), at<uint32_t>(scratch, L.off_dkeys0),
                               at<char>(scratch, L.off_hist), s->debug != 0, (cudaStream_t)stream);
}

int gsb_bin_sort(const GsbSettings* s, int P, void* saved, void* scratch, long long D_cap, int mode,
                 uint32_t* host_counts, void* event, void* stream) {
  if (!settings_ok(s) || P < 0 || !saved || !scratch) return GSB_E_INVALID;
  if (mode != GSB_BIN_TWO_LEVEL && mode != GSB_BIN_FLAT64) return GSB_E_INVALID;
  GsbLayout L;
  int rc = layout(P, s->image_height, s->image_width, D_cap, &L);
  if (rc) return rc;
  return launch_bin_sort(make_view(s), P, saved, scratch, L, D_cap, mode, host_counts, (cudaEvent_t)event,
                         s->debug != 0, (cudaStream_t)stream);
}

int gsb_render_fwd(const GsbSettings* s, int P, void* saved, long long D_cap, float* out_color,
                   float* out_depth, float* out_alpha, void* stream) {
  if (!settings_ok(s) || P < 0 || !saved || !out_color || !out_depth || !out_alpha) return GSB_E_INVALID;
  GsbLayout L;
  int rc = layout(P, s->image_height, s->image_width, D_cap, &L);
  if (rc) return rc;
  ProfScope ps(GSB_STAGE_RENDER_FWD, (cudaStream_t)stream);
  if (g_blend_variant.load() == 1)
    return launch_standin_fwd(make_view(s), at<Geom>(saved, L.off_geom), at<uint32_t>(saved, L.off_point_list),
                              at<uint2>(saved, L.off_ranges), out_color, out_depth, out_alpha,
                              at<uint32_t>(saved, L.off_n_contrib), at<float>(saved, L.off_final_T), s->debug != 0,
                              (cudaStream_t)stream);
  const bool record = s->forward_only == 0;
  return launch_render_fwd(make_view(s), P, at<Geom>(saved, L.off_geom), at<uint32_t>(saved, L.off_point_list),
                           at<uint2>(saved, L.off_ranges), at<uint32_t>(saved, L.off_tile_order), out_color,
                           out_depth, out_alpha,
                           at<uint32_t>(saved, L.off_n_contrib), at<float>(saved, L.off_final_T),
                           record ? at<uint2>(saved, L.off_hits) : nullptr,
                           record ? at<uint32_t>(saved, L.off_hit_count) : nullptr, g_fwd_variant.load(),
                           s->debug != 0, (cudaStream_t)stream);
}

int gsb_forward(const GsbSettings* s, int P, int K, const float* means3D, const float* scales,
                const float* rotations, const float* opacities, const float* shs,
                const float* colors_precomp, const float* cov3D_precomp, int32_t* radii_out,
                float* out_color, float* out_depth, float* out_alpha, void* saved, void* scratch,
                long long D_cap, int mode, uint32_t* host_counts, void* event, void* stream) {
  int rc = gsb_preprocess_fwd(s, P, K, means3D, scales, rotations, opacities, shs, colors_precomp,
                              cov3D_precomp, radii_out, saved, scratch, D_cap, stream);
  if (rc) return rc;
  rc = gsb_bin_sort(s, P, saved, scratch, D_cap, mode, host_counts, event, stream);
  if (rc) return rc;
  return gsb_render_fwd(s, P, saved, D_cap, out_color, out_depth, out_alpha, stream);
}

int gsb_read_counts(const void* saved, int P, int H, int W, long long D_cap, uint32_t* host_dst,
                    void* stream) {
  if (!saved || !host_dst) return GSB_E_INVALID;
  GsbLayout L;
  int rc = layout(P, H, W, D_cap, &L);
  if (rc) return rc;
  GSB_CUDA(cudaMemcpyAsync(host_dst, at<uint32_t>(saved, L.off_counts), 8 * sizeof(uint32_t),
                           cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return GSB_OK;
}

int gsb_render_bwd(const GsbSettings* s, int P, const void* saved, void* scratch, long long D_cap,
                   const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, void* stream) {
  if (!settings_ok(s) || P < 0 || !saved || !scratch || !dL_dcolor || !dL_ddepth || !dL_dalpha)
    return GSB_E_INVALID;
  GsbLayout L;
  int rc = layout(P, s->image_height, s->image_width, D_cap, &L);
  if (rc) return rc;
  ProfScope ps(GSB_STAGE_RENDER_BWD, (cudaStream_t)stream);
  if (g_blend_variant.load() == 1)
    return launch_standin_bwd(make_view(s), P, at<Geom>(saved, L.off_geom), at<uint32_t>(saved, L.off_point_list),
                              at<uint2>(saved, L.off_ranges), at<uint32_t>(saved, L.off_n_contrib),
                              at<float>(saved, L.off_final_T), dL_dcolor, dL_ddepth, dL_dalpha,
                              at<GGrad>(scratch, L.off_ggrad), s->debug != 0, (cudaStream_t)stream);
  if (s->forward_only) return GSB_E_INVALID;       // that forward kept nothing for a backward pass
  // variants: 0 transposed replay (default), 2 per-hit replay + butterfly, 3 rescan + butterfly (the round-1
  // kernel, record-free), 4 rescan + packed reduction
  const int variant = g_blend_variant.load();
  return launch_render_bwd(make_view(s), P, at<Geom>(saved, L.off_geom), at<uint32_t>(saved, L.off_point_list),
                           at<uint2>(saved, L.off_ranges), at<uint32_t>(saved, L.off_tile_order),
                           at<uint32_t>(saved, L.off_n_contrib), at<float>(saved, L.off_final_T),
                           at<uint2>(saved, L.off_hits), at<uint32_t>(saved, L.off_hit_count), dL_dcolor, dL_ddepth,
                           dL_dalpha, at<GGrad>(scratch, L.off_ggrad), variant == 0 ? 0 : (variant == 2 ? 1 : 2),
                           variant == 4, s->debug != 0, (cudaStream_t)stream);
}

int gsb_preprocess_bwd_views(int V, const GsbSettings* const* settings, int P, int K, const float* means3D,
                             const float* scales, const float* rotations, const float* opacities,
                             const float* shs, const float* colors_precomp, const float* cov3D_precomp,
                             const int32_t* const* radii, const void* const* saved, const void* const* scratch,
                             const long long* D_cap, float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs,
                             float* dL_dcolors, float* dL_dopacities, float* dL_dscales, float* dL_drotations,
                             float* dL_dcov3D, int accumulate, void* stream) {
  (void)opacities;
  if (V < 1 || V > GSB_MAX_VIEWS || P < 0 || !settings || !radii || !saved || !scratch || !D_cap)
    return GSB_E_INVALID;
  if (P == 0) return GSB_OK;
  if (!means3D || !dL_dmeans3D || !dL_dmeans2D || !dL_dopacities) return GSB_E_INVALID;
  if ((shs != nullptr) == (colors_precomp != nullptr)) return GSB_E_INVALID;
  if ((cov3D_precomp != nullptr) == (scales != nullptr && rotations != nullptr)) return GSB_E_INVALID;
  if (shs && !dL_dshs) return GSB_E_INVALID;
  if (colors_precomp && !dL_dcolors) return GSB_E_INVALID;
  if (cov3D_precomp && !dL_dcov3D) return GSB_E_INVALID;
  if (!cov3D_precomp && (!dL_dscales || !dL_drotations)) return GSB_E_INVALID;
  BwdBatch B;
  B.V = V;
  bool debug = false;
  for (int v = 0; v < V; ++v) {
    const GsbSettings* s = settings[v];
    if (!settings_ok(s) || !radii[v] || !saved[v] || !scratch[v]) return GSB_E_INVALID;
    if (shs && K < (s->sh_degree + 1) * (s->sh_degree + 1)) return GSB_E_INVALID;
    GsbLayout L;
    int rc = layout(P, s->image_height, s->image_width, D_cap[v], &L);
    if (rc) return rc;
    B.a[v].v = make_view(s);
    B.a[v].geom = at<Geom>(saved[v], L.off_geom);
    B.a[v].clamped = at<uint8_t>(saved[v], L.off_clamped);
    B.a[v].ggrad = at<GGrad>(scratch[v], L.off_ggrad);
    B.a[v].radii = radii[v];
    debug = debug || s->debug != 0;
  }
  ProfScope ps(GSB_STAGE_PREPROCESS_BWD, (cudaStream_t)stream);
  return launch_preprocess_bwd(B, P, K, means3D, scales, rotations, shs, colors_precomp, cov3D_precomp,
                               dL_dmeans3D, dL_dmeans2D, dL_dshs, dL_dcolors, dL_dopacities, dL_dscales,
                               dL_drotations, dL_dcov3D, accumulate, debug, (cudaStream_t)stream);
}

int gsb_preprocess_bwd(const GsbSettings* s, int P, int K, const float* means3D, const float* scales,
                       const float* rotations, const float* opacities, const float* shs,
                       const float* colors_precomp, const float* cov3D_precomp, const int32_t* radii,
                       const void* saved, const void* scratch, long long D_cap, float* dL_dmeans3D,
                       float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors, float* dL_dopacities,
                       float* dL_dscales, float* dL_drotations, float* dL_dcov3D, int accumulate,
                       void* stream) {
  return gsb_preprocess_bwd_views(1, &s, P, K, means3D, scales, rotations, opacities, shs, colors_precomp,
                                  cov3D_precomp, &radii, &saved, &scratch, &D_cap, dL_dmeans3D, dL_dmeans2D,
                                  dL_dshs, dL_dcolors, dL_dopacities, dL_dscales, dL_drotations, dL_dcov3D,
                                  accumulate, stream);
}

int gsb_backward(const GsbSettings* s, int P, int K, const float* means3D, const float* scales,
                 const float* rotations, const float* opacities, const float* shs,
                 const float* colors_precomp, const float* cov3D_precomp, const int32_t* radii,
                 const void* saved, void* scratch, long long D_cap, const float* dL_dcolor,
                 const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D, float* dL_dmeans2D,
                 float* dL_dshs, float* dL_dcolors, float* dL_dopacities, float* dL_dscales,
                 float* dL_drotations, float* dL_dcov3D, int accumulate, void* stream) {
  int rc = gsb_render_bwd(s, P, saved, scratch, D_cap, dL_dcolor, dL_ddepth, dL_dalpha, stream);
  if (rc) return rc;
  return gsb_preprocess_bwd(s, P, K, means3D, scales, rotations, opacities, shs, colors_precomp, cov3D_precomp,
                            radii, saved, scratch, D_cap, dL_dmeans3D, dL_dmeans2D, dL_dshs, dL_dcolors,
                            dL_dopacities, dL_dscales, dL_drotations, dL_dcov3D, accumulate, stream);
}

int gsb_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
  (void)projmatrix;
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return GSB_E_INVALID;
  return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

int gsb_exchange_config(int world, int rank, long long rows_per_rank, const void* const* peer_bases) {
  return set_exchange_peers(world, rank, rows_per_rank, peer_bases);
}

int gsb_exchange_set_aux(int32_t* radii_max, const float* scalar_in, float* scalar_out) {
  return set_exchange_aux(radii_max, scalar_in, scalar_out);
}

int gsb_exchange_gather(const float* local_base, float* multicast_base, int n_segments, const long long* offset_floats,
                        const long long* count_floats, void* stream) {
  if (n_segments < 0 || (n_segments > 0 && (!offset_floats || !count_floats))) return GSB_E_INVALID;
  return launch_exchange_gather(local_base, multicast_base, n_segments, offset_floats, count_floats,
                                (cudaStream_t)stream);
}

size_t gsb_mask_index_tmp_bytes(long long n) { return mask_index_tmp_bytes(n); }

int gsb_mask_to_index(long long n, const uint8_t* mask, int64_t* index, uint32_t* count, void* tmp, void* stream) {
  if (n < 0 || !count || (n > 0 && (!mask || !index || !tmp))) return GSB_E_INVALID;
  return launch_mask_to_index(n, mask, index, count, tmp, (cudaStream_t)stream);
}

int gsb_gather_rows(int n_tensors, const float* const* src, float* const* dst, const int* widths, long long n_rows,
                    const int64_t* index, long long dst_row0, void* stream) {
  if (n_tensors < 0 || n_rows < 0 || dst_row0 < 0) return GSB_E_INVALID;
  if (n_tensors > 0 && (!src || !dst || !widths)) return GSB_E_INVALID;
  return launch_gather_rows(n_tensors, src, dst, widths, n_rows, index, dst_row0, (cudaStream_t)stream);
}

size_t gsb_knn_scratch_bytes(long long P) { return knn_scratch_bytes(P); }

int gsb_knn_dist2(long long P, const float* points, float* mean_dist2, void* scratch, size_t scratch_bytes,
                  void* stream) {
  if (P < 0) return GSB_E_INVALID;
  return launch_knn_dist2(P, points, mean_dist2, scratch, scratch_bytes, false, (cudaStream_t)stream);
}

int gsb_debug_sorted_keys(int P, int H, int W, const void* saved, const void* scratch, long long D_cap,
                          uint64_t* keys_out, void* stream) {
  if (!saved || !scratch || !keys_out) return GSB_E_INVALID;
  GsbLayout L;
  int rc = layout(P, H, W, D_cap, &L);
  if (rc) return rc;
  View v;
  memset(&v, 0, sizeof(v));
  v.H = H; v.W = W; v.gx = (W + TILE_X - 1) / TILE_X; v.gy = (H + TILE_Y - 1) / TILE_Y;
  return launch_debug_sorted_keys(v, P, saved, scratch, L, D_cap, keys_out, (cudaStream_t)stream);
}

size_t gsb_radix_tmp_bytes(long long n, int key_bytes) {
  if (n < 0) return 0;
  const size_t nz = (size_t)n;
  // ping-pong key+value buffers (A and B) followed by the histogram block
  return 2 * (align_up(nz * (size_t)key_bytes) + align_up(nz * 4)) + align_up(radix_tmp_bytes(n));
}

static int radix_entry(long long n, const void* keys_in, const uint32_t* vals_in, void* keys_out,
                       uint32_t* vals_out, int end_bit, void* tmp, void* stream, int key_bytes) {
  if (n < 0 || end_bit < 0 || end_bit > key_bytes * 8) return GSB_E_INVALID;
  if (n == 0) return GSB_OK;
  if (!keys_in || !vals_in || !keys_out || !vals_out || !tmp) return GSB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nz = (size_t)n, kb = align_up(nz * key_bytes), vb = align_up(nz * 4);
  char* base = static_cast<char*>(tmp);
  void* kA = base; uint32_t* vA = reinterpret_cast<uint32_t*>(base + kb);
  void* kB = base + kb + vb; uint32_t* vB = reinterpret_cast<uint32_t*>(base + 2 * kb + vb);
  void* hist = base + 2 * (kb + vb);
  const int passes = radix_num_passes(end_bit);
  // make the final buffer the caller's output
  if (radix_result_in_A(passes)) { kA = keys_out; vA = vals_out; } else { kB = keys_out; vB = vals_out; }
  if (passes == 0) {
    GSB_CUDA(cudaMemcpyAsync(keys_out, keys_in, nz * key_bytes, cudaMemcpyDeviceToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(vals_out, vals_in, nz * 4, cudaMemcpyDeviceToDevice, st));
    return GSB_OK;
  }
  if (key_bytes == 4)
    return radix_sort_pairs<uint32_t>(n, nullptr, (const uint32_t*)keys_in, vals_in, (uint32_t*)kA, vA,
                                      (uint32_t*)kB, vB, end_bit, false, false, hist, false, st);
  return radix_sort_pairs<uint64_t>(n, nullptr, (const uint64_t*)keys_in, vals_in, (uint64_t*)kA, vA,
                                    (uint64_t*)kB, vB, end_bit, false, false, hist, false, st);
}

int gsb_radix_sort_pairs_u32(long long n, const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                             uint32_t* vals_out, int end_bit, void* tmp, void* stream) {
  return radix_entry(n, keys_in, vals_in, keys_out, vals_out, end_bit, tmp, stream, 4);
}

int gsb_radix_sort_pairs_u64(long long n, const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                             uint32_t* vals_out, int end_bit, void* tmp, void* stream) {
  return radix_entry(n, keys_in, vals_in, keys_out, vals_out, end_bit, tmp, stream, 8);
}

}  // extern "C"
