/* gsb.h — C ABI of the B200-native Gaussian-splatting rasterizer (libgsb.so).
 *
 * This is the drop-in boundary for the one native operator GaussianIP calls:
 * `diff_gaussian_rasterization._C` (external ashawkey depth+alpha fork; NOT vendored under
 * /root/reference).  The reference binds that operator from Python at
 *   gaussiansplatting/gaussian_renderer/__init__.py:14,36-51,85-93   (render)
 *   gaussiansplatting/gaussian_renderer/__init__.py:124-139,175-183  (render_with_smaller_scale)
 *   gaussiansplatting/gaussian_renderer/__init__.py:213-228,240-248  (render_deformed)
 *   gs_renderer.py:10-13,943-958,992-1001                            (Renderer.render)
 * Each entry point below names the `_C` function / stage it replaces.
 *
 * Conventions: plain C, raw DEVICE pointers and sizes, no torch types, `void* stream` is a
 * cudaStream_t and always the last argument, every function returns 0 on success or a
 * negative GSB_E_* code (never throws, never allocates or frees device memory, never
 * synchronises the device unless `debug` is set in the settings).  All float tensors are
 * contiguous fp32.  There is no CPU fallback.
 */
#ifndef GSB_H
#define GSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_ABI_VERSION 3

enum {
  GSB_OK = 0,
  GSB_E_INVALID = -1,     /* bad argument (null pointer, negative size, sh_degree > 3 ...) */
  GSB_E_CUDA = -2,        /* a CUDA runtime call / kernel launch failed (see gsb_last_cuda_error) */
  GSB_E_CAPACITY = -3,    /* D (num_rendered) exceeded D_cap; re-run with a larger workspace */
  GSB_E_UNSUPPORTED = -4
};

/* Additive: the activations GaussianModel applies before every render call
 * (gaussian_model.py:84-107: opacity = sigmoid(_opacity), scaling = exp(_scaling), rotation = normalize(_rotation))
 * can be evaluated inside the per-Gaussian kernels instead of by ~15 torch kernels per step: with a bit set, the
 * corresponding input holds the raw parameter, the forward activates it on load, and the backward returns the
 * gradient with respect to the RAW parameter (chain rule applied after the sum over views). */
#define GSB_RAW_OPACITY 1    /* opacities are logits */
#define GSB_RAW_SCALE 2      /* scales are log-scales */
#define GSB_RAW_ROTATION 4   /* rotations are unnormalised quaternions */

/* The 12 fields of GaussianRasterizationSettings
 * (gaussiansplatting/gaussian_renderer/__init__.py:36-49) that reach the kernels. */
typedef struct GsbSettings {
  int32_t image_height, image_width;
  float tanfovx, tanfovy;
  float scale_modifier;
  int32_t sh_degree;         /* active degree, 0..3 */
  int32_t prefiltered;       /* !=0: the caller asserts every point passes the near-plane test; a point that does not
                                raises counts[5] (the external operator traps the device there); call sites pass False */
  int32_t debug;             /* !=0: synchronise + check after every kernel */
  int32_t raw_inputs;        /* GSB_RAW_* bits: which inputs are the model's RAW parameters (0 = the reference's contract) */
  int32_t forward_only;      /* !=0: no backward pass will use `saved` (inference / playback): the forward blend skips
                                the hit records and `saved` may be saved_bytes_forward_only bytes long */
  const float* bg;           /* [3]  device */
  const float* viewmatrix;   /* [16] device, column-major world->view (row-vector convention) */
  const float* projmatrix;   /* [16] device, column-major full projection */
  const float* campos;       /* [3]  device */
  const float* tanfov_dev;   /* optional [2] device = {tanfovx, tanfovy}; when non-NULL the kernels read the intrinsics
                                from here instead of the two by-value fields above (which must still be > 0), so a
                                CUDA graph captured over these calls can be replayed with new cameras */
} GsbSettings;

/* Byte offsets of every sub-buffer inside the two blocks the HOST allocates per call
 * (the library never allocates).  `saved` lives until the backward pass of the same view;
 * `scratch` is transient within one call and may be shared by all calls on one stream. */
typedef struct GsbLayout {
  size_t saved_bytes;
  size_t off_geom;        /* P x 48 B  GsbGeom record {x,y,extx,exty | qa,qb,qc (conic scaled by -0.5log2e,-log2e,-0.5log2e),opacity | depth,r,g,b} */
  size_t off_clamped;     /* P x u8    SH clamp mask (bit c: channel c clamped at 0) */
  size_t off_counts;      /* 8 x u32   [0]=D (num_rendered) [1]=overflow flag [4]=binning mode [5]=prefiltered violated */
  size_t off_point_list;  /* D_cap x u32  Gaussian index per sorted instance */
  size_t off_ranges;      /* T x {u32 start,u32 end} */
  size_t off_n_contrib;   /* H*W x u32 */
  size_t off_final_T;     /* H*W x f32 */
  size_t off_tile_order;  /* T x u32  tile ids, heaviest first (CTA scheduling order of the blend kernels) */
  size_t off_hit_count;   /* T x 8 x u32  hit records written by each of the 8 warps of a tile's CTA (forward blend) */
  size_t off_hits;        /* 8 x D_cap x {u32 Gaussian id, u32 lane mask}: per (tile, warp) the Gaussians the warp
                             BLENDED, in list order, with the mask of its 32 pixels that blended them; warp w of the
                             tile with range [s, e) owns records [8 s + w (e - s), +(e - s)).  Written by the forward
                             blend, replayed back to front by the backward blend (no culling / alpha tests there).
                             LAST in the block: a forward that no backward follows (GsbSettings.forward_only) needs
                             only the first saved_bytes_forward_only bytes */
  size_t saved_bytes_forward_only;
  size_t scratch_bytes;
  size_t off_rect;        /* P x {u16 minx,miny,maxx,maxy} */
  size_t off_tiles;       /* P x u32  tiles_touched (written by the forward; the emission derives counts from off_rect) */
  size_t off_dkeys0;               /* P x u32  depth keys (0xFFFFFFFF when culled), kept intact */
  size_t off_dkeys1, off_dkeys2;   /* P x u32  depth-sort ping/pong keys */
  size_t off_didx0, off_didx1;     /* P x u32  depth-sort ping/pong Gaussian ids */
  size_t off_offsets;     /* P x u32  reserved (the exclusive scan lives in registers of the emission kernel) */
  size_t off_blocksums;   /* scan chain of the emission kernel (u64 words) */
  size_t off_hist;        /* radix per-block digit histograms + digit totals */
  size_t off_tkeys0, off_tkeys1;   /* D_cap x u32 tile ids ping/pong */
  size_t off_tvals_alt;   /* D_cap x u32 */
  size_t off_keys64_0, off_keys64_1; /* D_cap x u64 (tile<<32|depth) keys, flat-64 mode + debug */
  size_t off_ggrad;       /* P x 48 B  per-Gaussian gradient record written by render_bwd */
} GsbLayout;

int gsb_abi_version(void);
const char* gsb_strerror(int code);
/* Text of the last CUDA error seen by this thread ("" if none). */
const char* gsb_last_cuda_error(void);

/* Workspace sizing (replaces the resize callbacks `_C.rasterize_gaussians` receives for
 * geomBuffer / binningBuffer / imgBuffer).  D_cap is the instance capacity. */
int gsb_layout(int P, int H, int W, long long D_cap, GsbLayout* out);

/* Binning modes for gsb_bin_sort / gsb_forward. */
#define GSB_BIN_TWO_LEVEL 0   /* 32-bit depth sort of P Gaussians + stable tile partition of D instances */
#define GSB_BIN_FLAT64    1   /* duplicateWithKeys + 64-bit (tile|depth) LSD radix sort (reference structure) */

/* Stage 1 — replaces preprocessCUDA + InclusiveSum inside `_C.rasterize_gaussians`.
 * Exactly one of shs / colors_precomp, and exactly one of (scales, rotations) /
 * cov3D_precomp, is non-null.  K = coefficients per Gaussian present in `shs` ([P,K,3]). */
int gsb_preprocess_fwd(const GsbSettings* s, int P, int K,
                       const float* means3D, const float* scales, const float* rotations,
                       const float* opacities, const float* shs, const float* colors_precomp,
                       const float* cov3D_precomp,
                       int32_t* radii_out, void* saved, void* scratch, long long D_cap,
                       void* stream);

/* Stage 2 — replaces duplicateWithKeys + cub::DeviceRadixSort::SortPairs +
 * identifyTileRanges.  Result: saved.point_list, saved.ranges, saved.counts.
 * Where the reference synchronises the device to read num_rendered, this library copies
 * saved.counts (8 x u32) to `host_counts` (pinned host memory, may be null) right after the
 * emission kernel (which scans the instance counts and publishes D) and records `event` (a
 * cudaEvent_t, may be null) behind that copy: the host can wait on the event alone, while the
 * sort and blend kernels are already queued.  Must follow gsb_preprocess_fwd of the same view on
 * the same stream and scratch block (it consumes the tile rectangles, depth keys and the
 * prefiltered flag that call left there). */
int gsb_bin_sort(const GsbSettings* s, int P, void* saved, void* scratch, long long D_cap,
                 int mode, uint32_t* host_counts, void* event, void* stream);

/* Stage 3 — replaces renderCUDA (forward).  Outputs [3,H,W], [1,H,W], [1,H,W]. */
int gsb_render_fwd(const GsbSettings* s, int P, void* saved, long long D_cap,
                   float* out_color, float* out_depth, float* out_alpha, void* stream);

/* Stages 1-3 in one call — replaces `_C.rasterize_gaussians`. */
int gsb_forward(const GsbSettings* s, int P, int K,
                const float* means3D, const float* scales, const float* rotations,
                const float* opacities, const float* shs, const float* colors_precomp,
                const float* cov3D_precomp,
                int32_t* radii_out, float* out_color, float* out_depth, float* out_alpha,
                void* saved, void* scratch, long long D_cap, int mode,
                uint32_t* host_counts, void* event, void* stream);

/* Asynchronously copy saved.counts (8 x u32) to `host_dst` (pinned host memory). */
int gsb_read_counts(const void* saved, int P, int H, int W, long long D_cap,
                    uint32_t* host_dst, void* stream);

/* Backward stage 1 — replaces renderCUDA (backward): per-pixel reverse blend, gradients
 * reduced across each warp before one atomic per (warp, Gaussian, component).  Default: replays the hit
 * records the forward blend wrote into `saved` (see GsbLayout.off_hits); gsb_set_blend_variant selects the
 * record-free kernel that re-walks the tile lists instead. */
int gsb_render_bwd(const GsbSettings* s, int P, const void* saved, void* scratch, long long D_cap,
                   const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                   void* stream);

/* Backward stage 2 — replaces computeCov2DCUDA + preprocessCUDA (backward).  Output
 * pointers may be null when the matching input was not given.  accumulate != 0 adds into
 * the outputs (beta = 1) instead of overwriting, so several views can share one gradient
 * bucket that is then all-reduced once. */
int gsb_preprocess_bwd(const GsbSettings* s, int P, int K,
                       const float* means3D, const float* scales, const float* rotations,
                       const float* opacities, const float* shs, const float* colors_precomp,
                       const float* cov3D_precomp, const int32_t* radii,
                       const void* saved, const void* scratch, long long D_cap,
                       float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs,
                       float* dL_dcolors, float* dL_dopacities, float* dL_dscales,
                       float* dL_drotations, float* dL_dcov3D, int accumulate, void* stream);

/* Batched form of gsb_preprocess_bwd over V <= GSB_MAX_VIEWS views of the SAME Gaussians (the
 * views of one optimisation step): the per-view contributions are summed in registers and each
 * gradient tensor is written once.  settings[v], radii[v], saved[v], scratch[v], D_cap[v] describe
 * view v exactly as in gsb_preprocess_bwd (all views share P, K, resolution may differ). */
#define GSB_MAX_VIEWS 8
int gsb_preprocess_bwd_views(int V, const GsbSettings* const* settings, int P, int K,
                             const float* means3D, const float* scales, const float* rotations,
                             const float* opacities, const float* shs, const float* colors_precomp,
                             const float* cov3D_precomp, const int32_t* const* radii,
                             const void* const* saved, const void* const* scratch, const long long* D_cap,
                             float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs,
                             float* dL_dcolors, float* dL_dopacities, float* dL_dscales,
                             float* dL_drotations, float* dL_dcov3D, int accumulate, void* stream);

/* Both backward stages — replaces `_C.rasterize_gaussians_backward`. */
int gsb_backward(const GsbSettings* s, int P, int K,
                 const float* means3D, const float* scales, const float* rotations,
                 const float* opacities, const float* shs, const float* colors_precomp,
                 const float* cov3D_precomp, const int32_t* radii,
                 const void* saved, void* scratch, long long D_cap,
                 const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                 float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs,
                 float* dL_dcolors, float* dL_dopacities, float* dL_dscales,
                 float* dL_drotations, float* dL_dcov3D, int accumulate, void* stream);

/* Replaces `_C.mark_visible` (GaussianRasterizer.markVisible): present[i] = view z > 0.2. */
int gsb_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* Measurement: per-stage device timing and launch counting, live, inside the normal call path.
 * While enabled the library records a cudaEvent on the caller's stream at every stage
 * boundary of gsb_forward / gsb_backward (about 1 us of host time each). */
enum {
  GSB_STAGE_PREPROCESS_FWD = 0, GSB_STAGE_DEPTH_SORT = 1, GSB_STAGE_SCAN_EMIT = 2,
  GSB_STAGE_TILE_SORT = 3, GSB_STAGE_RANGES = 4, GSB_STAGE_RENDER_FWD = 5,
  GSB_STAGE_RENDER_BWD = 6, GSB_STAGE_PREPROCESS_BWD = 7, GSB_NUM_STAGES = 8
};
int gsb_profile_enable(int on);                   /* resets the accumulators */
/* Synchronises the recorded events; ms_out[GSB_NUM_STAGES] = summed device ms per stage,
 * calls_out[GSB_NUM_STAGES] = number of times each stage ran since gsb_profile_enable(1). */
int gsb_profile_read(float* ms_out, int* calls_out);
/* Kernels launched by this library (host-side counter, all threads) since process start. */
long long gsb_launch_count(void);

/* the step after the hot path — torch.optim.Adam over up to GSB_ADAM_MAX_GROUPS
 * parameter tensors (gaussian_model.py:138-159: six groups, own lr each, eps 1e-15) plus the
 * densification statistics (GaussianIP.py:456-457, gaussian_model.py:420-422) in ONE kernel.
 * `step` is the 1-based Adam step count; grad_scale multiplies every gradient first (AMP unscale, 1 = off);
 * stats_n = 0 skips the statistics.  Pointer arrays are HOST arrays of DEVICE pointers. */
#define GSB_ADAM_MAX_GROUPS 8
int gsb_adam_step(int n_groups, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const long long* counts, const float* lrs, double beta1, double beta2,
                  double eps, long long step, float grad_scale, long long stats_n, const float* viewspace_grad,
                  const int32_t* radii, float* xyz_gradient_accum, float* denom, float* max_radii2D,
                  void* stream);
/* Same launch with one 1-based step count PER GROUP (steps[n_groups], host array): torch.optim.Adam keeps `step`
 * per parameter, so a group whose gradient was None on earlier steps has its own bias correction. */
int gsb_adam_step_groups(int n_groups, float* const* params, const float* const* grads, float* const* exp_avg,
                         float* const* exp_avg_sq, const long long* counts, const float* lrs, double beta1,
                         double beta2, double eps, const long long* steps, float grad_scale, long long stats_n,
                         const float* viewspace_grad, const int32_t* radii, float* xyz_gradient_accum,
                         float* denom, float* max_radii2D, void* stream);

/* replaces `simple_knn._C.distCUDA2` (simple-knn/spatial.cu:15-26 -> SimpleKNN::knn,
 * simple_knn.cu:186-221; callers gaussian_model.py:123, gs_renderer.py:387): mean_dist2[i] = mean of the
 * squared distances from point i to its 3 nearest other points, in the reference's own fp32 rounding
 * (bit-identical output).  points: [P,3] fp32 device, mean_dist2: [P] fp32 device, scratch: device,
 * >= gsb_knn_scratch_bytes(P).  Asynchronous on `stream`; no host read-back. */
size_t gsb_knn_scratch_bytes(long long P);
int gsb_knn_dist2(long long P, const float* points, float* mean_dist2, void* scratch,
                  size_t scratch_bytes, void* stream);

/* gradient exchange fused into gsb_preprocess_bwd_views (multi-GPU, one process per GPU).
 * The gradient outputs live in a SYMMETRIC buffer: same layout on every rank, every rank's copy mapped into
 * every process (peer memory over NVLink) plus one NVLS multicast mapping of all copies.  `accumulate` selects:
 *   2  outputs are MULTICAST addresses: the kernel's epilogue multimem.red-adds every row into all copies
 *      (one kernel, no second phase; every GPU receives N gradients -> best for 2 ranks);
 *   3  outputs are addresses in the LOCAL copy: rows are owned by ranks in blocks of rows_per_rank; the epilogue
 *      red-adds each row into the owner's copy only, then gsb_exchange_gather lets every owner multicast its
 *      reduced block to all copies (every GPU receives ~2 gradients for any N).
 * Copies must be zero (mode 2: everywhere; mode 3: at least the owned block) before any rank's kernel starts and
 * ranks must meet at a barrier between the phases — the caller's job (gaussianip_b200/exchange.py).
 * gsb_exchange_config: process-wide table for mode 3; peer_bases[r] = base of rank r's copy as mapped in THIS
 * process (peer_bases[rank] = the local copy); rows_per_rank must be a multiple of 32. */
#define GSB_MAX_RANKS 16
#define GSB_EXCHANGE_MAX_SEGMENTS 8
int gsb_exchange_config(int world, int rank, long long rows_per_rank, const void* const* peer_bases);
/* Extras that ride in the next mode-2 / mode-3 launches of gsb_preprocess_bwd_views (process-wide, until changed):
 *   radii_max   [P] int32 inside the symmetric buffer (multicast address in mode 2, local copy in mode 3; zero before
 *               the step): receives the MAX over ranks of each Gaussian's largest radius over the launch's views;
 *   scalar_in   one float on this device (e.g. the step's loss), added into scalar_out of every copy: scalar_out is
 *               one float inside the symmetric buffer (multicast address in mode 2; in mode 3 the local address of
 *               a slot owned by rank 0, which gsb_exchange_gather then distributes).
 * NULL disables an extra.  With both, a view-sharded step needs no NCCL call at all. */
int gsb_exchange_set_aux(int32_t* radii_max, const float* scalar_in, float* scalar_out);
/* For each segment s < n_segments: floats [offset[s], offset[s] + count[s]) of the local copy are stored to the
 * same offsets of every copy through multicast_base (offsets multiples of 4 floats). */
int gsb_exchange_gather(const float* local_base, float* multicast_base, int n_segments,
                        const long long* offset_floats, const long long* count_floats, void* stream);

/* data movement of densify / clone / split / prune (gaussian_model.py:281-418).
 * gsb_mask_to_index: stable stream compaction — index[0..count) = positions of the non-zero mask bytes in
 *   ascending order (what `tensor[mask]` / torch.nonzero use), count[0] = how many; device-side, asynchronous.
 *   tmp: device, >= gsb_mask_index_tmp_bytes(n).
 * gsb_gather_rows: ONE launch for up to GSB_GATHER_MAX_TENSORS row-major fp32 tensors sharing the row index:
 *   dst[t][dst_row0 + r][:] = src[t][index ? index[r] : r][:] for r < n_rows, or zeros where src[t] == NULL
 *   (new points' Adam moments, cat_tensors_to_optimizer :334-335).  src/dst/widths are HOST arrays; widths in floats. */
#define GSB_GATHER_MAX_TENSORS 24
size_t gsb_mask_index_tmp_bytes(long long n);
int gsb_mask_to_index(long long n, const uint8_t* mask, int64_t* index, uint32_t* count, void* tmp, void* stream);
int gsb_gather_rows(int n_tensors, const float* const* src, float* const* dst, const int* widths,
                    long long n_rows, const int64_t* index, long long dst_row0, void* stream);

/* Blend-kernel variant used by gsb_render_fwd / gsb_render_bwd (process-wide, default 0):
 *   0  native kernels (the product path): the forward blend records per-warp hit lists, the backward blend replays
 *      them in two transposed phases (lane = pixel: sequential part; lane = record: the ten sums);
 *   1  reference-STRUCTURE stand-in (csrc/standin.cu): thread per pixel, CTA-synchronous batches, no
 *      culling, per-pixel global atomics — for measurement context and as a GPU cross-check only;
 *   2  native forward + replay backward that takes one hit per half-warp at a time and reduces with a shuffle
 *      butterfly (cross-check of 0);
 *   3  native forward + the record-free backward that re-walks the tile lists with per-warp culling (the
 *      round-1 kernel; cross-check of 0);
 *   4  as 3 with a packed shared-memory reduction;
 *   10 / 11 / 12  select the FORWARD blend only (backward choice unchanged): 10 per-hit blend staged with cp.async
 *      (default), 11 transposed two-phase blend, 12 per-hit blend staged with the sm_100 TMA row gather
 *      (cp.async.bulk.tensor ... tile::gather4; measured alternatives, all parity-green). */
int gsb_set_blend_variant(int variant);

/* Test / measurement helpers. */
/* Materialise the sorted 64-bit (tile<<32 | depth bits) key of every instance. */
int gsb_debug_sorted_keys(int P, int H, int W, const void* saved, const void* scratch,
                          long long D_cap, uint64_t* keys_out, void* stream);
/* Stable LSD radix sort of n (key,value) pairs on bits [0,end_bit); result in keys_out/vals_out.
 * `tmp` must hold gsb_radix_tmp_bytes(n, key_bytes). */
size_t gsb_radix_tmp_bytes(long long n, int key_bytes);
int gsb_radix_sort_pairs_u32(long long n, const uint32_t* keys_in, const uint32_t* vals_in,
                             uint32_t* keys_out, uint32_t* vals_out, int end_bit, void* tmp,
                             void* stream);
int gsb_radix_sort_pairs_u64(long long n, const uint64_t* keys_in, const uint32_t* vals_in,
                             uint64_t* keys_out, uint32_t* vals_out, int end_bit, void* tmp,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSB_H */
