// Stage 1: per-Gaussian projection.  Replaces preprocessCUDA of the external
// diff_gaussian_rasterization operator (forward steps 1-10).
//
// THIS FILE IS COMPILED WITH -fmad=false.  Every fp32 operation below is separately rounded
// and written in the same order as oracle/splat_torch.py::preprocess, so depth bits, pixel
// centres, radii and tile rectangles are bit-identical to the CPU oracle.  The kernel is
// HBM-bound (reads 44+12K B, writes ~70 B per Gaussian), so giving up FMA contraction costs
// nothing measurable.
#include <atomic>

#include "gsb_common.cuh"

namespace gsb {

namespace {

constexpr int PRE_CTAS_PER_SM = 4;   // 57 registers x 256 threads

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// SH -> RGB, term order of gaussiansplatting/utils/sh_utils.py:57-99.  sh points at this
// Gaussian's [K,3] block; returns the unclamped colour of channel c.
__device__ __forceinline__ float sh_channel(int deg, const float* __restrict__ sh, int c, float x, float y,
                                            float z) {
  const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
  float r = C0 * sh[c];
  if (deg > 0) {
    r = r - (C1 * y) * sh[3 + c];
    r = r + (C1 * z) * sh[6 + c];
    r = r - (C1 * x) * sh[9 + c];
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      r = r + (1.0925484305920792f * xy) * sh[12 + c];
      r = r + (-1.0925484305920792f * yz) * sh[15 + c];
      r = r + (0.31539156525252005f * ((2.0f * zz - xx) - yy)) * sh[18 + c];
      r = r + (-1.0925484305920792f * xz) * sh[21 + c];
      r = r + (0.5462742152960396f * (xx - yy)) * sh[24 + c];
      if (deg > 2) {
        r = r + ((-0.5900435899266435f * y) * (3.0f * xx - yy)) * sh[27 + c];
        r = r + ((2.890611442640554f * xy) * z) * sh[30 + c];
        r = r + ((-0.4570457994644658f * y) * ((4.0f * zz - xx) - yy)) * sh[33 + c];
        r = r + ((0.3731763325901154f * z) * ((2.0f * zz - 3.0f * xx) - 3.0f * yy)) * sh[36 + c];
        r = r + ((-0.4570457994644658f * x) * ((4.0f * zz - xx) - yy)) * sh[39 + c];
        r = r + ((1.445305721320277f * z) * (xx - yy)) * sh[42 + c];
        r = r + ((-0.5900435899266435f * x) * (xx - 3.0f * yy)) * sh[45 + c];
      }
    }
  }
  return r;
}

__device__ __forceinline__ void
preprocess_one(const View& v, int i, int K, const float* sV, const float* sM, const float* sCam, const float* sK,
               const float* __restrict__ means3D, const float* __restrict__ scales, const float* __restrict__ rots,
               const float* __restrict__ opac, const float* __restrict__ shs, const float* __restrict__ colors,
               const float* __restrict__ cov3Dp, int32_t* __restrict__ radii, Geom* __restrict__ geom,
               uint8_t* __restrict__ clamped, ushort4* __restrict__ rect, uint32_t* __restrict__ tiles,
               uint32_t* __restrict__ dkeys, uint32_t* s_h0, uint32_t* __restrict__ flag_word) {
  // All of this Gaussian's inputs are requested up front, before the first of them is needed: the kernel is bound by
  // the latency of its loads (ncu r2d: long-scoreboard 6.5 of ~10 stall cycles per issue, DRAM at 21 % of peak),
  // and loading lazily inside the visibility branches made them four dependent round trips.
  const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
  const float o_in = opac[i];
  float in_s0 = 0.f, in_s1 = 0.f, in_s2 = 0.f, in_c0 = 0.f, in_c1 = 0.f, in_c2 = 0.f;
  float4 in_q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!cov3Dp) {
    in_s0 = scales[3 * i]; in_s1 = scales[3 * i + 1]; in_s2 = scales[3 * i + 2];
    in_q = reinterpret_cast<const float4*>(rots)[i];
  }
  if (colors) {
    in_c0 = colors[3 * i]; in_c1 = colors[3 * i + 1]; in_c2 = colors[3 * i + 2];
  } else if (v.sh_degree == 0) {
    const float* sh0 = shs + (size_t)i * K * 3;
    in_c0 = sh0[0]; in_c1 = sh0[1]; in_c2 = sh0[2];
  }
  const float tx = ((sV[0] * px + sV[4] * py) + sV[8] * pz) + sV[12];
  const float ty = ((sV[1] * px + sV[5] * py) + sV[9] * pz) + sV[13];
  const float tz = ((sV[2] * px + sV[6] * py) + sV[10] * pz) + sV[14];

  bool visible = tz > NEAR_Z;
  // prefiltered = the caller asserts every point is in front of the camera; the external operator traps the
  // device when one is not (in_frustum: "Point is filtered although prefiltered is set"), here the host raises
  if (!visible && v.prefiltered) *flag_word = 1u;
  int rad = 0;
  uint32_t ntiles = 0;
  ushort4 rc = make_ushort4(0, 0, 0, 0);     // tile rectangle; stays empty for a culled Gaussian
  if (visible) {
    const float hx = ((sM[0] * px + sM[4] * py) + sM[8] * pz) + sM[12];
    const float hy = ((sM[1] * px + sM[5] * py) + sM[9] * pz) + sM[13];
    const float hw = ((sM[3] * px + sM[7] * py) + sM[11] * pz) + sM[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    const float projx = hx * pw, projy = hy * pw;

    float S00, S01, S02, S11, S12, S22;
    if (cov3Dp) {
      const float* c = cov3Dp + 6 * (size_t)i;
      S00 = c[0]; S01 = c[1]; S02 = c[2]; S11 = c[3]; S12 = c[4]; S22 = c[5];
    } else {
      float s0 = in_s0, s1 = in_s1, s2 = in_s2;
      if (v.raw & GSB_RAW_SCALE) { s0 = expf(s0); s1 = expf(s1); s2 = expf(s2); }
      const float sx = v.scale_mod * s0, sy = v.scale_mod * s1, sz = v.scale_mod * s2;
      float4 q = in_q;
      if (v.raw & GSB_RAW_ROTATION) q = act_normalize(q);
      const float r = q.x, x = q.y, y = q.z, z = q.w;
      const float m00 = (1.0f - 2.0f * (y * y + z * z)) * sx, m01 = (2.0f * (x * y - r * z)) * sy,
                  m02 = (2.0f * (x * z + r * y)) * sz;
      const float m10 = (2.0f * (x * y + r * z)) * sx, m11 = (1.0f - 2.0f * (x * x + z * z)) * sy,
                  m12 = (2.0f * (y * z - r * x)) * sz;
      const float m20 = (2.0f * (x * z - r * y)) * sx, m21 = (2.0f * (y * z + r * x)) * sy,
                  m22 = (1.0f - 2.0f * (x * x + y * y)) * sz;
      S00 = (m00 * m00 + m01 * m01) + m02 * m02;
      S01 = (m00 * m10 + m01 * m11) + m02 * m12;
      S02 = (m00 * m20 + m01 * m21) + m02 * m22;
      S11 = (m10 * m10 + m11 * m11) + m12 * m12;
      S12 = (m10 * m20 + m11 * m21) + m12 * m22;
      S22 = (m20 * m20 + m21 * m21) + m22 * m22;
    }

    // EWA projection
    const float limx = 1.3f * sK[0], limy = 1.3f * sK[1];
    const float txtz = tx / tz, tytz = ty / tz;
    const float txc = fminf(limx, fmaxf(-limx, txtz)) * tz;
    const float tyc = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float tz2 = tz * tz;
    const float J00 = sK[2] / tz, J02 = -(sK[2] * txc) / tz2;
    const float J11 = sK[3] / tz, J12 = -(sK[3] * tyc) / tz2;
    // W(i,j) = sV[i + 4j]
    const float T00 = J00 * sV[0] + J02 * sV[2], T01 = J00 * sV[4] + J02 * sV[6],
                T02 = J00 * sV[8] + J02 * sV[10];
    const float T10 = J11 * sV[1] + J12 * sV[2], T11 = J11 * sV[5] + J12 * sV[6],
                T12 = J11 * sV[9] + J12 * sV[10];
    const float U00 = (T00 * S00 + T01 * S01) + T02 * S02, U01 = (T00 * S01 + T01 * S11) + T02 * S12,
                U02 = (T00 * S02 + T01 * S12) + T02 * S22;
    const float U10 = (T10 * S00 + T11 * S01) + T12 * S02, U11 = (T10 * S01 + T11 * S11) + T12 * S12,
                U12 = (T10 * S02 + T11 * S12) + T12 * S22;
    const float c00 = (U00 * T00 + U01 * T01) + U02 * T02;
    const float c01 = (U00 * T10 + U01 * T11) + U02 * T12;
    const float c11 = (U10 * T10 + U11 * T11) + U12 * T12;
    const float a = c00 + LOWPASS, b = c01, c = c11 + LOWPASS;
    const float det = a * c - b * b;
    if (det == 0.0f) visible = false;
    if (visible) {
      const float det_inv = 1.0f / det;
      const float cA = c * det_inv, cB = (-b) * det_inv, cC = a * det_inv;
      const float mid = 0.5f * (a + c);
      const float disc = sqrtf(fmaxf(mid * mid - det, 0.1f));
      const float lam = fmaxf(mid + disc, mid - disc);
      const float radf = ceilf(3.0f * sqrtf(lam));
      const float pixx = ((projx + 1.0f) * (float)v.W - 1.0f) * 0.5f;
      const float pixy = ((projy + 1.0f) * (float)v.H - 1.0f) * 0.5f;
      const float gxf = (float)v.gx, gyf = (float)v.gy;
      const int minx = (int)clampf(truncf((pixx - radf) / 16.0f), 0.0f, gxf);
      const int maxx = (int)clampf(truncf(((pixx + radf) + 15.0f) / 16.0f), 0.0f, gxf);
      const int miny = (int)clampf(truncf((pixy - radf) / 16.0f), 0.0f, gyf);
      const int maxy = (int)clampf(truncf(((pixy + radf) + 15.0f) / 16.0f), 0.0f, gyf);
      ntiles = (uint32_t)((maxx - minx) * (maxy - miny));
      if (ntiles == 0) visible = false;
      if (visible) {
        rad = (radf < 1073741824.0f) ? (int)radf : 1073741824;
        float r_, g_, b_;
        uint8_t cl = 0;
        if (colors) {
          r_ = in_c0; g_ = in_c1; b_ = in_c2;
        } else if (v.sh_degree == 0) {
          // degree 0: C0 * sh[c] + 0.5 (the first term of sh_channel), coefficients already loaded
          r_ = 0.28209479177387814f * in_c0 + 0.5f;
          g_ = 0.28209479177387814f * in_c1 + 0.5f;
          b_ = 0.28209479177387814f * in_c2 + 0.5f;
          if (r_ < 0.0f) { cl |= 1; r_ = 0.0f; }
          if (g_ < 0.0f) { cl |= 2; g_ = 0.0f; }
          if (b_ < 0.0f) { cl |= 4; b_ = 0.0f; }
        } else {
          const float dx = px - sCam[0], dy = py - sCam[1], dz = pz - sCam[2];
          const float ln = sqrtf((dx * dx + dy * dy) + dz * dz);
          const float x = dx / ln, y = dy / ln, z = dz / ln;
          const float* sh = shs + (size_t)i * K * 3;
          r_ = sh_channel(v.sh_degree, sh, 0, x, y, z) + 0.5f;
          g_ = sh_channel(v.sh_degree, sh, 1, x, y, z) + 0.5f;
          b_ = sh_channel(v.sh_degree, sh, 2, x, y, z) + 0.5f;
          if (r_ < 0.0f) { cl |= 1; r_ = 0.0f; }
          if (g_ < 0.0f) { cl |= 2; g_ = 0.0f; }
          if (b_ < 0.0f) { cl |= 4; b_ = 0.0f; }
        }
        const float o = (v.raw & GSB_RAW_OPACITY) ? act_sigmoid(o_in) : o_in;
        // Conservative half-extent (pixels) of the region where alpha = o*exp(power) can
        // reach 1/255: bounding box of {d : 0.5 d^T Q d <= ln(255 o)} is sqrt(2 tau cov_ii).
        // Used only to SKIP work that would be discarded anyway; padded against rounding.
        float ex, ey;
        const float tau = logf(255.0f * o);
        if (!(det > 0.0f) || !isfinite(a) || !isfinite(c) || !isfinite(tau)) {
          ex = ey = 3.0e38f;
        } else if (tau <= 0.0f) {
          ex = ey = -3.0e38f;
        } else {
          ex = sqrtf(2.0f * tau * a) * 1.002f + 0.05f;
          ey = sqrtf(2.0f * tau * c) * 1.002f + 0.05f;
        }
        float4* dst = reinterpret_cast<float4*>(geom + i);
        dst[0] = make_float4(pixx, pixy, ex, ey);
        dst[1] = make_float4(cA * CONIC_SCALE_AC, cB * CONIC_SCALE_B, cC * CONIC_SCALE_AC, o);
        dst[2] = make_float4(tz, r_, g_, b_);
        clamped[i] = cl;
        rc = make_ushort4((unsigned short)minx, (unsigned short)miny, (unsigned short)maxx, (unsigned short)maxy);
      }
    }
  }
  radii[i] = visible ? rad : 0;
  rect[i] = rc;                              // the emission kernel derives the instance count from it
  tiles[i] = visible ? ntiles : 0u;
  const uint32_t dk = visible ? __float_as_uint(tz) : 0xFFFFFFFFu;
  dkeys[i] = dk;
  // culled Gaussians all carry key 0xFFFFFFFF: aggregate them per warp before touching the counter
  const uint32_t culled = __ballot_sync(__activemask(), !visible);
  if (visible) atomicAdd(&s_h0[dk & 0xFFu], 1u);
  else if ((threadIdx.x & 31) == __ffs(culled) - 1) atomicAdd(&s_h0[255], (uint32_t)__popc(culled));
}

__global__ void __launch_bounds__(256)
preprocess_fwd_kernel(View v, int P, int K, const float* __restrict__ means3D,
                      const float* __restrict__ scales, const float* __restrict__ rots,
                      const float* __restrict__ opac, const float* __restrict__ shs,
                      const float* __restrict__ colors, const float* __restrict__ cov3Dp,
                      int32_t* __restrict__ radii, Geom* __restrict__ geom,
                      uint8_t* __restrict__ clamped, ushort4* __restrict__ rect,
                      uint32_t* __restrict__ tiles, uint32_t* __restrict__ dkeys,
                      uint32_t* __restrict__ ghist0, uint32_t* __restrict__ flag_word) {
  __shared__ float sV[16], sM[16], sCam[3], sK[4];   // sK: tanfovx, tanfovy, focal_x, focal_y
  __shared__ uint32_t s_h0[256];   // digit-0 histogram of the depth keys (first pass of the depth sort)
  s_h0[threadIdx.x] = 0;
  if (threadIdx.x < 16) { sV[threadIdx.x] = v.view[threadIdx.x]; sM[threadIdx.x] = v.proj[threadIdx.x]; }
  if (threadIdx.x < 3) sCam[threadIdx.x] = v.campos[threadIdx.x];
  if (threadIdx.x == 32) load_intrinsics(v, sK);
  __syncthreads();
  // Persistent CTAs (one wave, 4 per SM) stride over the 256-Gaussian chunks: the grid of round 1 (4 chunks per
  // CTA) was 1.65 waves at 1 M Gaussians, i.e. a third of the launch ran at partial occupancy (ncu r2:
  // sm__cycles_active / elapsed 0.78).  Few CTAs also means few flushes of the 256-bin histogram to the same
  // 256 global counters (one flush per 256 Gaussians cost +22 us of same-address L2 atomics).
#pragma unroll 1
  for (int c = blockIdx.x; c * 256 < P; c += gridDim.x) {
    const int i = c * 256 + threadIdx.x;
    if (i < P) preprocess_one(v, i, K, sV, sM, sCam, sK, means3D, scales, rots, opac, shs, colors, cov3Dp, radii, geom,
                              clamped, rect, tiles, dkeys, s_h0, flag_word);
  }
  __syncthreads();
  if (s_h0[threadIdx.x]) atomicAdd(&ghist0[threadIdx.x], s_h0[threadIdx.x]);
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
  const float tz = ((view[2] * px + view[6] * py) + view[10] * pz) + view[14];
  present[i] = tz > NEAR_Z ? 1 : 0;
}

}  // namespace

int launch_preprocess_fwd(const View& v, int P, int K, const float* means3D, const float* scales,
                          const float* rots, const float* opac, const float* shs,
                          const float* colors, const float* cov3D, int32_t* radii, Geom* geom,
                          uint8_t* clamped, ushort4* rect, uint32_t* tiles, uint32_t* dkeys,
                          void* radix_tmp, bool debug, cudaStream_t st) {
  if (P == 0) return GSB_OK;
  int sms = 0;
  { const int rc_sm = device_sm_count(&sms); if (rc_sm) return rc_sm; }
  const int chunks = (P + 255) / 256;
  const int grid = chunks < sms * PRE_CTAS_PER_SM ? chunks : sms * PRE_CTAS_PER_SM;
  // the kernel also accumulates the digit-0 histogram of the depth keys for the depth sort
  int rc = radix_prepare(P, 32, radix_tmp, st);
  if (rc) return rc;
  preprocess_fwd_kernel<<<grid, 256, 0, st>>>(v, P, K, means3D, scales, rots, opac, shs, colors, cov3D,
                                              radii, geom, clamped, rect, tiles, dkeys, radix_hist0(radix_tmp),
                                              radix_flag_word(radix_tmp));
  GSB_POST_LAUNCH(debug, st, "preprocess_fwd_kernel");
  return GSB_OK;
}

int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present,
                        cudaStream_t st) {
  if (P == 0) return GSB_OK;
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, view, present);
  GSB_POST_LAUNCH(false, st, "mark_visible_kernel");
  return GSB_OK;
}

}  // namespace gsb
