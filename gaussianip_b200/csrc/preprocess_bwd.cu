// Backward stage 2: per-Gaussian gradients.  Replaces computeCov2DCUDA + preprocessCUDA
// (backward) of the external operator ("Backward preprocess") in one
// kernel: conic -> cov2D -> (cov3D, view-space mean) -> (scale, rotation, mean), plus the
// mean gradient through the perspective divide (means2D), the depth output and the SH
// colour.  One thread per Gaussian; the forward intermediates are recomputed from the
// inputs instead of being stored (saves 24 B/Gaussian of cov3D traffic each way).
// With accumulate != 0 results are added to the outputs (beta = 1) so all views of a step
// land in one gradient bucket.  The kernel takes up to GSB_MAX_VIEWS views at once: the
// per-view contributions are summed in registers and every output is written ONCE, instead of
// one read-modify-write of all gradient tensors per view.
#include <atomic>
#include <mutex>

#include "gsb_common.cuh"

namespace gsb {

namespace {

__device__ __forceinline__ void put(float* p, float v, int acc) { *p = acc ? (*p + v) : v; }

// ---- accumulate mode 2: the outputs are NVLS MULTICAST addresses of zero-initialised symmetric buffers
// mapped on every GPU of the step; a multimem.red adds the value into the copy of EVERY rank inside the
// NVSwitch, so when all ranks' kernels have finished each copy holds the sum over ranks — the gradient
// all-reduce happens in this kernel's epilogue and overlaps its compute (fused exchange).
//
// ---- accumulate mode 3 (scales with the number of ranks): rows are OWNED by ranks in contiguous blocks; the
// epilogue adds each row with a plain red.global into the OWNER's copy only (peer memory over NVLink, every
// GPU receives (N-1)/N of one gradient instead of N gradients), and a small second kernel lets every owner
// multicast its reduced block to all copies (gsb_exchange_gather).
template <bool MULTIMEM>
__device__ __forceinline__ void mc_red(float* p, float v) {
  if (MULTIMEM) asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
  else asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
template <bool MULTIMEM>
__device__ __forceinline__ void mc_red4(float* p, const float4 v) {
  if (MULTIMEM)
    asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
  else
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
// One warp's rows of a [P, W] tensor are 32*W contiguous floats starting 16 B aligned (first row of a warp is
// a multiple of 32): stage them through shared memory and emit full 16 B reds instead of W strided 4 B ones.
template <int W, bool MULTIMEM>
__device__ __forceinline__ void mc_rows(float* __restrict__ out, float* __restrict__ stage, int first, int n_valid,
                                        int lane, const float (&v)[W]) {
  __syncwarp();
#pragma unroll
  for (int k = 0; k < W; ++k) stage[lane * W + k] = v[k];
  __syncwarp();
  const int floats = n_valid * W;
  float* base = out + (size_t)first * W;
#pragma unroll
  for (int q = lane; q < (32 * W) / 4; q += 32) {
    if (4 * q + 3 < floats) {
      const float4 t = reinterpret_cast<const float4*>(stage)[q];
      if (t.x != 0.f || t.y != 0.f || t.z != 0.f || t.w != 0.f) mc_red4<MULTIMEM>(base + 4 * q, t);
    } else {
      for (int e = 4 * q; e < floats; ++e)
        if (stage[e] != 0.f) mc_red<MULTIMEM>(base + e, stage[e]);
    }
  }
}

template <int DEG>
struct Accum {            // per-Gaussian gradient sums over the views of one launch (registers)
  float dp[3], d2[2], dop, dsc[3], dq[4], dcov[6], dcol[3];
  float dsh[3 * (DEG + 1) * (DEG + 1)];
};

// Adds the contribution of ONE view to the sums in A (Gaussian i is visible in that view).
// DEG = active SH degree (compile time, so the basis / coefficient loops unroll and the 3(DEG+1)^2
// SH gradient sums stay in registers); sh = this Gaussian's coefficient row (staged in shared memory).
// What one view contributes for Gaussian i, as loaded from that view's blocks: the gradient record written by the
// backward blend, the pre-scaled conic + activated opacity of the forward record, the SH clamp mask.  Loaded one
// view AHEAD of the arithmetic (the kernel is bound by the latency of these dependent loads, ncu r2d: long
// scoreboard 4.5 of ~7 stall cycles per issue at 23 % occupancy).
struct ViewRec {
  float4 g0, g1, g2, con;
  uint32_t cl;
};
__device__ __forceinline__ void load_view_rec(const BwdView& bv, int i, bool want_clamped, ViewRec& r) {
  const float4* gp = reinterpret_cast<const float4*>(bv.ggrad + i);
  r.g0 = gp[0]; r.g1 = gp[1]; r.g2 = gp[2];
  r.con = reinterpret_cast<const float4*>(bv.geom + i)[1];
  r.cl = want_clamped ? bv.clamped[i] : 0u;
}

template <int DEG>
__device__ __forceinline__ void
view_contrib(const View& v, int i, const float* sV, const float* sM, const float* sCam, const float* sK,
             const float px, const float py, const float pz, const float (&sc_act)[3], const float4 q_act,
             const float* sh, const float* __restrict__ cov3Dp, const ViewRec& rec, bool precomp_color,
             Accum<DEG>& A) {
  constexpr int ncoef = (DEG + 1) * (DEG + 1);
  const float4 g0 = rec.g0, g1 = rec.g1, g2 = rec.g2;
  // moments -> gradients of the screen-space mean (NDC units) and of the conic
  float4 con = rec.con;                                              // pre-scaled conic, opacity
  con.x *= 1.0f / CONIC_SCALE_AC; con.y *= 1.0f / CONIC_SCALE_B; con.z *= 1.0f / CONIC_SCALE_AC;   // back to A, B, C
  const float gx = -(con.x * g0.x + con.y * g0.y) * (0.5f * (float)v.W);
  const float gy = -(con.z * g0.y + con.y * g0.x) * (0.5f * (float)v.H);
  const float gA = -0.5f * g0.z, gB = -g0.w, gC = -0.5f * g1.x, gop = g1.y, gdepth = g1.z;
  float grgb[3] = {g1.w, g2.x, g2.y};

  const float tx = sV[0] * px + sV[4] * py + sV[8] * pz + sV[12];
  const float ty = sV[1] * px + sV[5] * py + sV[9] * pz + sV[13];
  const float tz = sV[2] * px + sV[6] * py + sV[10] * pz + sV[14];
  float dpx = 0.f, dpy = 0.f, dpz = 0.f;

  // ---- colour -------------------------------------------------------------------------
  if (precomp_color) {
    A.dcol[0] += grgb[0]; A.dcol[1] += grgb[1]; A.dcol[2] += grgb[2];
  } else {
    const uint32_t cl = rec.cl;
    if (cl & 1) grgb[0] = 0.f;
    if (cl & 2) grgb[1] = 0.f;
    if (cl & 4) grgb[2] = 0.f;
    const float ox = px - sCam[0], oy = py - sCam[1], oz = pz - sCam[2];
    const float sum2 = ox * ox + oy * oy + oz * oz;
    const float inv = rsqrtf(sum2);
    const float x = ox * inv, y = oy * inv, z = oz * inv;
    const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
    float basis[16];
    float ddx[3] = {0.f, 0.f, 0.f}, ddy[3] = {0.f, 0.f, 0.f}, ddz[3] = {0.f, 0.f, 0.f};
    basis[0] = C0;
    if (DEG > 0) {
      basis[1] = -C1 * y; basis[2] = C1 * z; basis[3] = -C1 * x;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ddx[c] = -C1 * sh[9 + c]; ddy[c] = -C1 * sh[3 + c]; ddz[c] = C1 * sh[6 + c];
      }
      if (DEG > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        const float a0 = 1.0925484305920792f, a1 = -1.0925484305920792f, a2 = 0.31539156525252005f,
                    a3 = -1.0925484305920792f, a4 = 0.5462742152960396f;
        basis[4] = a0 * xy; basis[5] = a1 * yz; basis[6] = a2 * (2.f * zz - xx - yy);
        basis[7] = a3 * xz; basis[8] = a4 * (xx - yy);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          ddx[c] += a0 * y * sh[12 + c] + a2 * (-2.f * x) * sh[18 + c] + a3 * z * sh[21 + c] + a4 * 2.f * x * sh[24 + c];
          ddy[c] += a0 * x * sh[12 + c] + a1 * z * sh[15 + c] + a2 * (-2.f * y) * sh[18 + c] + a4 * (-2.f * y) * sh[24 + c];
          ddz[c] += a1 * y * sh[15 + c] + a2 * 4.f * z * sh[18 + c] + a3 * x * sh[21 + c];
        }
        if (DEG > 2) {
          const float b0 = -0.5900435899266435f, b1 = 2.890611442640554f, b2 = -0.4570457994644658f,
                      b3 = 0.3731763325901154f, b4 = -0.4570457994644658f, b5 = 1.445305721320277f,
                      b6 = -0.5900435899266435f;
          basis[9] = b0 * y * (3.f * xx - yy);
          basis[10] = b1 * xy * z;
          basis[11] = b2 * y * (4.f * zz - xx - yy);
          basis[12] = b3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
          basis[13] = b4 * x * (4.f * zz - xx - yy);
          basis[14] = b5 * z * (xx - yy);
          basis[15] = b6 * x * (xx - 3.f * yy);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            ddx[c] += b0 * sh[27 + c] * 6.f * xy + b1 * sh[30 + c] * yz + b2 * sh[33 + c] * (-2.f * xy) +
                      b3 * sh[36 + c] * (-6.f * xz) + b4 * sh[39 + c] * (4.f * zz - 3.f * xx - yy) +
                      b5 * sh[42 + c] * 2.f * xz + b6 * sh[45 + c] * (3.f * xx - 3.f * yy);
            ddy[c] += b0 * sh[27 + c] * (3.f * xx - 3.f * yy) + b1 * sh[30 + c] * xz +
                      b2 * sh[33 + c] * (4.f * zz - xx - 3.f * yy) + b3 * sh[36 + c] * (-6.f * yz) +
                      b4 * sh[39 + c] * (-2.f * xy) + b5 * sh[42 + c] * (-2.f * yz) + b6 * sh[45 + c] * (-6.f * xy);
            ddz[c] += b1 * sh[30 + c] * xy + b2 * sh[33 + c] * 8.f * yz + b3 * sh[36 + c] * (6.f * zz - 3.f * xx - 3.f * yy) +
                      b4 * sh[39 + c] * 8.f * xz + b5 * sh[42 + c] * (xx - yy);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < ncoef; ++k) {
#pragma unroll
      for (int c = 0; c < 3; ++c) A.dsh[3 * k + c] += basis[k] * grgb[c];
    }
    // through the view direction normalisation
    const float dLx = ddx[0] * grgb[0] + ddx[1] * grgb[1] + ddx[2] * grgb[2];
    const float dLy = ddy[0] * grgb[0] + ddy[1] * grgb[1] + ddy[2] * grgb[2];
    const float dLz = ddz[0] * grgb[0] + ddz[1] * grgb[1] + ddz[2] * grgb[2];
    const float inv3 = inv * inv * inv;
    dpx += ((sum2 - ox * ox) * dLx - oy * ox * dLy - oz * ox * dLz) * inv3;
    dpy += (-ox * oy * dLx + (sum2 - oy * oy) * dLy - oz * oy * dLz) * inv3;
    dpz += (-ox * oz * dLx - oy * oz * dLy + (sum2 - oz * oz) * dLz) * inv3;
  }

  // ---- 3D covariance (recomputed) -------------------------------------------------------
  float S00, S01, S02, S11, S12, S22;
  float m[3][3], R[3][3], s[3] = {0.f, 0.f, 0.f};
  float qr = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
  if (cov3Dp) {
    const float* c = cov3Dp + 6 * (size_t)i;
    S00 = c[0]; S01 = c[1]; S02 = c[2]; S11 = c[3]; S12 = c[4]; S22 = c[5];
  } else {
    s[0] = v.scale_mod * sc_act[0]; s[1] = v.scale_mod * sc_act[1]; s[2] = v.scale_mod * sc_act[2];
    const float4 q = q_act;
    qr = q.x; qx = q.y; qy = q.z; qz = q.w;
    R[0][0] = 1.f - 2.f * (qy * qy + qz * qz); R[0][1] = 2.f * (qx * qy - qr * qz); R[0][2] = 2.f * (qx * qz + qr * qy);
    R[1][0] = 2.f * (qx * qy + qr * qz); R[1][1] = 1.f - 2.f * (qx * qx + qz * qz); R[1][2] = 2.f * (qy * qz - qr * qx);
    R[2][0] = 2.f * (qx * qz - qr * qy); R[2][1] = 2.f * (qy * qz + qr * qx); R[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int k = 0; k < 3; ++k) m[a][k] = R[a][k] * s[k];
    S00 = m[0][0] * m[0][0] + m[0][1] * m[0][1] + m[0][2] * m[0][2];
    S01 = m[0][0] * m[1][0] + m[0][1] * m[1][1] + m[0][2] * m[1][2];
    S02 = m[0][0] * m[2][0] + m[0][1] * m[2][1] + m[0][2] * m[2][2];
    S11 = m[1][0] * m[1][0] + m[1][1] * m[1][1] + m[1][2] * m[1][2];
    S12 = m[1][0] * m[2][0] + m[1][1] * m[2][1] + m[1][2] * m[2][2];
    S22 = m[2][0] * m[2][0] + m[2][1] * m[2][1] + m[2][2] * m[2][2];
  }

  // ---- EWA projection (recomputed) and its backward ------------------------------------------
  const float limx = 1.3f * sK[0], limy = 1.3f * sK[1];
  const float tz_inv = 1.0f / tz;
  const float txtz = tx * tz_inv, tytz = ty * tz_inv;
  const float x_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  const float y_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  const float txc = fminf(limx, fmaxf(-limx, txtz)) * tz;
  const float tyc = fminf(limy, fmaxf(-limy, tytz)) * tz;
  const float tz2 = tz_inv * tz_inv, tz3 = tz2 * tz_inv;
  const float fx = sK[2], fy = sK[3];
  const float J00 = fx * tz_inv, J02 = -fx * txc * tz2, J11 = fy * tz_inv, J12 = -fy * tyc * tz2;
  float T0[3], T1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    T0[j] = J00 * sV[0 + 4 * j] + J02 * sV[2 + 4 * j];
    T1[j] = J11 * sV[1 + 4 * j] + J12 * sV[2 + 4 * j];
  }
  const float Sg[3][3] = {{S00, S01, S02}, {S01, S11, S12}, {S02, S12, S22}};
  float U0[3], U1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    U0[j] = T0[0] * Sg[0][j] + T0[1] * Sg[1][j] + T0[2] * Sg[2][j];
    U1[j] = T1[0] * Sg[0][j] + T1[1] * Sg[1][j] + T1[2] * Sg[2][j];
  }
  const float a = U0[0] * T0[0] + U0[1] * T0[1] + U0[2] * T0[2] + LOWPASS;
  const float b = U0[0] * T1[0] + U0[1] * T1[1] + U0[2] * T1[2];
  const float c = U1[0] * T1[0] + U1[1] * T1[1] + U1[2] * T1[2] + LOWPASS;
  const float det = a * c - b * b;
  const float d2 = 1.0f / (det * det + 0.0000001f);
  float ga = 0.f, gb = 0.f, gc = 0.f;
  if (det != 0.0f) {
    ga = d2 * (-c * c * gA + b * c * gB + (det - a * c) * gC);
    gc = d2 * (-a * a * gC + a * b * gB + (det - a * c) * gA);
    gb = d2 * (2.f * b * c * gA - (det + 2.f * b * b) * gB + 2.f * a * b * gC);
  }
  // dL/dSigma (all nine entries independent): Gs = T^T [[ga, gb/2],[gb/2, gc]] T
  float Gs[3][3];
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int q = 0; q < 3; ++q)
      Gs[p][q] = T0[p] * T0[q] * ga + 0.5f * (T0[p] * T1[q] + T1[p] * T0[q]) * gb + T1[p] * T1[q] * gc;

  float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float dT0 = 2.f * ga * U0[j] + gb * U1[j];
    const float dT1 = gb * U0[j] + 2.f * gc * U1[j];
    dJ00 += dT0 * sV[0 + 4 * j]; dJ02 += dT0 * sV[2 + 4 * j];
    dJ11 += dT1 * sV[1 + 4 * j]; dJ12 += dT1 * sV[2 + 4 * j];
  }
  const float dtx = x_mul * (-fx * tz2 * dJ02);
  const float dty = y_mul * (-fy * tz2 * dJ12);
  const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + 2.f * fx * txc * tz3 * dJ02 + 2.f * fy * tyc * tz3 * dJ12;
  dpx += sV[0] * dtx + sV[1] * dty + sV[2] * dtz;
  dpy += sV[4] * dtx + sV[5] * dty + sV[6] * dtz;
  dpz += sV[8] * dtx + sV[9] * dty + sV[10] * dtz;

  // ---- means2D (perspective divide) and depth -------------------------------------------
  const float hx = sM[0] * px + sM[4] * py + sM[8] * pz + sM[12];
  const float hy = sM[1] * px + sM[5] * py + sM[9] * pz + sM[13];
  const float hw = sM[3] * px + sM[7] * py + sM[11] * pz + sM[15];
  const float pw = 1.0f / (hw + 0.0000001f);
  const float mul1 = hx * pw * pw, mul2 = hy * pw * pw;
  dpx += (sM[0] * pw - sM[3] * mul1) * gx + (sM[1] * pw - sM[3] * mul2) * gy + sV[2] * gdepth;
  dpy += (sM[4] * pw - sM[7] * mul1) * gx + (sM[5] * pw - sM[7] * mul2) * gy + sV[6] * gdepth;
  dpz += (sM[8] * pw - sM[11] * mul1) * gx + (sM[9] * pw - sM[11] * mul2) * gy + sV[10] * gdepth;

  A.dp[0] += dpx; A.dp[1] += dpy; A.dp[2] += dpz;
  A.d2[0] += gx; A.d2[1] += gy;
  A.dop += gop;

  // ---- cov3D -> scale / rotation ---------------------------------------------------------
  if (cov3Dp) {
    A.dcov[0] += Gs[0][0]; A.dcov[1] += 2.f * Gs[0][1]; A.dcov[2] += 2.f * Gs[0][2];
    A.dcov[3] += Gs[1][1]; A.dcov[4] += 2.f * Gs[1][2]; A.dcov[5] += Gs[2][2];
  } else {
    float dM[3][3];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        dM[p][k] = 2.f * (Gs[p][0] * m[0][k] + Gs[p][1] * m[1][k] + Gs[p][2] * m[2][k]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      A.dsc[k] += v.scale_mod * (R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k]);
    float dR[3][3];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int k = 0; k < 3; ++k) dR[p][k] = s[k] * dM[p][k];
    A.dq[0] += 2.f * (-qz * dR[0][1] + qy * dR[0][2] + qz * dR[1][0] - qx * dR[1][2] - qy * dR[2][0] + qx * dR[2][1]);
    A.dq[1] += 2.f * (qy * dR[0][1] + qz * dR[0][2] + qy * dR[1][0] - 2.f * qx * dR[1][1] - qr * dR[1][2] +
                      qz * dR[2][0] + qr * dR[2][1] - 2.f * qx * dR[2][2]);
    A.dq[2] += 2.f * (-2.f * qy * dR[0][0] + qx * dR[0][1] + qr * dR[0][2] + qx * dR[1][0] + qz * dR[1][2] -
                      qr * dR[2][0] + qz * dR[2][1] - 2.f * qy * dR[2][2]);
    A.dq[3] += 2.f * (-2.f * qz * dR[0][0] - qr * dR[0][1] + qx * dR[0][2] + qr * dR[1][0] - 2.f * qz * dR[1][1] +
                      qy * dR[1][2] + qx * dR[2][0] + qy * dR[2][1]);
  }
}

#ifndef GSB_PBWD_MINB
#define GSB_PBWD_MINB 2
#endif
// LPG = lanes per Gaussian (MODE 0 only): the views of a launch are dealt to LPG adjacent lanes, each lane loads and
// differentiates its views, and the partial sums are added across the lanes with xor-shuffles.  The kernel is bound by
// the latency of its per-view dependent loads at ~25 % occupancy (128 registers), so 4x more threads with a quarter
// of the views each put 4x more loads in flight for the same arithmetic.
template <int DEG, int MODE, int LPG>   // MODE 0: store / accumulate locally, 2: multicast red into all copies, 3: red into the owner's copy
__global__ void __launch_bounds__(256, (DEG <= 1 ? GSB_PBWD_MINB : 2))
preprocess_bwd_kernel(BwdBatch B, int P, int K, const float* __restrict__ means3D,
                      const float* __restrict__ scales, const float* __restrict__ rots,
                      const float* __restrict__ shs, const float* __restrict__ cov3Dp,
                      float* __restrict__ dmeans3D, float* __restrict__ dmeans2D, float* __restrict__ dshs,
                      float* __restrict__ dcolors, float* __restrict__ dopac, float* __restrict__ dscales,
                      float* __restrict__ drots, float* __restrict__ dcov3D, int acc, const ExchangePeers X) {
  constexpr int NC3 = 3 * (DEG + 1) * (DEG + 1);
  constexpr bool MC = MODE != 0;
  constexpr bool MM = MODE == 2;
  __shared__ float sV[GSB_MAX_VIEWS][16], sM[GSB_MAX_VIEWS][16], sCam[GSB_MAX_VIEWS][4], sK[GSB_MAX_VIEWS][4];
  for (int t = threadIdx.x; t < B.V * 16; t += blockDim.x) {
    sV[t >> 4][t & 15] = B.a[t >> 4].v.view[t & 15];
    sM[t >> 4][t & 15] = B.a[t >> 4].v.proj[t & 15];
  }
  for (int t = threadIdx.x; t < B.V * 3; t += blockDim.x) sCam[t / 3][t % 3] = B.a[t / 3].v.campos[t % 3];
  if (threadIdx.x >= 128 && threadIdx.x < 128 + B.V) load_intrinsics(B.a[threadIdx.x - 128].v, sK[threadIdx.x - 128]);
  __syncthreads();
  static_assert(MODE == 0 || LPG == 1, "the exchange epilogue stages one row per lane");
  const int i = (blockIdx.x * 256 + threadIdx.x) / LPG;
  const int sub = threadIdx.x % LPG;
  // a Gaussian's coefficient row is K*12 contiguous bytes; the rows of a warp are contiguous too, so
  // reading one's own row touches every fetched sector completely (re-reads across views hit L1)
  const float* my_sh = shs ? shs + (size_t)i * K * 3 : nullptr;

  Accum<DEG> A;
  int rmax = 0;                      // largest radius of this Gaussian over the views of the launch
#pragma unroll
  for (int k = 0; k < 3; ++k) { A.dp[k] = 0.f; A.dsc[k] = 0.f; A.dcol[k] = 0.f; }
  A.d2[0] = A.d2[1] = 0.f; A.dop = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) A.dq[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) A.dcov[k] = 0.f;
#pragma unroll
  for (int k = 0; k < NC3; ++k) A.dsh[k] = 0.f;

  const unsigned wm = __ballot_sync(0xffffffffu, i < P);     // the lanes that take the branch below (whole LPG groups)
  if (i < P) {
    // scale / rotation are loaded (and, for raw model parameters, activated) once for all views
    const int raw = B.a[0].v.raw;
    float sc[3] = {0.f, 0.f, 0.f}, qn = 1.f, o_act = 0.f;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!cov3Dp) {
      sc[0] = scales[3 * i]; sc[1] = scales[3 * i + 1]; sc[2] = scales[3 * i + 2];
      if (raw & GSB_RAW_SCALE) { sc[0] = expf(sc[0]); sc[1] = expf(sc[1]); sc[2] = expf(sc[2]); }
      q = reinterpret_cast<const float4*>(rots)[i];
      if (raw & GSB_RAW_ROTATION) q = act_normalize(q, &qn);
    }
    // radii of all views first (independent loads), then the per-view records one view ahead of the arithmetic
    const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
    uint32_t seen = 0;
#pragma unroll
    for (int vi = 0; vi < GSB_MAX_VIEWS; ++vi) {
      const int rv = vi < B.V ? B.a[vi].radii[i] : 0;
      if (rv > 0) seen |= 1u << vi;                                 // gradients of culled Gaussians are exactly 0
      rmax = max(rmax, rv);
    }
    if (LPG > 1) seen &= (LPG == 4 ? 0x11u : 0x55u) << sub;          // this lane's share of the views
    const bool want_cl = dcolors == nullptr;
    ViewRec cur, nxt;
    int vi = seen ? __ffs(seen) - 1 : -1;
    if (vi >= 0) load_view_rec(B.a[vi], i, want_cl, cur);
    while (vi >= 0) {
      seen &= seen - 1;
      const int vn = seen ? __ffs(seen) - 1 : -1;
      if (vn >= 0) load_view_rec(B.a[vn], i, want_cl, nxt);
      view_contrib<DEG>(B.a[vi].v, i, sV[vi], sM[vi], sCam[vi], sK[vi], px, py, pz, sc, q, my_sh, cov3Dp, cur,
                        dcolors != nullptr, A);
      o_act = cur.con.w;                                            // activated opacity, as the forward stored it
      cur = nxt;
      vi = vn;
    }
    if (LPG > 1) {
      // add the lanes' partial sums (all LPG lanes of a Gaussian are in one warp and take the same branches here)
#pragma unroll
      for (int d = 1; d < LPG; d <<= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          A.dp[k] += __shfl_xor_sync(wm, A.dp[k], d);
          A.dsc[k] += __shfl_xor_sync(wm, A.dsc[k], d);
          A.dcol[k] += __shfl_xor_sync(wm, A.dcol[k], d);
        }
        A.d2[0] += __shfl_xor_sync(wm, A.d2[0], d);
        A.d2[1] += __shfl_xor_sync(wm, A.d2[1], d);
        A.dop += __shfl_xor_sync(wm, A.dop, d);
#pragma unroll
        for (int k = 0; k < 4; ++k) A.dq[k] += __shfl_xor_sync(wm, A.dq[k], d);
        if (cov3Dp) {
#pragma unroll
          for (int k = 0; k < 6; ++k) A.dcov[k] += __shfl_xor_sync(wm, A.dcov[k], d);
        }
#pragma unroll
        for (int k = 0; k < NC3; ++k) A.dsh[k] += __shfl_xor_sync(wm, A.dsh[k], d);
        o_act = fmaxf(o_act, __shfl_xor_sync(wm, o_act, d));
      }
    }
    // chain rule to the RAW parameters, applied once to the sum over views
    if (raw & GSB_RAW_OPACITY) A.dop = A.dop * (1.0f - o_act) * o_act;                    // sigmoid'
    if (!cov3Dp) {
      if (raw & GSB_RAW_SCALE) { A.dsc[0] *= sc[0]; A.dsc[1] *= sc[1]; A.dsc[2] *= sc[2]; }   // exp'
      if (raw & GSB_RAW_ROTATION) {                                                       // (I - q q^T) / |raw|
        const float dot = q.x * A.dq[0] + q.y * A.dq[1] + q.z * A.dq[2] + q.w * A.dq[3];
        A.dq[0] = (A.dq[0] - q.x * dot) / qn; A.dq[1] = (A.dq[1] - q.y * dot) / qn;
        A.dq[2] = (A.dq[2] - q.z * dot) / qn; A.dq[3] = (A.dq[3] - q.w * dot) / qn;
      }
    }
  }

  if (MC) {
    // every lane takes part (shared-memory staging is warp-wide); rows past P carry zeros and are cut off
    __shared__ float s_stage[8][32 * 6];
    const int lane = threadIdx.x & 31;
    float* stage = s_stage[threadIdx.x >> 5];
    const int first = i - lane;
    const int n_valid = min(32, P - first);
    // extras riding in the exchange: the step's scalar (one thread of the launch) and the radii maximum
    if (X.scalar_in && blockIdx.x == 0 && threadIdx.x == 0) {
      float* dst = X.scalar_out;
      if (MODE == 3) dst = reinterpret_cast<float*>(reinterpret_cast<char*>(dst) + (X.base[0] - X.base[X.rank]));
      mc_red<MM>(dst, *X.scalar_in);
    }
    if (n_valid <= 0) return;               // warp-uniform
    if (X.radii_max && i < P && rmax > 0) {
      int32_t* dst = X.radii_max + i;
      if (MODE == 3) {
        int owner = (int)(first / X.rows_per_rank);
        owner = owner < X.world ? owner : X.world - 1;
        dst = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(dst) + (X.base[owner] - X.base[X.rank]));
        asm volatile("red.relaxed.sys.global.max.s32 [%0], %1;" ::"l"(dst), "r"(rmax) : "memory");
      } else {
        asm volatile("multimem.red.relaxed.sys.global.max.s32 [%0], %1;" ::"l"(dst), "r"(rmax) : "memory");
      }
    }
    if (MODE == 3) {
      // this warp's 32 rows belong to one owner (blocks are multiples of 32 rows): retarget every output
      // from the local copy to the same offset inside the owner's copy
      int owner = (int)(first / X.rows_per_rank);
      owner = owner < X.world ? owner : X.world - 1;
      const ptrdiff_t shift = X.base[owner] - X.base[X.rank];
      auto to_owner = [&](float* p) { return p ? reinterpret_cast<float*>(reinterpret_cast<char*>(p) + shift) : p; };
      dmeans3D = to_owner(dmeans3D); dmeans2D = to_owner(dmeans2D); dopac = to_owner(dopac);
      dcolors = to_owner(dcolors); dcov3D = to_owner(dcov3D); dscales = to_owner(dscales);
      drots = to_owner(drots); dshs = to_owner(dshs);
    }
    mc_rows<3, MM>(dmeans3D, stage, first, n_valid, lane, A.dp);
    { const float t[3] = {A.d2[0], A.d2[1], 0.f}; mc_rows<3, MM>(dmeans2D, stage, first, n_valid, lane, t); }
    { const float t[1] = {A.dop}; mc_rows<1, MM>(dopac, stage, first, n_valid, lane, t); }
    if (dcolors) mc_rows<3, MM>(dcolors, stage, first, n_valid, lane, A.dcol);
    if (cov3Dp) {
      mc_rows<6, MM>(dcov3D, stage, first, n_valid, lane, A.dcov);
    } else {
      mc_rows<3, MM>(dscales, stage, first, n_valid, lane, A.dsc);
      if (i < P && (A.dq[0] != 0.f || A.dq[1] != 0.f || A.dq[2] != 0.f || A.dq[3] != 0.f))
        mc_red4<MM>(drots + 4 * (size_t)i, make_float4(A.dq[0], A.dq[1], A.dq[2], A.dq[3]));
    }
    if (dshs) {
      if (NC3 == 3 && K == 1) {
        const float t[3] = {A.dsh[0], A.dsh[1], A.dsh[2]};
        mc_rows<3, MM>(dshs, stage, first, n_valid, lane, t);
      } else if (i < P) {
        float* dsh = dshs + (size_t)i * K * 3;
        if (((K * 3) & 3) == 0) {
#pragma unroll
          for (int k = 0; k + 3 < NC3; k += 4)
            if (A.dsh[k] != 0.f || A.dsh[k + 1] != 0.f || A.dsh[k + 2] != 0.f || A.dsh[k + 3] != 0.f)
              mc_red4<MM>(dsh + k, make_float4(A.dsh[k], A.dsh[k + 1], A.dsh[k + 2], A.dsh[k + 3]));
#pragma unroll
          for (int k = NC3 & ~3; k < NC3; ++k)
            if (A.dsh[k] != 0.f) mc_red<MM>(dsh + k, A.dsh[k]);
        } else {
#pragma unroll
          for (int k = 0; k < NC3; ++k)
            if (A.dsh[k] != 0.f) mc_red<MM>(dsh + k, A.dsh[k]);
        }
      }
    }
    return;
  }

  if (i < P && sub == 0) {
    put(dmeans3D + 3 * i, A.dp[0], acc); put(dmeans3D + 3 * i + 1, A.dp[1], acc); put(dmeans3D + 3 * i + 2, A.dp[2], acc);
    put(dmeans2D + 3 * i, A.d2[0], acc); put(dmeans2D + 3 * i + 1, A.d2[1], acc);
    if (!acc) dmeans2D[3 * i + 2] = 0.f;
    put(dopac + i, A.dop, acc);
    if (dcolors) {
      put(dcolors + 3 * i, A.dcol[0], acc); put(dcolors + 3 * i + 1, A.dcol[1], acc); put(dcolors + 3 * i + 2, A.dcol[2], acc);
    }
    if (cov3Dp) {
      float* o = dcov3D + 6 * (size_t)i;
#pragma unroll
      for (int k = 0; k < 6; ++k) put(o + k, A.dcov[k], acc);
    } else {
      put(dscales + 3 * i, A.dsc[0], acc); put(dscales + 3 * i + 1, A.dsc[1], acc); put(dscales + 3 * i + 2, A.dsc[2], acc);
      float4* o = reinterpret_cast<float4*>(drots) + i;
      if (acc) {
        const float4 old = *o;
        *o = make_float4(old.x + A.dq[0], old.y + A.dq[1], old.z + A.dq[2], old.w + A.dq[3]);
      } else {
        *o = make_float4(A.dq[0], A.dq[1], A.dq[2], A.dq[3]);
      }
    }
    if (dshs) {
      float* dsh = dshs + (size_t)i * K * 3;
      if (((K * 3) & 3) == 0) {              // rows are 16 B aligned: vector stores
#pragma unroll
        for (int k = 0; k + 3 < NC3; k += 4) {
          float4* o = reinterpret_cast<float4*>(dsh + k);
          float4 val = make_float4(A.dsh[k], A.dsh[k + 1], A.dsh[k + 2], A.dsh[k + 3]);
          if (acc) { const float4 old = *o; val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w; }
          *o = val;
        }
#pragma unroll
        for (int k = NC3 & ~3; k < NC3; ++k) put(dsh + k, A.dsh[k], acc);
      } else {
#pragma unroll
        for (int k = 0; k < NC3; ++k) put(dsh + k, A.dsh[k], acc);
      }
      if (!acc) for (int k = NC3; k < 3 * K; ++k) dsh[k] = 0.f;   // coefficients above the active degree
    }
  }
}

// Second half of the owner-push exchange: copy `count` floats at `off` of the local copy to the same offset of
// EVERY copy through the multicast mapping (16 B stores; offsets and counts are multiples of 4 floats except
// possibly the last segment's tail, which goes out as scalars).
__global__ void __launch_bounds__(256) exchange_gather_kernel(const float* __restrict__ local, float* __restrict__ mc,
                                                              ExchangeSegments S) {
  const int seg = blockIdx.y;
  const long long off = S.off[seg], n = S.count[seg];
  const float4* src = reinterpret_cast<const float4*>(local + off);
  float* dst = mc + off;
  const long long n4 = n >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[q];
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long e = (n4 << 2) + threadIdx.x;
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst + e), "f"(local[off + e]) : "memory");
  }
}

}  // namespace

static ExchangePeers g_peers = {};
static std::mutex g_peers_mu;       // autograd may run the backward on its own thread

int set_exchange_peers(int world, int rank, long long rows_per_rank, const void* const* bases) {
  if (world < 1 || world > GSB_MAX_RANKS || rank < 0 || rank >= world || rows_per_rank <= 0 || (rows_per_rank & 31) || !bases)
    return GSB_E_INVALID;
  ExchangePeers x = {};
  x.world = world; x.rank = rank; x.rows_per_rank = rows_per_rank;
  for (int r = 0; r < world; ++r) {
    if (!bases[r]) return GSB_E_INVALID;
    x.base[r] = static_cast<char*>(const_cast<void*>(bases[r]));
  }
  std::lock_guard<std::mutex> lk(g_peers_mu);
  x.radii_max = g_peers.radii_max; x.scalar_in = g_peers.scalar_in; x.scalar_out = g_peers.scalar_out;
  g_peers = x;
  return GSB_OK;
}

int set_exchange_aux(int32_t* radii_max, const float* scalar_in, float* scalar_out) {
  if ((scalar_in != nullptr) != (scalar_out != nullptr)) return GSB_E_INVALID;
  std::lock_guard<std::mutex> lk(g_peers_mu);
  g_peers.radii_max = radii_max; g_peers.scalar_in = scalar_in; g_peers.scalar_out = scalar_out;
  return GSB_OK;
}

int launch_exchange_gather(const float* local, float* mc, int n_seg, const long long* off, const long long* count,
                           cudaStream_t st) {
  if (n_seg <= 0) return GSB_OK;
  if (n_seg > GSB_EXCHANGE_MAX_SEGMENTS || !local || !mc) return GSB_E_INVALID;
  ExchangeSegments S = {};
  long long most = 0;
  for (int i = 0; i < n_seg; ++i) {
    if (off[i] < 0 || count[i] < 0 || (off[i] & 3)) return GSB_E_INVALID;
    S.off[i] = off[i]; S.count[i] = count[i];
    most = count[i] > most ? count[i] : most;
  }
  if (most == 0) return GSB_OK;
  long long blocks = (most / 4 + 255) / 256;
  blocks = blocks < 1 ? 1 : (blocks > 148 * 4 ? 148 * 4 : blocks);
  exchange_gather_kernel<<<dim3((unsigned)blocks, (unsigned)n_seg), 256, 0, st>>>(local, mc, S);
  GSB_POST_LAUNCH(false, st, "exchange_gather_kernel");
  return GSB_OK;
}

int launch_preprocess_bwd(const BwdBatch& B, int P, int K, const float* means3D, const float* scales,
                          const float* rots, const float* shs, const float* colors, const float* cov3D,
                          float* dmeans3D, float* dmeans2D, float* dshs, float* dcolors, float* dopac,
                          float* dscales, float* drots, float* dcov3D, int accumulate, bool debug,
                          cudaStream_t st) {
  if (P == 0) return GSB_OK;
  if (B.V < 1 || B.V > GSB_MAX_VIEWS) return GSB_E_INVALID;
  ExchangePeers peers;                       // snapshot of the process-wide table for this launch
  {
    std::lock_guard<std::mutex> lk(g_peers_mu);
    peers = g_peers;
  }
  if (accumulate == 3 && peers.world < 1) return GSB_E_INVALID;       // gsb_exchange_config was not called
  if (accumulate == 2 && peers.world < 1) { peers.world = 1; peers.rank = 0; peers.rows_per_rank = 32; }
  const int deg = B.a[0].v.sh_degree;
  for (int v = 1; v < B.V; ++v)
    if (B.a[v].v.sh_degree != deg) return GSB_E_INVALID;   // one degree per launch
  const int nc3 = 3 * (deg + 1) * (deg + 1);
  const size_t smem = 0;
  (void)nc3;
  const int grid = (P + 255) / 256;
  // lanes per Gaussian (local modes): 4 for three or more views, 2 for two, 1 otherwise; degree >= 2 keeps 1 (its
  // 27-48 SH gradient sums per lane would cost more shuffles than the split saves)
#ifndef GSB_PBWD_LPG
#define GSB_PBWD_LPG 1
#endif
  const int lpg = (accumulate >= 2 || deg >= 2) ? 1
                  : (B.V >= 3 ? (GSB_PBWD_LPG >= 4 ? 4 : GSB_PBWD_LPG) : (B.V == 2 && GSB_PBWD_LPG >= 2 ? 2 : 1));
#define GSB_LAUNCH_BWD(D)                                                                                       \
  do {                                                                                                          \
    if (smem > 48 * 1024) {                                                                                     \
      static std::atomic<unsigned long long> configured{0};                                                     \
      int dev = 0;                                                                                              \
      GSB_CUDA(cudaGetDevice(&dev));                                                                            \
      if (!(configured.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {                                 \
        GSB_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<D, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                      64 * 1024));                                                              \
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);                                     \
      }                                                                                                         \
    }                                                                                                           \
    if (accumulate == 2)                                                                                        \
      preprocess_bwd_kernel<D, 2, 1><<<grid, 256, smem, st>>>(B, P, K, means3D, scales, rots, shs, cov3D,       \
                                                           dmeans3D, dmeans2D, shs ? dshs : nullptr,            \
                                                           colors ? dcolors : nullptr, dopac, dscales,          \
                                                           drots, dcov3D, accumulate, peers);                 \
    else if (accumulate == 3)                                                                                   \
      preprocess_bwd_kernel<D, 3, 1><<<grid, 256, smem, st>>>(B, P, K, means3D, scales, rots, shs, cov3D,       \
                                                           dmeans3D, dmeans2D, shs ? dshs : nullptr,            \
                                                           colors ? dcolors : nullptr, dopac, dscales,          \
                                                           drots, dcov3D, accumulate, peers);                 \
    else if (lpg == 4)                                                                                          \
      preprocess_bwd_kernel<D, 0, 4><<<(int)(((long long)P * 4 + 255) / 256), 256, smem, st>>>(                 \
          B, P, K, means3D, scales, rots, shs, cov3D, dmeans3D, dmeans2D, shs ? dshs : nullptr,                 \
          colors ? dcolors : nullptr, dopac, dscales, drots, dcov3D, accumulate, peers);                        \
    else if (lpg == 2)                                                                                          \
      preprocess_bwd_kernel<D, 0, 2><<<(int)(((long long)P * 2 + 255) / 256), 256, smem, st>>>(                 \
          B, P, K, means3D, scales, rots, shs, cov3D, dmeans3D, dmeans2D, shs ? dshs : nullptr,                 \
          colors ? dcolors : nullptr, dopac, dscales, drots, dcov3D, accumulate, peers);                        \
    else                                                                                                        \
      preprocess_bwd_kernel<D, 0, 1><<<grid, 256, smem, st>>>(B, P, K, means3D, scales, rots, shs, cov3D,       \
                                                           dmeans3D, dmeans2D, shs ? dshs : nullptr,            \
                                                           colors ? dcolors : nullptr, dopac, dscales,          \
                                                           drots, dcov3D, accumulate, peers);                 \
  } while (0)
  switch (shs ? deg : 0) {
    case 0: GSB_LAUNCH_BWD(0); break;
    case 1: GSB_LAUNCH_BWD(1); break;
    case 2: GSB_LAUNCH_BWD(2); break;
    case 3: GSB_LAUNCH_BWD(3); break;
    default: return GSB_E_INVALID;
  }
#undef GSB_LAUNCH_BWD
  GSB_POST_LAUNCH(debug, st, "preprocess_bwd_kernel");
  return GSB_OK;
}

}  // namespace gsb
