// dlt_device.cuh — device code of the 4-lanes-per-problem 4x4 symmetric eigen-solver and of the
// DLT triangulation of one key-point, shared by dlt.cu (stand-alone kernels) and geom.cu.
#pragma once

#include "common.cuh"

namespace mvgeo {

constexpr int kDltThreads = 128;
// 3 sweeps already reach output rounding once the FP64 correction below is applied (simulated
// against the float64 SVD for V = 2..8 with 0..3 px noise); 4 leaves a margin.
constexpr int kDltSweeps = 4;

__device__ __forceinline__ double shfl4(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 4); }
__device__ __forceinline__ float shfl4(float v, int src) { return __shfl_sync(0xffffffffu, v, src, 4); }
__device__ __forceinline__ double xor4(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m, 4); }
__device__ __forceinline__ float xor4(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m, 4); }

// One Jacobi rotation in the (P,Q) plane of the 4x4 symmetric matrix distributed one column
// per lane (A[i] = A_{i,lane}); Vc is the matching column of the accumulated eigenvectors.
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(float (&A)[4], float (&Vc)[4], int lane) {
  const float app = shfl4(A[P], P);
  const float aqq = shfl4(A[Q], Q);
  const float apq = shfl4(A[P], Q);
  float c = 1.0f, s = 0.0f;
  if (fabsf(apq) > 1e-37f) {
    const float tau = (aqq - app) / (2.0f * apq);
    const float t = copysignf(1.0f, tau) / (fabsf(tau) + sqrtf(1.0f + tau * tau));
    c = rsqrtf(1.0f + t * t);
    s = t * c;
  }
  // column step (A J, V J): lanes P and Q mix their columns, the others keep theirs
  const int partner = lane == P ? Q : (lane == Q ? P : lane);
  float oa[4], ov[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    oa[i] = shfl4(A[i], partner);
    ov[i] = shfl4(Vc[i], partner);
  }
  const float cs = (lane == P || lane == Q) ? c : 1.0f;
  const float sn = lane == P ? -s : (lane == Q ? s : 0.0f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    A[i] = cs * A[i] + sn * oa[i];
    Vc[i] = cs * Vc[i] + sn * ov[i];
  }
  // row step (J^T .): lane-local
  const float ap = A[P], aq = A[Q];
  A[P] = c * ap - s * aq;
  A[Q] = s * ap + c * aq;
}

// DLT for the kDltThreads / 4 key-points of CTA `blk`. sP: the V projection matrices in shared memory.
__device__ __forceinline__ void dlt_body(const float* __restrict__ kp, const float* __restrict__ w, const float* sP,
                                         int64_t B, int V, int K, float min_weight, int weighted,
                                         float* __restrict__ X, float* __restrict__ resid,
                                         int32_t* __restrict__ n_views, int64_t blk) {
  const int lane = threadIdx.x & 3;
  const int64_t n_pts = B * K;
  int64_t pid = (int64_t)blk * (kDltThreads / 4) + (threadIdx.x >> 2);
  const bool active = pid < n_pts;
  if (!active) pid = n_pts - 1;  // keep the whole warp in the shuffles
  const int64_t b = pid / K;
  const int k = (int)(pid - b * K);

  // ---- M = A^T A, column `lane`, FP64 -------------------------------------------------
  double Mc[4] = {0.0, 0.0, 0.0, 0.0};
  int nv = 0;
  for (int v = 0; v < V; ++v) {
    const int64_t o = (b * V + v) * K + k;
    const float u_ = kp[2 * o], v_ = kp[2 * o + 1];
    const float wt = w ? w[o] : 1.0f;
    const bool ok = (wt >= min_weight) && isfinite(u_) && isfinite(v_);
    if (ok) {
      ++nv;
      const float* P = sP + 12 * v;
      const double s = weighted ? (double)wt : 1.0;
      double ra[4], rb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ra[i] = s * ((double)u_ * (double)P[8 + i] - (double)P[i]);
        rb[i] = s * ((double)v_ * (double)P[8 + i] - (double)P[4 + i]);
      }
      const double aj = lane == 0 ? ra[0] : lane == 1 ? ra[1] : lane == 2 ? ra[2] : ra[3];
      const double bj = lane == 0 ? rb[0] : lane == 1 ? rb[1] : lane == 2 ? rb[2] : rb[3];
#pragma unroll
      for (int i = 0; i < 4; ++i) Mc[i] += ra[i] * aj + rb[i] * bj;
    }
  }
  // trace normalisation
  double diag = lane == 0 ? Mc[0] : lane == 1 ? Mc[1] : lane == 2 ? Mc[2] : Mc[3];
  double tr = diag + xor4(diag, 1);
  tr += xor4(tr, 2);
  const double inv_tr = tr > 0.0 ? 1.0 / tr : 0.0;
  double Md[4];
  float A[4], Vc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Md[i] = Mc[i] * inv_tr;
    A[i] = (float)Md[i];
    Vc[i] = (i == lane) ? 1.0f : 0.0f;
  }

  // ---- cyclic Jacobi, FP32, parallel ordering (0,1)(2,3) (0,2)(1,3) (0,3)(1,2) ---------
#pragma unroll 1
  for (int sweep = 0; sweep < kDltSweeps; ++sweep) {
    jacobi_rotate<0, 1>(A, Vc, lane);
    jacobi_rotate<2, 3>(A, Vc, lane);
    jacobi_rotate<0, 2>(A, Vc, lane);
    jacobi_rotate<1, 3>(A, Vc, lane);
    jacobi_rotate<0, 3>(A, Vc, lane);
    jacobi_rotate<1, 2>(A, Vc, lane);
  }
  const float lam = lane == 0 ? A[0] : lane == 1 ? A[1] : lane == 2 ? A[2] : A[3];
  // arg-min over the 4 lanes (ties to the lower lane)
  float lmin = lam;
  int imin = lane;
#pragma unroll
  for (int m = 1; m < 4; m <<= 1) {
    const float ol = xor4(lmin, m);
    const int oi = __shfl_xor_sync(0xffffffffu, imin, m, 4);
    if (ol < lmin || (ol == lmin && oi < imin)) {
      lmin = ol;
      imin = oi;
    }
  }

  // ---- FP64 first-order refinement of the smallest eigenvector ---------------------------
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = (double)shfl4(Vc[i], imin);
  const double xj = lane == 0 ? x[0] : lane == 1 ? x[1] : lane == 2 ? x[2] : x[3];
  double Mx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double t = Md[i] * xj;  // column `lane` of M times x[lane]
    t += xor4(t, 1);
    t += xor4(t, 2);
    Mx[i] = t;
  }
  const double xx = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
  const double lam0 = (x[0] * Mx[0] + x[1] * Mx[1] + x[2] * Mx[2] + x[3] * Mx[3]) / xx;
  double coef = 0.0;
  if (lane != imin) {
    double vr = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) vr += (double)Vc[i] * (Mx[i] - lam0 * x[i]);
    const double den = lam0 - (double)lam;
    if (fabs(den) > 1e-12) coef = vr / den;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double t = coef * (double)Vc[i];
    t += xor4(t, 1);
    t += xor4(t, 2);
    x[i] += t;
  }

  const bool good = (nv >= 2) && (tr > 0.0) && (x[3] != 0.0);
  const float qnan = __int_as_float(0x7fc00000);
  float Xp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) Xp[i] = good ? (float)(x[i] / x[3]) : qnan;

  // ---- reprojection residual: lane j takes views j, j+4, ... -----------------------------
  float e2 = 0.f;
  if (resid) {
    for (int v = lane; v < V; v += 4) {
      const int64_t o = (b * V + v) * K + k;
      const float u_ = kp[2 * o], v_ = kp[2 * o + 1];
      const float wt = w ? w[o] : 1.0f;
      if ((wt >= min_weight) && isfinite(u_) && isfinite(v_)) {
        const float* P = sP + 12 * v;
        const float hx = P[0] * Xp[0] + P[1] * Xp[1] + P[2] * Xp[2] + P[3];
        const float hy = P[4] * Xp[0] + P[5] * Xp[1] + P[6] * Xp[2] + P[7];
        const float hz = P[8] * Xp[0] + P[9] * Xp[1] + P[10] * Xp[2] + P[11];
        const float du = hx / hz - u_, dv = hy / hz - v_;
        e2 += du * du + dv * dv;
      }
    }
    e2 += xor4(e2, 1);
    e2 += xor4(e2, 2);
  }
  if (active) {
    if (lane < 3) X[3 * pid + lane] = Xp[lane];
    if (lane == 3) {
      if (resid) resid[pid] = good ? sqrtf(e2 / (float)nv) : qnan;
      if (n_views) n_views[pid] = nv;
    }
  }
}

}  // namespace mvgeo
