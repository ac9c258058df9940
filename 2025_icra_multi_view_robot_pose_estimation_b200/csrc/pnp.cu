// pnp.cu — batched camera-pose refinement (PnP) per (frame, view), sm_100a.
//
// "Next" row 3 of the scope table: the step right after the hot path in the reference,
// estimate_camera_pose (model/Fr5_model_train.ipynb:4707-4753, model/Franka_research3_model_train.ipynb:
// 3667-3708): FK points (3-D) + decoded key-points (2-D, score >= threshold, at least 4 of them)
// -> cv2.solvePnPRansac(EPNP) -> plausibility gate 0.5 < |t| < 5 m (Franka_research3_model_train.ipynb:
// 3696-3701) -> on failure fall back to the ArUco prior (Fr5_model_train.ipynb:4978-4993).
// Here: Levenberg-Marquardt on the reprojection error, started from the prior pose stored in the
// camera record (so "failure" degrades to the prior exactly like the reference), one thread per
// (frame, view). Specification / oracle: oracle.mvgeo_oracle.pnp_refine (float64 LM, cross-checked
// against cv2.solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess=True), which minimises the same cost).
//
// Pose update is a left perturbation R <- exp([dw]x) R, t <- t + dt, so with Y = R X the point
// Jacobian is d(Xc)/d(dw, dt) = [ -[Y]x | I ]. The 2x3 projection Jacobian (Brown-Conrady) is
// evaluated in FP32; the 6x6 normal equations are accumulated and solved (Cholesky) in FP64 —
// rotation and translation columns differ by orders of magnitude and K <= 9 points make this free.
// Roofline: FP32/FP64 latency (~1.5 kFLOP per point per iteration); bytes < 1 KB per frame-view.
#include "common.cuh"

namespace mvgeo {

constexpr int kPnpThreads = 128;
constexpr int kPnpMaxPts = 16;

struct PnpCam {
  float fx, fy, cx, cy, k1, k2, p1, p2, k3;
};

// (u, v) and the 2x3 Jacobian d(u,v)/d(Xc) at camera-frame point (xc, yc, zc)
__device__ __forceinline__ void pnp_project(const PnpCam& c, float xc, float yc, float zc, float& u, float& v, float* J) {
  const float iz = 1.0f / zc;
  const float xp = xc * iz, yp = yc * iz;
  const float r2 = xp * xp + yp * yp;
  const float rad = 1.0f + r2 * (c.k1 + r2 * (c.k2 + r2 * c.k3));
  u = c.fx * (xp * rad + 2.0f * c.p1 * xp * yp + c.p2 * (r2 + 2.0f * xp * xp)) + c.cx;
  v = c.fy * (yp * rad + c.p1 * (r2 + 2.0f * yp * yp) + 2.0f * c.p2 * xp * yp) + c.cy;
  const float dr = c.k1 + r2 * (2.0f * c.k2 + 3.0f * c.k3 * r2);
  const float a00 = rad + 2.0f * xp * xp * dr + 2.0f * c.p1 * yp + 6.0f * c.p2 * xp;
  const float a01 = 2.0f * xp * yp * dr + 2.0f * c.p1 * xp + 2.0f * c.p2 * yp;
  const float a11 = rad + 2.0f * yp * yp * dr + 6.0f * c.p1 * yp + 2.0f * c.p2 * xp;
  J[0] = c.fx * a00 * iz;
  J[1] = c.fx * a01 * iz;
  J[2] = -c.fx * (a00 * xp + a01 * yp) * iz;
  J[3] = c.fy * a01 * iz;
  J[4] = c.fy * a11 * iz;
  J[5] = -c.fy * (a01 * xp + a11 * yp) * iz;
}

// In-place Cholesky solve of the 6x6 SPD system A x = b (A row-major full). False if not SPD.
__device__ __forceinline__ bool chol_solve6(double (&A)[6][6], double (&b)[6]) {
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < j) d -= A[j][k] * A[j][k];
    if (!(d > 0.0)) return false;
    const double l = sqrt(d);
    A[j][j] = l;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i > j) {
        double s = A[i][j];
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k < j) s -= A[i][k] * A[j][k];
        A[i][j] = s / l;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {  // L y = b
    double s = b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < i) s -= A[i][k] * b[k];
    b[i] = s / A[i][i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {  // L^T x = y
    double s = b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k > i) s -= A[k][i] * b[k];
    b[i] = s / A[i][i];
  }
  return true;
}

// R <- exp([w]x) R (Rodrigues formula, FP64)
__device__ __forceinline__ void rot_update(double (&R)[9], const double* w) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double a, b;  // exp = I + a [w]x + b [w]x^2
  if (th < 1e-8) {
    a = 1.0 - th2 / 6.0;
    b = 0.5 - th2 / 24.0;
  } else {
    a = sin(th) / th;
    b = (1.0 - cos(th)) / th2;
  }
  const double K[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
  double E[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double k2 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) k2 += K[3 * i + k] * K[3 * k + j];
      E[3 * i + j] = (i == j ? 1.0 : 0.0) + a * K[3 * i + j] + b * k2;
    }
  double N[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) N[3 * i + j] = E[3 * i] * R[j] + E[3 * i + 1] * R[3 + j] + E[3 * i + 2] * R[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = N[i];
}

// rvec = log(R) (cv2.Rodrigues inverse)
__device__ __forceinline__ void rot_to_rvec(const double (&R)[9], float* rv) {
  const double c = fmin(1.0, fmax(-1.0, 0.5 * (R[0] + R[4] + R[8] - 1.0)));
  const double th = acos(c);
  double ax = R[7] - R[5], ay = R[2] - R[6], az = R[3] - R[1];
  const double s = 0.5 * sqrt(ax * ax + ay * ay + az * az);  // sin(theta)
  double k;
  if (s < 1e-10) {
    if (c > 0.0) {
      k = 0.5;  // theta ~ 0: rvec ~ 0.5 * (R - R^T) vee
    } else {     // theta ~ pi: axis from the diagonal
      const double xx = sqrt(fmax(0.0, 0.5 * (R[0] + 1.0))), yy = sqrt(fmax(0.0, 0.5 * (R[4] + 1.0))),
                   zz = sqrt(fmax(0.0, 0.5 * (R[8] + 1.0)));
      ax = xx;
      ay = (R[1] + R[3] < 0.0) ? -yy : yy;
      az = (R[2] + R[6] < 0.0) ? -zz : zz;
      k = th;
    }
  } else {
    k = 0.5 * th / s;
  }
  rv[0] = (float)(k * ax);
  rv[1] = (float)(k * ay);
  rv[2] = (float)(k * az);
}

// Levenberg-Marquardt on the reprojection error of the points in `valid` (n of them), from (R, t) in place.
// Returns the status bits (1 solved, 2 converged, 4 plausible |t|); `cost` = final sum of squared residuals.
// prior_R / prior_t: pose handed back when the cost is not finite (failed solve).
__device__ __forceinline__ int pnp_lm(const PnpCam& c, const float* __restrict__ Xp, const float* __restrict__ kpp,
                                      int K, uint32_t valid, int max_iters, double (&R)[9], double (&t)[3],
                                      double& cost) {
  const double R0[9] = {R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8]};
  const double t0[3] = {t[0], t[1], t[2]};
  int st = 1;
  auto eval_cost = [&](const double (&Rr)[9], const double (&tt)[3]) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) {
      if (!((valid >> k) & 1u)) continue;
      const float x = Xp[3 * k], y = Xp[3 * k + 1], z = Xp[3 * k + 2];
      const float xc = (float)(Rr[0] * x + Rr[1] * y + Rr[2] * z + tt[0]);
      const float yc = (float)(Rr[3] * x + Rr[4] * y + Rr[5] * z + tt[1]);
      const float zc = (float)(Rr[6] * x + Rr[7] * y + Rr[8] * z + tt[2]);
      float u, vv, J[6];
      pnp_project(c, xc, yc, zc, u, vv, J);
      const double du = (double)u - (double)kpp[2 * k], dv = (double)vv - (double)kpp[2 * k + 1];
      s += du * du + dv * dv;
    }
    return s;
  };
  double lambda = 1e-3;
  cost = eval_cost(R, t);
  for (int it = 0; it < max_iters; ++it) {
    double H[6][6], g[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      g[r] = 0.0;
#pragma unroll
      for (int q = 0; q < 6; ++q) H[r][q] = 0.0;
    }
    for (int k = 0; k < K; ++k) {
      if (!((valid >> k) & 1u)) continue;
      const float x = Xp[3 * k], y = Xp[3 * k + 1], z = Xp[3 * k + 2];
      const float Yx = (float)(R[0] * x + R[1] * y + R[2] * z), Yy = (float)(R[3] * x + R[4] * y + R[5] * z),
                  Yz = (float)(R[6] * x + R[7] * y + R[8] * z);
      float u, vv, J[6];
      pnp_project(c, Yx + (float)t[0], Yy + (float)t[1], Yz + (float)t[2], u, vv, J);
      const double ru = (double)u - (double)kpp[2 * k], rv = (double)vv - (double)kpp[2 * k + 1];
      // rows of the 2x6 Jacobian: [ J3 * (-[Y]x) | J3 ],  -[Y]x = [[0, Yz, -Yy], [-Yz, 0, Yx], [Yy, -Yx, 0]]
      double ju[6], jv[6];
      ju[0] = -J[1] * Yz + J[2] * Yy;
      ju[1] = J[0] * Yz - J[2] * Yx;
      ju[2] = -J[0] * Yy + J[1] * Yx;
      ju[3] = J[0];
      ju[4] = J[1];
      ju[5] = J[2];
      jv[0] = -J[4] * Yz + J[5] * Yy;
      jv[1] = J[3] * Yz - J[5] * Yx;
      jv[2] = -J[3] * Yy + J[4] * Yx;
      jv[3] = J[3];
      jv[4] = J[4];
      jv[5] = J[5];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        g[r] += ju[r] * ru + jv[r] * rv;
#pragma unroll
        for (int q = 0; q < 6; ++q)
          if (q <= r) H[r][q] += ju[r] * ju[q] + jv[r] * jv[q];
      }
    }
    bool accepted = false;
    for (int tries = 0; tries < 6 && !accepted; ++tries) {
      double A[6][6], d[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        d[r] = -g[r];
#pragma unroll
        for (int q = 0; q < 6; ++q) A[r][q] = q <= r ? H[r][q] : H[q][r];
        A[r][r] += lambda * H[r][r] + 1e-12;
      }
      if (chol_solve6(A, d)) {
        double Rn[9], tn[3];
#pragma unroll
        for (int j = 0; j < 9; ++j) Rn[j] = R[j];
        rot_update(Rn, d);
        tn[0] = t[0] + d[3];
        tn[1] = t[1] + d[4];
        tn[2] = t[2] + d[5];
        const double cn = eval_cost(Rn, tn);
        if (cn <= cost) {
          const double step2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5];
          const bool small = (cost - cn) <= 1e-10 * cost + 1e-12 || step2 < 1e-18;
#pragma unroll
          for (int j = 0; j < 9; ++j) R[j] = Rn[j];
          t[0] = tn[0];
          t[1] = tn[1];
          t[2] = tn[2];
          cost = cn;
          lambda = fmax(lambda * 0.1, 1e-9);
          accepted = true;
          if (small) {
            st |= 2;
            it = max_iters;
          }
          continue;
        }
      }
      lambda *= 10.0;
    }
    if (!accepted) {  // no downhill step at any damping: at a minimum to working precision
      if (isfinite(cost)) st |= 2;  // a non-finite cost (prior puts a point on the camera plane) is not convergence
      break;
    }
  }
  if (!isfinite(cost)) {  // failed solve: hand back the starting pose, as the header promises
    st &= ~2;
#pragma unroll
    for (int j = 0; j < 9; ++j) R[j] = R0[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) t[j] = t0[j];
  }
  const double tn2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
  if (tn2 > 0.25 && tn2 < 25.0) st |= 4;  // 0.5 m < |t| < 5 m (Franka_research3_model_train.ipynb:3696-3701)
  return st;
}

// which of the K points take part: weight passes, image AND object coordinates finite
// (X_tri is NaN for under-observed key-points: one NaN would turn every cost comparison false)
__device__ __forceinline__ uint32_t pnp_valid_mask(const float* __restrict__ Xp, const float* __restrict__ kpp,
                                                   const float* __restrict__ wp, int K, float min_weight, int& n) {
  uint32_t valid = 0;
  n = 0;
  for (int k = 0; k < K; ++k) {
    const float wt = wp ? wp[k] : 1.0f;
    if ((wt >= min_weight) && isfinite(kpp[2 * k]) && isfinite(kpp[2 * k + 1]) && isfinite(Xp[3 * k]) &&
        isfinite(Xp[3 * k + 1]) && isfinite(Xp[3 * k + 2])) {
      valid |= 1u << k;
      ++n;
    }
  }
  return valid;
}

__global__ void __launch_bounds__(kPnpThreads)
    pnp_refine_kernel(const float* __restrict__ X, int x_per_view, const float* __restrict__ kp,
                      const float* __restrict__ w, const mvgeo_camera* __restrict__ cams, int64_t B, int V, int K,
                      float min_weight, int max_iters, float* __restrict__ rvec, float* __restrict__ tvec,
                      float* __restrict__ rms, int32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * kPnpThreads + threadIdx.x;
  if (i >= B * V) return;
  const int v = (int)(i % V);
  const int64_t b = i / V;
  const mvgeo_camera& cam = cams[v];
  PnpCam c = {cam.fx, cam.fy, cam.cx, cam.cy, cam.dist[0], cam.dist[1], cam.dist[2], cam.dist[3], cam.dist[4]};
  double R[9], t[3];
#pragma unroll
  for (int j = 0; j < 9; ++j) R[j] = (double)cam.R[j];
#pragma unroll
  for (int j = 0; j < 3; ++j) t[j] = (double)cam.t[j];
  const float* Xp = X + (x_per_view ? (b * V + v) : b) * (int64_t)K * 3;
  const float* kpp = kp + (b * V + v) * (int64_t)K * 2;
  const float* wp = w ? w + (b * V + v) * (int64_t)K : nullptr;
  int n = 0;
  const uint32_t valid = pnp_valid_mask(Xp, kpp, wp, K, min_weight, n);
  int st = 0;
  double cost = 0.0;
  // the reference refuses PnP with fewer than 4 confident points (Fr5_model_train.ipynb:4728)
  if (n >= 4) st = pnp_lm(c, Xp, kpp, K, valid, max_iters, R, t, cost);
  rot_to_rvec(R, rvec + 3 * i);
  tvec[3 * i] = (float)t[0];
  tvec[3 * i + 1] = (float)t[1];
  tvec[3 * i + 2] = (float)t[2];
  if (rms) rms[i] = n >= 4 ? (float)sqrt(cost / n) : __int_as_float(0x7fc00000);
  if (status) status[i] = st;
}

// ----------------------------------------------------------------------------------------
// Pose WITHOUT a prior: replaces cv2.solvePnPRansac(obj, img, K, dist, flags=SOLVEPNP_EPNP) in
// estimate_camera_pose (model/Fr5_model_train.ipynb:4735-4741). With K <= 16 key-points random sampling is
// pointless: EVERY point triplet is a hypothesis. One WARP per (frame, view); lane l takes triplets
// l, l+32, ...: P3P (Grunert: a quartic in the depth ratio v = s3/s1, coefficients by polynomial
// arithmetic, roots by Durand-Kerner in double complex + Newton polish) -> up to 4 poses, each scored on
// all valid points (inliers within `thresh` px, then squared error of the inliers). The warp's best
// hypothesis (most inliers, then lowest error, then lowest triplet number: deterministic) is broadcast
// and refined by the same LM as above on its inliers. Specification: oracle.mvgeo_oracle.pnp_solve.
// ----------------------------------------------------------------------------------------
constexpr int kSolveWarps = 4;

struct Cplx {
  double re, im;
};
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cdiv(Cplx a, Cplx b) {
  const double d = b.re * b.re + b.im * b.im;
  return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// real roots > 0 of q[0] v^4 + q[1] v^3 + q[2] v^2 + q[3] v + q[4]; returns their number
__device__ __forceinline__ int quartic_positive_roots(const double (&q)[5], double (&roots)[4]) {
  double mx = 0.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) mx = fmax(mx, fabs(q[i]));
  if (!(fabs(q[0]) > 1e-13 * mx) || !isfinite(mx)) return 0;  // degenerate leading coefficient: skip the triplet
  double a[4];  // monic: v^4 + a[0] v^3 + a[1] v^2 + a[2] v + a[3]
  double bound = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = q[i + 1] / q[0];
    bound = fmax(bound, fabs(a[i]));
  }
  bound = 1.0 + bound;  // Cauchy bound
  Cplx z[4];
  {
    Cplx s = {0.4 * 0.9, 0.9 * 0.9}, cur = {1.0, 0.0};  // distinct, non-real, non-symmetric starting points
    const double rad = fmin(bound, 1e3);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      z[i] = {cur.re * rad * 0.5, cur.im * rad * 0.5};
      cur = cmul(cur, s);
    }
  }
  for (int it = 0; it < 60; ++it) {
    double moved = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Cplx pz = {1.0, 0.0};  // Horner
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pz = cmul(pz, z[i]);
        pz.re += a[k];
      }
      Cplx den = {1.0, 0.0};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j != i) den = cmul(den, csub(z[i], z[j]));
      const Cplx d = cdiv(pz, den);
      z[i] = csub(z[i], d);
      moved = fmax(moved, fabs(d.re) + fabs(d.im));
    }
    if (moved < 1e-14 * bound) break;
  }
  int n = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (!(z[i].re > 0.0) || fabs(z[i].im) > 1e-6 * (1.0 + fabs(z[i].re))) continue;
    double v = z[i].re;
    for (int it = 0; it < 3; ++it) {  // Newton polish on the real polynomial
      const double p = (((v + a[0]) * v + a[1]) * v + a[2]) * v + a[3];
      const double dp = ((4.0 * v + 3.0 * a[0]) * v + 2.0 * a[1]) * v + a[2];
      if (fabs(dp) > 1e-300) v -= p / dp;
    }
    if (v > 0.0 && isfinite(v)) roots[n++] = v;
  }
  return n;
}

struct PnpHyp {
  int inl;      // inlier count (-1: none)
  double err;   // squared error of the inliers
  int tri;      // triplet number (tie-break)
  uint32_t mask;
  double R[9], t[3];
};

__global__ void __launch_bounds__(kSolveWarps * 32)
    pnp_solve_kernel(const float* __restrict__ X, int x_per_view, const float* __restrict__ kp,
                     const float* __restrict__ w, const mvgeo_camera* __restrict__ cams, int64_t B, int V, int K,
                     float min_weight, float thresh, int max_iters, float* __restrict__ rvec,
                     float* __restrict__ tvec, float* __restrict__ rms, int32_t* __restrict__ status,
                     int32_t* __restrict__ inliers) {
  __shared__ double sf[kSolveWarps][kPnpMaxPts][3];  // unit bearings of the undistorted key-points
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kSolveWarps + warp;
  if (i >= B * V) return;  // whole warps leave together
  const int v = (int)(i % V);
  const int64_t b = i / V;
  const mvgeo_camera& cam = cams[v];
  const PnpCam c = {cam.fx, cam.fy, cam.cx, cam.cy, cam.dist[0], cam.dist[1], cam.dist[2], cam.dist[3], cam.dist[4]};
  const float* Xp = X + (x_per_view ? (b * V + v) : b) * (int64_t)K * 3;
  const float* kpp = kp + (b * V + v) * (int64_t)K * 2;
  const float* wp = w ? w + (b * V + v) * (int64_t)K : nullptr;
  int n = 0;
  const uint32_t valid = pnp_valid_mask(Xp, kpp, wp, K, min_weight, n);
  if (lane < K && ((valid >> lane) & 1u)) {
    // cv2.undistortPoints (5 fixed-point iterations) in double, then the unit bearing
    const double x0 = ((double)kpp[2 * lane] - c.cx) / c.fx, y0 = ((double)kpp[2 * lane + 1] - c.cy) / c.fy;
    double x = x0, y = y0;
    for (int it = 0; it < 5; ++it) {
      const double r2 = x * x + y * y;
      const double icd = 1.0 / (1.0 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2);
      const double dx = 2.0 * c.p1 * x * y + c.p2 * (r2 + 2.0 * x * x);
      const double dy = c.p1 * (r2 + 2.0 * y * y) + 2.0 * c.p2 * x * y;
      x = (x0 - dx) * icd;
      y = (y0 - dy) * icd;
    }
    const double inv = rsqrt(x * x + y * y + 1.0);
    sf[warp][lane][0] = x * inv;
    sf[warp][lane][1] = y * inv;
    sf[warp][lane][2] = inv;
  }
  __syncwarp();

  PnpHyp best;
  best.inl = -1;
  best.err = 0.0;
  best.tri = 0x7fffffff;
  best.mask = 0;
  const double th2 = (double)thresh * (double)thresh;
  if (n >= 4) {
    int tri = 0;
    for (int i0 = 0; i0 < K; ++i0) {
      if (!((valid >> i0) & 1u)) continue;
      for (int i1 = i0 + 1; i1 < K; ++i1) {
        if (!((valid >> i1) & 1u)) continue;
        for (int i2 = i1 + 1; i2 < K; ++i2) {
          if (!((valid >> i2) & 1u)) continue;
          const int my = tri++;
          if ((my & 31) != lane) continue;
          const int id[3] = {i0, i1, i2};
          double P[3][3], f[3][3];
#pragma unroll
          for (int a_ = 0; a_ < 3; ++a_)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              P[a_][d] = (double)Xp[3 * id[a_] + d];
              f[a_][d] = sf[warp][id[a_]][d];
            }
          auto d2 = [&](int a_, int b_) {
            const double x = P[a_][0] - P[b_][0], y = P[a_][1] - P[b_][1], z = P[a_][2] - P[b_][2];
            return x * x + y * y + z * z;
          };
          auto dotf = [&](int a_, int b_) { return f[a_][0] * f[b_][0] + f[a_][1] * f[b_][1] + f[a_][2] * f[b_][2]; };
          const double a2 = d2(1, 2), b2 = d2(0, 2), c2 = d2(0, 1);
          if (fmin(a2, fmin(b2, c2)) < 1e-18) continue;
          const double ca = dotf(1, 2), cb = dotf(0, 2), cg = dotf(0, 1);
          const double k = (a2 - c2) / b2, cr = c2 / b2;
          // u = N(v) / D(v); quartic = D^2 + N^2 - 2 cg N D - cr (1 + v^2 - 2 cb v) D^2
          const double N2 = k - 1.0, N1 = -2.0 * k * cb, N0 = 1.0 + k;
          const double D1 = -2.0 * ca, D0 = 2.0 * cg;
          const double DD2 = D1 * D1, DD1 = 2.0 * D1 * D0, DD0 = D0 * D0;
          double q[5];
          q[0] = N2 * N2 - cr * DD2;
          q[1] = 2.0 * N2 * N1 - 2.0 * cg * (N2 * D1) - cr * (DD1 - 2.0 * cb * DD2);
          q[2] = DD2 + 2.0 * N2 * N0 + N1 * N1 - 2.0 * cg * (N2 * D0 + N1 * D1) - cr * (DD0 - 2.0 * cb * DD1 + DD2);
          q[3] = DD1 + 2.0 * N1 * N0 - 2.0 * cg * (N1 * D0 + N0 * D1) - cr * (DD1 - 2.0 * cb * DD0);
          q[4] = DD0 + N0 * N0 - 2.0 * cg * (N0 * D0) - cr * DD0;
          double roots[4];
          const int nr = quartic_positive_roots(q, roots);
          for (int r_ = 0; r_ < nr; ++r_) {
            const double vv = roots[r_];
            const double den = 2.0 * (cg - vv * ca);
            if (fabs(den) < 1e-12) continue;
            const double uu = (N2 * vv * vv + N1 * vv + N0) / den;
            const double ww = 1.0 + vv * vv - 2.0 * vv * cb;
            if (!(uu > 0.0) || !(ww > 0.0)) continue;
            const double s1 = sqrt(b2 / ww);
            const double sc[3] = {s1, uu * s1, vv * s1};
            double Q[3][3];
#pragma unroll
            for (int a_ = 0; a_ < 3; ++a_)
#pragma unroll
              for (int d = 0; d < 3; ++d) Q[a_][d] = sc[a_] * f[a_][d];
            // orthonormal frames on both triangles: columns e1, e3 x e1, e3
            double Fp[3][3], Fq[3][3];
            bool okf = true;
            auto frame = [&](const double (&A)[3][3], double (&F)[3][3]) {
              double e1[3] = {A[1][0] - A[0][0], A[1][1] - A[0][1], A[1][2] - A[0][2]};
              const double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
              const double g2[3] = {A[2][0] - A[0][0], A[2][1] - A[0][1], A[2][2] - A[0][2]};
              for (int d = 0; d < 3; ++d) e1[d] /= n1;
              double e3[3] = {e1[1] * g2[2] - e1[2] * g2[1], e1[2] * g2[0] - e1[0] * g2[2], e1[0] * g2[1] - e1[1] * g2[0]};
              const double n3 = sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
              if (!(n3 > 1e-12) || !(n1 > 0.0)) {
                okf = false;
                return;
              }
              for (int d = 0; d < 3; ++d) e3[d] /= n3;
              const double e2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
              for (int d = 0; d < 3; ++d) {
                F[d][0] = e1[d];
                F[d][1] = e2[d];
                F[d][2] = e3[d];
              }
            };
            frame(P, Fp);
            frame(Q, Fq);
            if (!okf) continue;
            double Rh[9], th[3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int cc = 0; cc < 3; ++cc)
                Rh[3 * r + cc] = Fq[r][0] * Fp[cc][0] + Fq[r][1] * Fp[cc][1] + Fq[r][2] * Fp[cc][2];
#pragma unroll
            for (int r = 0; r < 3; ++r)
              th[r] = Q[0][r] - (Rh[3 * r] * P[0][0] + Rh[3 * r + 1] * P[0][1] + Rh[3 * r + 2] * P[0][2]);
            // score on every valid point
            int inl = 0;
            double err = 0.0;
            uint32_t mask = 0;
            bool front = true;
            for (int kk = 0; kk < K; ++kk) {
              if (!((valid >> kk) & 1u)) continue;
              const double x = Xp[3 * kk], y = Xp[3 * kk + 1], z = Xp[3 * kk + 2];
              const double zc = Rh[6] * x + Rh[7] * y + Rh[8] * z + th[2];
              if (!(zc > 1e-9)) {
                front = false;
                break;
              }
              float u, v2, J[6];
              pnp_project(c, (float)(Rh[0] * x + Rh[1] * y + Rh[2] * z + th[0]),
                          (float)(Rh[3] * x + Rh[4] * y + Rh[5] * z + th[1]), (float)zc, u, v2, J);
              const double du = (double)u - (double)kpp[2 * kk], dv = (double)v2 - (double)kpp[2 * kk + 1];
              const double e2 = du * du + dv * dv;
              if (e2 < th2) {
                ++inl;
                err += e2;
                mask |= 1u << kk;
              }
            }
            if (!front) continue;
            if (inl > best.inl || (inl == best.inl && err < best.err)) {
              best.inl = inl;
              best.err = err;
              best.tri = my;
              best.mask = mask;
#pragma unroll
              for (int j = 0; j < 9; ++j) best.R[j] = Rh[j];
#pragma unroll
              for (int j = 0; j < 3; ++j) best.t[j] = th[j];
            }
          }
        }
      }
    }
  }
  // warp arg-best: (inliers desc, error asc, triplet asc)
  int win = lane;
  {
    int inl = best.inl, tri = best.tri;
    double err = best.err;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const int oinl = __shfl_xor_sync(0xffffffffu, inl, off);
      const double oerr = __shfl_xor_sync(0xffffffffu, err, off);
      const int otri = __shfl_xor_sync(0xffffffffu, tri, off);
      const int owin = __shfl_xor_sync(0xffffffffu, win, off);
      const bool better = oinl > inl || (oinl == inl && (oerr < err || (oerr == err && otri < tri)));
      if (better) {
        inl = oinl;
        err = oerr;
        tri = otri;
        win = owin;
      }
    }
  }
  double R[9], t[3];
#pragma unroll
  for (int j = 0; j < 9; ++j) R[j] = __shfl_sync(0xffffffffu, best.R[j], win);
#pragma unroll
  for (int j = 0; j < 3; ++j) t[j] = __shfl_sync(0xffffffffu, best.t[j], win);
  const int inl = __shfl_sync(0xffffffffu, best.inl, win);
  const uint32_t mask = __shfl_sync(0xffffffffu, best.mask, win);
  if (lane != 0) return;
  int st = 0;
  double cost = 0.0;
  const float nanf_ = __int_as_float(0x7fc00000);
  if (inl >= 4) {  // the reference requires `success and num_inliers >= 4` (Fr5_model_train.ipynb:4743)
    st = pnp_lm(c, Xp, kpp, K, mask, max_iters, R, t, cost);
    rot_to_rvec(R, rvec + 3 * i);
    tvec[3 * i] = (float)t[0];
    tvec[3 * i + 1] = (float)t[1];
    tvec[3 * i + 2] = (float)t[2];
  } else {  // the reference returns (None, None): NaN pose, status 0
    for (int j = 0; j < 3; ++j) rvec[3 * i + j] = tvec[3 * i + j] = nanf_;
  }
  if (rms) rms[i] = inl >= 4 ? (float)sqrt(cost / inl) : nanf_;
  if (status) status[i] = st;
  if (inliers) inliers[i] = inl >= 4 ? (int32_t)mask : 0;
}

}  // namespace mvgeo

using namespace mvgeo;

extern "C" int mvgeo_pnp_refine(const float* X, int x_per_view, const float* kp, const float* w,
                                const mvgeo_camera* cams, int64_t B, int V, int K, float min_weight, int max_iters,
                                float* rvec, float* tvec, float* rms, int32_t* status, void* stream) {
  if (B < 0 || V < 1 || V > MVGEO_MAX_VIEWS || K < 1 || K > kPnpMaxPts || max_iters < 0 || max_iters > 100)
    return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!X || !kp || !cams || !rvec || !tvec) return MVGEO_ENULL;
  const int64_t n = B * V;
  const unsigned grid = (unsigned)((n + kPnpThreads - 1) / kPnpThreads);
  pnp_refine_kernel<<<grid, kPnpThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      X, x_per_view, kp, w, cams, B, V, K, min_weight, max_iters, rvec, tvec, rms, status);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

extern "C" int mvgeo_pnp_solve(const float* X, int x_per_view, const float* kp, const float* w,
                               const mvgeo_camera* cams, int64_t B, int V, int K, float min_weight,
                               float reproj_thresh, int max_iters, float* rvec, float* tvec, float* rms,
                               int32_t* status, int32_t* inliers, void* stream) {
  if (B < 0 || V < 1 || V > MVGEO_MAX_VIEWS || K < 1 || K > kPnpMaxPts || max_iters < 0 || max_iters > 100 ||
      !(reproj_thresh > 0.f))
    return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!X || !kp || !cams || !rvec || !tvec) return MVGEO_ENULL;
  const int64_t n = B * V;
  const unsigned grid = (unsigned)((n + kSolveWarps - 1) / kSolveWarps);
  pnp_solve_kernel<<<grid, kSolveWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      X, x_per_view, kp, w, cams, B, V, K, min_weight, reproj_thresh, max_iters, rvec, tvec, rms, status, inliers);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}
