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

  uint32_t valid = 0;
  int n = 0;
  for (int k = 0; k < K; ++k) {
    const float wt = wp ? wp[k] : 1.0f;
    // a point takes part only if its weight passes, and BOTH its image and object coordinates are finite
    // (X_tri is NaN for under-observed key-points: one NaN would turn every cost comparison false)
    if ((wt >= min_weight) && isfinite(kpp[2 * k]) && isfinite(kpp[2 * k + 1]) && isfinite(Xp[3 * k]) &&
        isfinite(Xp[3 * k + 1]) && isfinite(Xp[3 * k + 2])) {
      valid |= 1u << k;
      ++n;
    }
  }
  int st = 0;
  double cost = 0.0;
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
  if (n >= 4) {  // the reference refuses PnP with fewer than 4 confident points (Fr5_model_train.ipynb:4728)
    st |= 1;
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
    if (!isfinite(cost)) {  // failed solve: hand back the prior, as the header promises
      st &= ~2;
#pragma unroll
      for (int j = 0; j < 9; ++j) R[j] = (double)cam.R[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) t[j] = (double)cam.t[j];
    }
    const double tn2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
    if (tn2 > 0.25 && tn2 < 25.0) st |= 4;  // 0.5 m < |t| < 5 m (Franka_research3_model_train.ipynb:3696-3701)
  }
  rot_to_rvec(R, rvec + 3 * i);
  tvec[3 * i] = (float)t[0];
  tvec[3 * i + 1] = (float)t[1];
  tvec[3 * i + 2] = (float)t[2];
  if (rms) rms[i] = n >= 4 ? (float)sqrt(cost / n) : __int_as_float(0x7fc00000);
  if (status) status[i] = st;
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
