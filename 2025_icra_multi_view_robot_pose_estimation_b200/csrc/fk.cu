// fk.cu — DH forward kinematics, pinhole/Brown-Conrady projection and the FK-consistency
// (reprojection) loss with a hand-written backward (kernel 3 of the hot path), sm_100a.
//
// Replaces angle_to_joint_coordinate (FR3 model/MvRoPose_FR3.py:90-131, Fr5
// model/Fr5_model_train.ipynb:256-288), forward_kinematics (Meca500
// visualization/Meca500_vis.ipynb:62-82), ForwardKinematics.forward (model/MV-model.ipynb:858-874),
// joint_coordinate_to_pixel_plane / project_to_pixel / project_3d_to_2d (cv2.projectPoints call
// sites, model/MvRoPose_FR3.py:133-141 and twins) and the FK term of robot_pose_loss
// (model/MV-model.ipynb:942-950). The reference has no backward; the gradient here is analytic.
//
// Roofline: FP32 latency. ~0.6 kFLOP + J sincos per frame for the chain and ~40 FLOP per
// projected point; bytes are < 1 KB per frame. One thread owns one (frame, view): the chain is a serial
// product of J 3x4 transforms held in registers, the J joint axes needed by the backward are
// kept alongside, and per-frame results are written without atomics (deterministic).
//
// Geometry of the backward. For a revolute joint i with axis z_i through the point o_i (both in
// the base frame), d p_k / d theta_i = z_i x (p_k - o_i) for every chain point k at or beyond the
// joint. Standard DH: axis and point are the z column / origin of the frame BEFORE link i.
// Modified (Craig) DH: they are the z column / origin of the frame AFTER link i (Rz and Tz leave
// the z axis in place). With g_k = d loss / d p_k the joint gradient collapses to
//     d loss / d theta_i = z_i . ( sum_{k>=i} p_k x g_k  -  o_i x sum_{k>=i} g_k ),
// two suffix sums accumulated from the end of the chain.
#include "common.cuh"

namespace mvgeo {

constexpr int kFkThreads = 128;
constexpr int kMaxPts = MVGEO_MAX_JOINTS + 1;

struct Vec3 {
  float x, y, z;
};
__device__ __forceinline__ Vec3 cross(const Vec3& a, const Vec3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// Serial DH chain in the base frame. pts[0..K) are the emitted points; when AXES, axis[i] /
// apt[i] describe joint i for the backward and first_pt[i] is the first emitted point that
// moves with joint i.
template <bool AXES, bool BASE>
__device__ __forceinline__ void chain_forward(const mvgeo_chain& ch, const float* __restrict__ q, Vec3* pts,
                                              Vec3* axis, Vec3* apt) {
  // T = [r0 r1 r2 | p], columns of the rotation kept as three vectors. BASE (= emit_base) is a
  // template parameter so that every pts[] index is a compile-time constant (registers, no stack).
  Vec3 cx = {1.f, 0.f, 0.f}, cy = {0.f, 1.f, 0.f}, cz = {0.f, 0.f, 1.f}, p = {0.f, 0.f, 0.f};
  if (BASE) pts[0] = p;
#pragma unroll
  for (int i = 0; i < MVGEO_MAX_JOINTS; ++i) {
    if (i < ch.n_joints) {
      const float th = (q[i] + ch.theta_offset[i]) * ch.angle_scale;
      float st, ct;
      sincosf(th, &st, &ct);
      const float ca = ch.cos_alpha[i], sa = ch.sin_alpha[i], a = ch.a[i], d = ch.d[i];
      if (ch.convention == MVGEO_DH_STANDARD) {
        if (AXES) {
          axis[i] = cz;
          apt[i] = p;
        }
        // columns of T_i: (ct, st, 0), (-st ca, ct ca, sa), (st sa, -ct sa, ca), (a ct, a st, d)
        const Vec3 nx = {cx.x * ct + cy.x * st, cx.y * ct + cy.y * st, cx.z * ct + cy.z * st};
        const Vec3 ty = {cy.x * ct - cx.x * st, cy.y * ct - cx.y * st, cy.z * ct - cx.z * st};  // Rz(theta) e_y image
        const Vec3 ny = {ty.x * ca + cz.x * sa, ty.y * ca + cz.y * sa, ty.z * ca + cz.z * sa};
        const Vec3 nz = {cz.x * ca - ty.x * sa, cz.y * ca - ty.y * sa, cz.z * ca - ty.z * sa};
        p = {p.x + a * nx.x + d * cz.x, p.y + a * nx.y + d * cz.y, p.z + a * nx.z + d * cz.z};
        cx = nx;
        cy = ny;
        cz = nz;
      } else {
        // Craig: columns of T_i: (ct, st ca, st sa), (-st, ct ca, ct sa), (0, -sa, ca), (a, -d sa, d ca)
        const Vec3 ry = {cy.x * ca + cz.x * sa, cy.y * ca + cz.y * sa, cy.z * ca + cz.z * sa};  // Rx(alpha) e_y image
        const Vec3 nz = {cz.x * ca - cy.x * sa, cz.y * ca - cy.y * sa, cz.z * ca - cy.z * sa};
        const Vec3 nx = {cx.x * ct + ry.x * st, cx.y * ct + ry.y * st, cx.z * ct + ry.z * st};
        const Vec3 ny = {ry.x * ct - cx.x * st, ry.y * ct - cx.y * st, ry.z * ct - cx.z * st};
        p = {p.x + a * cx.x + d * nz.x, p.y + a * cx.y + d * nz.y, p.z + a * cx.z + d * nz.z};
        cx = nx;
        cy = ny;
        cz = nz;
        if (AXES) {
          axis[i] = cz;
          apt[i] = p;
        }
      }
      pts[i + (BASE ? 1 : 0)] = p;
    }
  }
}

struct CamRegs {
  float M[9];  // R_cam * R_view
  float t[3];
  float fx, fy, cx, cy, k1, k2, p1, p2, k3;
};

__device__ __forceinline__ CamRegs load_cam(const mvgeo_camera* __restrict__ cams, const float* __restrict__ R_view,
                                            int v) {
  CamRegs c;
  const float* R = cams[v].R;
  if (R_view) {
    const float* Rv = R_view + 9 * v;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        c.M[3 * i + j] = R[3 * i] * Rv[j] + R[3 * i + 1] * Rv[3 + j] + R[3 * i + 2] * Rv[6 + j];
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) c.M[i] = R[i];
  }
  c.t[0] = cams[v].t[0];
  c.t[1] = cams[v].t[1];
  c.t[2] = cams[v].t[2];
  c.fx = cams[v].fx;
  c.fy = cams[v].fy;
  c.cx = cams[v].cx;
  c.cy = cams[v].cy;
  c.k1 = cams[v].dist[0];
  c.k2 = cams[v].dist[1];
  c.p1 = cams[v].dist[2];
  c.p2 = cams[v].dist[3];
  c.k3 = cams[v].dist[4];
  return c;
}

// cv2.projectPoints: x = K * distort((M X + t) / z). When JAC, also returns d(u,v)/d(X) (2x3).
template <bool JAC>
__device__ __forceinline__ void project_point(const CamRegs& c, const Vec3& X, float& u, float& v, float* J) {
  const float xc = c.M[0] * X.x + c.M[1] * X.y + c.M[2] * X.z + c.t[0];
  const float yc = c.M[3] * X.x + c.M[4] * X.y + c.M[5] * X.z + c.t[1];
  const float zc = c.M[6] * X.x + c.M[7] * X.y + c.M[8] * X.z + c.t[2];
  const float iz = 1.0f / zc;
  const float xp = xc * iz, yp = yc * iz;
  const float r2 = xp * xp + yp * yp;
  const float rad = 1.0f + r2 * (c.k1 + r2 * (c.k2 + r2 * c.k3));
  const float xpp = xp * rad + 2.0f * c.p1 * xp * yp + c.p2 * (r2 + 2.0f * xp * xp);
  const float ypp = yp * rad + c.p1 * (r2 + 2.0f * yp * yp) + 2.0f * c.p2 * xp * yp;
  u = c.fx * xpp + c.cx;
  v = c.fy * ypp + c.cy;
  if (JAC) {
    const float dr = c.k1 + r2 * (2.0f * c.k2 + 3.0f * c.k3 * r2);  // d rad / d r2
    const float a00 = rad + 2.0f * xp * xp * dr + 2.0f * c.p1 * yp + 6.0f * c.p2 * xp;
    const float a01 = 2.0f * xp * yp * dr + 2.0f * c.p1 * xp + 2.0f * c.p2 * yp;
    const float a10 = a01;
    const float a11 = rad + 2.0f * yp * yp * dr + 6.0f * c.p1 * yp + 2.0f * c.p2 * xp;
    // d(xp,yp)/d(xc,yc,zc) = [[iz,0,-xp iz],[0,iz,-yp iz]]
    const float du[3] = {c.fx * a00 * iz, c.fx * a01 * iz, -c.fx * (a00 * xp + a01 * yp) * iz};
    const float dv[3] = {c.fy * a10 * iz, c.fy * a11 * iz, -c.fy * (a10 * xp + a11 * yp) * iz};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J[j] = du[0] * c.M[j] + du[1] * c.M[3 + j] + du[2] * c.M[6 + j];
      J[3 + j] = dv[0] * c.M[j] + dv[1] * c.M[3 + j] + dv[2] * c.M[6 + j];
    }
  }
}

// ------------------------------------------------------------------------------ kernels
template <bool BASE>
__global__ void __launch_bounds__(kFkThreads) fk_kernel(const mvgeo_chain ch, const float* __restrict__ q, int64_t B,
                                                        const float* __restrict__ R_view, int V,
                                                        float* __restrict__ X) {
  const int64_t b = (int64_t)blockIdx.x * kFkThreads + threadIdx.x;
  if (b >= B) return;
  Vec3 pts[kMaxPts];
  chain_forward<false, BASE>(ch, q + b * ch.n_joints, pts, nullptr, nullptr);
  const int K = ch.n_joints + (BASE ? 1 : 0);
  for (int v = 0; v < V; ++v) {
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (R_view) {
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = R_view[9 * v + i];
    }
    float* out = X + ((b * V + v) * K) * 3;
#pragma unroll
    for (int k = 0; k < kMaxPts; ++k) {
      if (k < K) {
        out[3 * k + 0] = R[0] * pts[k].x + R[1] * pts[k].y + R[2] * pts[k].z;
        out[3 * k + 1] = R[3] * pts[k].x + R[4] * pts[k].y + R[5] * pts[k].z;
        out[3 * k + 2] = R[6] * pts[k].x + R[7] * pts[k].y + R[8] * pts[k].z;
      }
    }
  }
}

__global__ void __launch_bounds__(kFkThreads) project_kernel(const float* __restrict__ X, int x_per_view,
                                                             const mvgeo_camera* __restrict__ cams, int64_t B, int V,
                                                             int K, float* __restrict__ uv) {
  // one thread per (frame, view, point)
  const int64_t i = (int64_t)blockIdx.x * kFkThreads + threadIdx.x;
  if (i >= B * V * K) return;
  const int k = (int)(i % K);
  const int v = (int)((i / K) % V);
  const int64_t b = i / ((int64_t)K * V);
  const CamRegs c = load_cam(cams, nullptr, v);
  const float* xp = X + (x_per_view ? ((b * V + v) * K + k) : (b * K + k)) * 3;
  float u, w;
  project_point<false>(c, Vec3{xp[0], xp[1], xp[2]}, u, w, nullptr);
  uv[2 * i] = u;
  uv[2 * i + 1] = w;
}

// One thread per (frame, view): the chain is recomputed per view (J sincos, cheap) so that a
// small batch still fills the machine and the serial projection loop is K points, not V*K.
// blockDim.x = frames_per_cta * V; the per-frame sums over views go through shared memory in
// fixed view order (deterministic, no atomics).
template <bool BASE>
__global__ void __launch_bounds__(kFkThreads)
    fk_reproj_fwd_kernel(const mvgeo_chain ch, const float* __restrict__ q, int64_t B,
                         const float* __restrict__ R_view, const mvgeo_camera* __restrict__ cams, int V,
                         const float* __restrict__ gt_uv, const float* __restrict__ w, float scale,
                         float* __restrict__ X_out, float* __restrict__ uv_out, float* __restrict__ frame_loss) {
  __shared__ float part[kFkThreads];
  const int fpc = blockDim.x / V;
  const int fl = threadIdx.x / V, v = threadIdx.x - fl * V;
  const int64_t b = (int64_t)blockIdx.x * fpc + fl;
  float acc = 0.f;
  if (b < B) {
    Vec3 pts[kMaxPts];
    chain_forward<false, BASE>(ch, q + b * ch.n_joints, pts, nullptr, nullptr);
    const int K = ch.n_joints + (BASE ? 1 : 0);
    const CamRegs c = load_cam(cams, R_view, v);
    const int64_t base = (b * V + v) * K;
    float Rv[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (X_out && R_view) {
#pragma unroll
      for (int i = 0; i < 9; ++i) Rv[i] = R_view[9 * v + i];
    }
#pragma unroll
    for (int k = 0; k < kMaxPts; ++k) {
      if (k < K) {
        float u, vv;
        project_point<false>(c, pts[k], u, vv, nullptr);
        if (uv_out) {
          uv_out[2 * (base + k)] = u;
          uv_out[2 * (base + k) + 1] = vv;
        }
        if (X_out) {
          float* o = X_out + 3 * (base + k);
          o[0] = Rv[0] * pts[k].x + Rv[1] * pts[k].y + Rv[2] * pts[k].z;
          o[1] = Rv[3] * pts[k].x + Rv[4] * pts[k].y + Rv[5] * pts[k].z;
          o[2] = Rv[6] * pts[k].x + Rv[7] * pts[k].y + Rv[8] * pts[k].z;
        }
        if (gt_uv) {
          const float gu = gt_uv[2 * (base + k)], gv = gt_uv[2 * (base + k) + 1];
          const float wt = w ? w[base + k] : 1.0f;
          if (isfinite(gu) && isfinite(gv)) {
            const float du = u - gu, dv = vv - gv;
            acc += wt * (du * du + dv * dv);
          }
        }
      }
    }
  }
  if (frame_loss) {
    part[threadIdx.x] = acc;
    __syncthreads();
    if (b < B && v == 0) {
      float t = 0.f;
      for (int i = 0; i < V; ++i) t += part[fl * V + i];
      frame_loss[b] = t * scale;
    }
  }
}

template <bool BASE>
__global__ void __launch_bounds__(kFkThreads)
    fk_reproj_bwd_kernel(const mvgeo_chain ch, const float* __restrict__ q, int64_t B,
                         const float* __restrict__ R_view, const mvgeo_camera* __restrict__ cams, int V,
                         const float* __restrict__ gt_uv, const float* __restrict__ w, float scale,
                         const float* __restrict__ dloss, float* __restrict__ dq) {
  __shared__ float part[kFkThreads][MVGEO_MAX_JOINTS];
  const int fpc = blockDim.x / V;
  const int fl = threadIdx.x / V, v = threadIdx.x - fl * V;
  const int64_t b = (int64_t)blockIdx.x * fpc + fl;
  float dqv[MVGEO_MAX_JOINTS];
#pragma unroll
  for (int i = 0; i < MVGEO_MAX_JOINTS; ++i) dqv[i] = 0.f;
  if (b < B) {
    Vec3 pts[kMaxPts], axis[MVGEO_MAX_JOINTS], apt[MVGEO_MAX_JOINTS];
    chain_forward<true, BASE>(ch, q + b * ch.n_joints, pts, axis, apt);
    const int K = ch.n_joints + (BASE ? 1 : 0);
    const float up = (dloss ? dloss[0] : 1.0f) * scale * 2.0f;
    const CamRegs c = load_cam(cams, R_view, v);
    const int64_t base = (b * V + v) * K;
    // suffix sums from the end of the chain; point index of joint i is i + emit_base
    Vec3 G = {0.f, 0.f, 0.f}, N = {0.f, 0.f, 0.f};
    constexpr int off = BASE ? 1 : 0;
#pragma unroll
    for (int i = MVGEO_MAX_JOINTS - 1; i >= 0; --i) {
      if (i < ch.n_joints) {
        const int k = i + off;
        Vec3 g = {0.f, 0.f, 0.f};
        const float gu = gt_uv[2 * (base + k)], gv = gt_uv[2 * (base + k) + 1];
        if (isfinite(gu) && isfinite(gv)) {
          float u, vv, J[6];
          project_point<true>(c, pts[k], u, vv, J);
          const float wt = (w ? w[base + k] : 1.0f) * up;
          const float ru = wt * (u - gu), rv = wt * (vv - gv);
          g = {ru * J[0] + rv * J[3], ru * J[1] + rv * J[4], ru * J[2] + rv * J[5]};
        }
        G = {G.x + g.x, G.y + g.y, G.z + g.z};
        const Vec3 pxg = cross(pts[k], g);
        N = {N.x + pxg.x, N.y + pxg.y, N.z + pxg.z};
        const Vec3 oxG = cross(apt[i], G);
        const Vec3 m = {N.x - oxG.x, N.y - oxG.y, N.z - oxG.z};
        dqv[i] = dot(axis[i], m) * ch.angle_scale;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MVGEO_MAX_JOINTS; ++i) part[threadIdx.x][i] = dqv[i];
  __syncthreads();
  // threads (fl, j) with j < n_joints add the V per-view partials of joint j in fixed order
  for (int j = v; j < ch.n_joints; j += V) {
    if (b < B) {
      float t = 0.f;
      for (int i = 0; i < V; ++i) t += part[fl * V + i][j];
      dq[b * ch.n_joints + j] = t;
    }
  }
}

// cv2.undistortPoints(src, K, dist, P=K): pixel -> normalised -> fixed-point inversion of the
// Brown-Conrady model (OpenCV's default: 5 iterations) -> pixel of the ideal pinhole camera.
__global__ void __launch_bounds__(kFkThreads) undistort_kernel(const float* __restrict__ kp,
                                                               const mvgeo_camera* __restrict__ cams, int64_t n,
                                                               int V, int K, int iters, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kFkThreads + threadIdx.x;
  if (i >= n) return;
  const int v = (int)((i / K) % V);
  const mvgeo_camera& c = cams[v];
  const float k1 = c.dist[0], k2 = c.dist[1], p1 = c.dist[2], p2 = c.dist[3], k3 = c.dist[4];
  const float x0 = (kp[2 * i] - c.cx) / c.fx, y0 = (kp[2 * i + 1] - c.cy) / c.fy;
  float x = x0, y = y0;
  for (int it = 0; it < iters; ++it) {
    const float r2 = x * x + y * y;
    const float icdist = 1.0f / (1.0f + ((k3 * r2 + k2) * r2 + k1) * r2);
    const float dx = 2.0f * p1 * x * y + p2 * (r2 + 2.0f * x * x);
    const float dy = p1 * (r2 + 2.0f * y * y) + 2.0f * p2 * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  out[2 * i] = c.fx * x + c.cx;
  out[2 * i + 1] = c.fy * y + c.cy;
}

// Deterministic fixed-order sum of n floats into out[0]: one CTA, strided per-thread partials,
// then a shared-memory tree.
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float sh[1024];
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) a += x[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

int launch_sum(const float* x, int64_t n, float* out, cudaStream_t st) {
  sum_kernel<<<1, 1024, 0, st>>>(x, n, out);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

static int check_chain(const mvgeo_chain* c) {
  if (!c) return MVGEO_ENULL;
  if (c->n_joints < 1 || c->n_joints > MVGEO_MAX_JOINTS) return MVGEO_EINVAL;
  if (c->convention != MVGEO_DH_STANDARD && c->convention != MVGEO_DH_MODIFIED) return MVGEO_EINVAL;
  return MVGEO_OK;
}

}  // namespace mvgeo

using namespace mvgeo;

extern "C" int mvgeo_chain_builtin(int robot, mvgeo_chain* out) {
  if (!out) return MVGEO_ENULL;
  // (a, d, alpha_deg, theta_offset) per joint, numbers verbatim from the reference tables
  struct Row {
    double a, d, alpha, off;
  };
  static const Row fr3[7] = {{0, 0.333, 0, 0},       {0, 0, -90, 0}, {0, 0.316, 90, 0}, {0.0825, 0, 90, 0},
                             {-0.0825, 0.384, -90, 0}, {0, 0, 90, 0},  {0.088, 0, 90, 0}};
  static const Row fr5[6] = {{0, 0.152, 90, 0},  {-0.425, 0, 0, 0},  {-0.395, 0, 0, 0},
                             {0, 0.102, 90, 0},  {0, 0.102, -90, 0}, {0, 0.100, 0, 0}};
  static const Row meca[6] = {{0, 0.135, -90, 0}, {0.135, 0, 0, -90}, {0.038, 0, -90, 0},
                              {0, 0.120, 90, 0},  {0, 0, -90, 0},     {0, 0.070, 0, 0}};
  const Row* rows;
  int n;
  const double kPi = 3.14159265358979323846;
  mvgeo_chain c = {};
  switch (robot) {
    case MVGEO_ROBOT_FR3:
      rows = fr3; n = 7; c.convention = MVGEO_DH_MODIFIED; c.angle_scale = 1.0f;
      break;
    case MVGEO_ROBOT_FR5:
      rows = fr5; n = 6; c.convention = MVGEO_DH_STANDARD; c.angle_scale = (float)(kPi / 180.0);
      break;
    case MVGEO_ROBOT_MECA500:
      rows = meca; n = 6; c.convention = MVGEO_DH_STANDARD; c.angle_scale = (float)(kPi / 180.0);
      break;
    default:
      return MVGEO_EINVAL;
  }
  c.n_joints = n;
  c.emit_base = 1;
  for (int i = 0; i < n; ++i) {
    const double al = rows[i].alpha * (kPi / 180.0);  // math.radians
    c.a[i] = (float)rows[i].a;
    c.d[i] = (float)rows[i].d;
    c.cos_alpha[i] = (float)cos(al);
    c.sin_alpha[i] = (float)sin(al);
    c.theta_offset[i] = (float)rows[i].off;
  }
  *out = c;
  return MVGEO_OK;
}

extern "C" int mvgeo_fk(const mvgeo_chain* chain, const float* q, int64_t B, const float* R_view, int V, float* X,
                        void* stream) {
  int rc = check_chain(chain);
  if (rc) return rc;
  if (B < 0 || V < 1 || V > MVGEO_MAX_VIEWS) return MVGEO_EINVAL;
  if (!R_view && V != 1) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!q || !X) return MVGEO_ENULL;
  const unsigned grid = (unsigned)((B + kFkThreads - 1) / kFkThreads);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (chain->emit_base) fk_kernel<true><<<grid, kFkThreads, 0, st>>>(*chain, q, B, R_view, V, X);
  else fk_kernel<false><<<grid, kFkThreads, 0, st>>>(*chain, q, B, R_view, V, X);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

extern "C" int mvgeo_project(const float* X, int x_per_view, const mvgeo_camera* cams, int64_t B, int V, int K,
                             float* uv, void* stream) {
  if (B < 0 || V < 1 || V > MVGEO_MAX_VIEWS || K < 1) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!X || !cams || !uv) return MVGEO_ENULL;
  const int64_t n = B * V * K;
  const unsigned grid = (unsigned)((n + kFkThreads - 1) / kFkThreads);
  project_kernel<<<grid, kFkThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, x_per_view, cams, B, V, K, uv);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

extern "C" int mvgeo_undistort_points(const float* kp, const mvgeo_camera* cams, int64_t B, int V, int K, int iters,
                                      float* out, void* stream) {
  if (B < 0 || V < 1 || V > MVGEO_MAX_VIEWS || K < 1 || iters < 0 || iters > 100) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!kp || !cams || !out) return MVGEO_ENULL;
  const int64_t n = B * V * K;
  const unsigned grid = (unsigned)((n + kFkThreads - 1) / kFkThreads);
  undistort_kernel<<<grid, kFkThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(kp, cams, n, V, K, iters, out);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}

extern "C" int mvgeo_fk_reproj_fwd(const mvgeo_chain* chain, const float* q, int64_t B, const float* R_view,
                                   const mvgeo_camera* cams, int V, const float* gt_uv, const float* w, float lambda,
                                   float* X_out, float* uv_out, float* frame_loss, float* loss, void* stream) {
  int rc = check_chain(chain);
  if (rc) return rc;
  if (B < 0 || V < 1 || V > MVGEO_MAX_VIEWS) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!q || !cams) return MVGEO_ENULL;
  if (loss && !frame_loss) return MVGEO_ENULL;
  if (frame_loss && !gt_uv) return MVGEO_ENULL;
  const int K = chain->n_joints + (chain->emit_base ? 1 : 0);
  const float scale = (float)((double)lambda / ((double)B * V * K * 2.0));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int fpc = kFkThreads / V;  // V <= 16, so at least 8 frames per CTA
  const unsigned grid = (unsigned)((B + fpc - 1) / fpc);
  if (chain->emit_base)
    fk_reproj_fwd_kernel<true><<<grid, fpc * V, 0, st>>>(*chain, q, B, R_view, cams, V, gt_uv, w, scale, X_out,
                                                            uv_out, frame_loss);
  else
    fk_reproj_fwd_kernel<false><<<grid, fpc * V, 0, st>>>(*chain, q, B, R_view, cams, V, gt_uv, w, scale, X_out,
                                                             uv_out, frame_loss);
  MVGEO_CHECK_LAUNCH();
  if (loss) return launch_sum(frame_loss, B, loss, st);
  return MVGEO_OK;
}

extern "C" int mvgeo_fk_reproj_bwd(const mvgeo_chain* chain, const float* q, int64_t B, const float* R_view,
                                   const mvgeo_camera* cams, int V, const float* gt_uv, const float* w, float lambda,
                                   const float* dloss, float* dq, void* stream) {
  int rc = check_chain(chain);
  if (rc) return rc;
  if (B < 0 || V < 1 || V > MVGEO_MAX_VIEWS) return MVGEO_EINVAL;
  if (B == 0) return MVGEO_OK;
  if (!q || !cams || !gt_uv || !dq) return MVGEO_ENULL;
  const int K = chain->n_joints + (chain->emit_base ? 1 : 0);
  const float scale = (float)((double)lambda / ((double)B * V * K * 2.0));
  const int fpc = kFkThreads / V;
  const unsigned grid = (unsigned)((B + fpc - 1) / fpc);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (chain->emit_base)
    fk_reproj_bwd_kernel<true><<<grid, fpc * V, 0, st>>>(*chain, q, B, R_view, cams, V, gt_uv, w, scale, dloss, dq);
  else
    fk_reproj_bwd_kernel<false><<<grid, fpc * V, 0, st>>>(*chain, q, B, R_view, cams, V, gt_uv, w, scale, dloss, dq);
  MVGEO_CHECK_LAUNCH();
  return MVGEO_OK;
}
