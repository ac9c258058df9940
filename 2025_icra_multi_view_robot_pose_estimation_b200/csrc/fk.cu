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
#include "fk_device.cuh"

namespace mvgeo {

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

template <bool BASE>
__global__ void __launch_bounds__(kFkThreads)
    fk_reproj_fwd_kernel(const mvgeo_chain ch, const float* __restrict__ q, int64_t B,
                         const float* __restrict__ R_view, const mvgeo_camera* __restrict__ cams, int V,
                         const float* __restrict__ gt_uv, const float* __restrict__ w, float scale,
                         float* __restrict__ X_out, float* __restrict__ uv_out, float* __restrict__ frame_loss) {
  __shared__ float part[kFkThreads];
  fk_reproj_fwd_body<BASE>(ch, q, B, R_view, cams, V, gt_uv, w, scale, X_out, uv_out, frame_loss, part, blockIdx.x);
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
