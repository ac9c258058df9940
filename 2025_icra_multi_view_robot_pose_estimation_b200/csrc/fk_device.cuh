// fk_device.cuh — device code of the DH chain, projection and the reprojection-consistency
// forward pass, shared by fk.cu (stand-alone kernels) and geom.cu (fused geometry kernel).
#pragma once

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

// Forward pass of the FK-consistency loss for CTA `blk` of the (frame, view) decomposition:
// kFkThreads / V frames per CTA, one thread per (frame, view) (the chain is recomputed per view — J
// sincos, cheap — so a small batch still fills the machine); the per-frame sums over views go
// through `part` (shared, kFkThreads floats) in fixed view order (deterministic, no atomics).
template <bool BASE>
__device__ __forceinline__ void fk_reproj_fwd_body(const mvgeo_chain& ch, const float* __restrict__ q, int64_t B,
                                                   const float* __restrict__ R_view,
                                                   const mvgeo_camera* __restrict__ cams, int V,
                                                   const float* __restrict__ gt_uv, const float* __restrict__ w,
                                                   float scale, float* __restrict__ X_out,
                                                   float* __restrict__ uv_out, float* __restrict__ frame_loss,
                                                   float* part, int64_t blk) {
  const int fpc = kFkThreads / V;  // frames per CTA; threads beyond fpc * V idle
  const int fl = threadIdx.x / V, v = threadIdx.x - fl * V;
  const int64_t b = (int64_t)blk * fpc + fl;
  const bool active = (int)threadIdx.x < fpc * V && b < B;
  float acc = 0.f;
  if (active) {
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
    if (active && v == 0) {
      float t = 0.f;
      for (int i = 0; i < V; ++i) t += part[fl * V + i];
      frame_loss[b] = t * scale;
    }
  }
}

}  // namespace mvgeo
