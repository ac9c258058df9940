"""
oracle/mvgeo_oracle.py — CPU restatement of the reference's geometry hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it. The product package
(2025_icra_multi_view_robot_pose_estimation_b200) never imports anything from oracle/
and has no CPU fallback.

Reference: Najongs/2025_ICRA_Multi_View_Robot_Pose_Estimation (pure Python; paths below
are relative to the reference checkout, notebook citations are raw .ipynb line numbers).

Pinning status
  * decode (arg-max), the three DH chains, the generic FK class, projection and the GT
    belief-map encoder exist in the reference. Their restatements here are pinned against
    golden vectors produced by the UNMODIFIED reference functions, imported in the build
    container by tests/golden/make_golden.py (fixtures: tests/golden/*.npz), and against
    the known-answer values listed in SURVEY.md section 8c.
  * Projection delegates in the reference to cv2.projectPoints / cv2.Rodrigues (OpenCV, not
    vendored, unpinned by the reference). The restatement follows OpenCV's published
    Brown-Conrady model and is pinned against cv2 4.13.0 wherever cv2 imports.
  * Multi-view triangulation, soft-arg-max and the differentiable FK/reprojection loss DO
    NOT EXIST in the reference ("parity unpinned" against the reference for these three):
    the float64 definitions here are the specification. They are cross-checked against
    cv2.triangulatePoints (V=2), against closed-loop geometry (FK -> project -> triangulate
    returns FK) and against torch.autograd.gradcheck.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------------------
# Kinematic tables (verbatim numbers from the reference)
# --------------------------------------------------------------------------------------
#: model/MvRoPose_FR3.py:94-101 — Craig modified DH, (a [m], d [m], alpha [deg]); 8th row
#: (flange) exists in the table but is never applied because the loop runs over the 7
#: input angles (:121).
FR3_DH = [
    (0.0, 0.333, 0.0),
    (0.0, 0.0, -90.0),
    (0.0, 0.316, 90.0),
    (0.0825, 0.0, 90.0),
    (-0.0825, 0.384, -90.0),
    (0.0, 0.0, 90.0),
    (0.088, 0.0, 90.0),
    (0.0, 0.107, 0.0),
]
#: model/Fr5_model_train.ipynb:258-265 — standard DH, (alpha [deg], a [m], d [m], theta0 [deg])
FR5_DH = [
    (90.0, 0.0, 0.152, 0.0),
    (0.0, -0.425, 0.0, 0.0),
    (0.0, -0.395, 0.0, 0.0),
    (90.0, 0.0, 0.102, 0.0),
    (-90.0, 0.0, 0.102, 0.0),
    (0.0, 0.0, 0.100, 0.0),
]
#: visualization/Meca500_vis.ipynb:65-70 — standard DH, (alpha [deg], a, d, theta_offset [deg])
MECA500_DH = [
    (-90.0, 0.0, 0.135, 0.0),
    (0.0, 0.135, 0.0, -90.0),
    (-90.0, 0.038, 0.0, 0.0),
    (90.0, 0.0, 0.120, 0.0),
    (-90.0, 0.0, 0.0, 0.0),
    (0.0, 0.0, 0.070, 0.0),
]
#: Per-view base rotations, scipy `R.from_euler('zyx', angles, degrees=True)`:
#: FR3 model/MvRoPose_FR3.py:105-110, Fr5 model/Fr5_model_train.ipynb:269-273.
VIEW_EULER_ZYX_DEG = {
    "fr3": {"view1": (90, 180, 0), "view2": (90, 180, 0), "view3": (90, 180, 0), "view4": (90, 180, 0)},
    "fr5": {"top": (-85, 0, 180), "left": (180, 0, 90), "right": (0, 0, 90)},
    "meca500": {},
}


def euler_zyx_extrinsic(angles_deg) -> np.ndarray:
    """scipy Rotation.from_euler('zyx', [a, b, c], degrees=True).as_matrix(): lower-case
    axes are extrinsic, applied in order z, y, x, i.e. R = Rx(c) @ Ry(b) @ Rz(a)."""
    a, b, c = (math.radians(float(v)) for v in angles_deg)
    ca, sa, cb, sb, cc, sc = math.cos(a), math.sin(a), math.cos(b), math.sin(b), math.cos(c), math.sin(c)
    Rz = np.array([[ca, -sa, 0.0], [sa, ca, 0.0], [0.0, 0.0, 1.0]])
    Ry = np.array([[cb, 0.0, sb], [0.0, 1.0, 0.0], [-sb, 0.0, cb]])
    Rx = np.array([[1.0, 0.0, 0.0], [0.0, cc, -sc], [0.0, sc, cc]])
    return Rx @ Ry @ Rz


def view_rotation(robot: str, view) -> np.ndarray:
    """Base correction of `angle_to_joint_coordinate`; unknown view -> identity
    (`if selected_view in view_rotations`, model/MvRoPose_FR3.py:113-114)."""
    table = VIEW_EULER_ZYX_DEG[robot]
    if view in table:
        return euler_zyx_extrinsic(table[view])
    return np.eye(3)


# --------------------------------------------------------------------------------------
# Forward kinematics (a4-a7)
# --------------------------------------------------------------------------------------
def modified_dh_matrix(a, d, alpha_deg, theta_deg) -> np.ndarray:
    """model/MvRoPose_FR3.py:75-88 (Craig): Rx(alpha) Tx(a) Rz(theta) Tz(d)."""
    al, th = math.radians(alpha_deg), math.radians(theta_deg)
    ct, st, ca, sa = np.cos(th), np.sin(th), np.cos(al), np.sin(al)
    return np.array(
        [[ct, -st, 0.0, a], [st * ca, ct * ca, -sa, -d * sa], [st * sa, ct * sa, ca, d * ca], [0.0, 0.0, 0.0, 1.0]]
    )


def standard_dh_matrix(a, d, alpha_deg, theta_deg) -> np.ndarray:
    """model/Fr5_model_train.ipynb:246-254, visualization/Meca500_vis.ipynb:51-60:
    Rz(theta) Tz(d) Tx(a) Rx(alpha)."""
    al, th = math.radians(alpha_deg), math.radians(theta_deg)
    ct, st, ca, sa = np.cos(th), np.sin(th), np.cos(al), np.sin(al)
    return np.array(
        [[ct, -st * ca, st * sa, a * ct], [st, ct * ca, -ct * sa, a * st], [0.0, sa, ca, d], [0.0, 0.0, 0.0, 1.0]]
    )


def fk_fr3(joint_angles, view=None, dtype=np.float32) -> np.ndarray:
    """angle_to_joint_coordinate, model/MvRoPose_FR3.py:90-131. Radians in, (8,3) out."""
    T = np.eye(4)
    T[:3, :3] = view_rotation("fr3", view)
    pts = [np.zeros(3)]
    for i, q in enumerate(joint_angles):
        a, d, alpha = FR3_DH[i]
        T = T @ modified_dh_matrix(a, d, alpha, math.degrees(float(q)))
        pts.append(T[:3, 3].copy())
    return np.array(pts, dtype=dtype)


def fk_fr5(joint_angles_deg, view=None, dtype=np.float32) -> np.ndarray:
    """angle_to_joint_coordinate, model/Fr5_model_train.ipynb:256-288. Degrees in, (7,3) out."""
    T = np.eye(4)
    T[:3, :3] = view_rotation("fr5", view)
    pts = [np.zeros(3)]
    for i in range(6):
        alpha, a, d, th0 = FR5_DH[i]
        T = T @ standard_dh_matrix(a, d, alpha, float(joint_angles_deg[i]) + th0)
        pts.append(T[:3, 3].copy())
    return np.array(pts, dtype=dtype)


def fk_meca500(joint_angles_deg, dtype=np.float32) -> np.ndarray:
    """forward_kinematics, visualization/Meca500_vis.ipynb:62-82. Degrees in, (7,3) out."""
    T = np.eye(4)
    pts = [np.zeros(3)]
    for i in range(6):
        alpha, a, d, th0 = MECA500_DH[i]
        T = T @ standard_dh_matrix(a, d, alpha, float(joint_angles_deg[i]) + th0)
        pts.append(T[:3, 3].copy())
    return np.array(pts, dtype=dtype)


def fk_generic(dh_params, angles) -> np.ndarray:
    """ForwardKinematics.forward, model/MV-model.ipynb:841-874: standard DH, tuples
    (theta0, d, a, alpha) in radians, float32 link matrices (:856), joints only (no base).
    angles (B,J) -> float32 (B,J,3)."""
    angles = np.asarray(angles)
    out = np.zeros(angles.shape + (3,), dtype=np.float32)
    for b in range(angles.shape[0]):
        T = np.eye(4)
        for j in range(angles.shape[1]):
            th0, d, a, al = dh_params[j]
            th = th0 + float(angles[b, j])
            ct, st, ca, sa = np.cos(th), np.sin(th), np.cos(al), np.sin(al)
            Tj = np.array(
                [[ct, -st * ca, st * sa, a * ct], [st, ct * ca, -ct * sa, a * st], [0, sa, ca, d], [0, 0, 0, 1]],
                dtype=np.float32,
            )
            T = T @ Tj
            out[b, j] = T[:3, 3]
    return out


def chain_spec(robot: str) -> dict:
    """Uniform description used by the batched restatements below and mirrored by the
    product's mvgeo_chain struct (include/mvgeo.h)."""
    if robot == "fr3":
        rows = FR3_DH[:7]
        return dict(convention="modified", emit_base=True, angle_scale=1.0,
                    a=[r[0] for r in rows], d=[r[1] for r in rows], alpha_deg=[r[2] for r in rows],
                    theta_offset=[0.0] * 7)
    if robot == "fr5":
        return dict(convention="standard", emit_base=True, angle_scale=math.pi / 180.0,
                    a=[r[1] for r in FR5_DH], d=[r[2] for r in FR5_DH], alpha_deg=[r[0] for r in FR5_DH],
                    theta_offset=[r[3] for r in FR5_DH])
    if robot == "meca500":
        return dict(convention="standard", emit_base=True, angle_scale=math.pi / 180.0,
                    a=[r[1] for r in MECA500_DH], d=[r[2] for r in MECA500_DH],
                    alpha_deg=[r[0] for r in MECA500_DH], theta_offset=[r[3] for r in MECA500_DH])
    raise KeyError(robot)


def fk_chain(spec: dict, q, R_view=None) -> np.ndarray:
    """Vectorised float64 FK for any chain spec. q (B,J) in the chain's native unit,
    R_view (V,3,3) or None -> (B,V,K,3) float64. Same arithmetic as the per-robot
    functions above (left-multiplied base rotation, cumulative product, frame origins)."""
    q = np.asarray(q, dtype=np.float64)
    B, J = q.shape
    Rv = np.eye(3)[None] if R_view is None else np.asarray(R_view, dtype=np.float64)
    V = Rv.shape[0]
    T = np.zeros((B, V, 4, 4))
    T[..., :3, :3] = Rv[None]
    T[..., 3, 3] = 1.0
    pts = [np.zeros((B, V, 3))] if spec["emit_base"] else []
    for i in range(J):
        th = (q[:, i] + spec["theta_offset"][i]) * spec["angle_scale"]
        al = math.radians(spec["alpha_deg"][i]) if "alpha_deg" in spec else spec["alpha_rad"][i]
        ct, st, ca, sa = np.cos(th), np.sin(th), np.cos(al), np.sin(al)
        a, d = spec["a"][i], spec["d"][i]
        Ti = np.zeros((B, 4, 4))
        Ti[:, 3, 3] = 1.0
        if spec["convention"] == "modified":
            Ti[:, 0, 0], Ti[:, 0, 1], Ti[:, 0, 3] = ct, -st, a
            Ti[:, 1, 0], Ti[:, 1, 1], Ti[:, 1, 2], Ti[:, 1, 3] = st * ca, ct * ca, -sa, -d * sa
            Ti[:, 2, 0], Ti[:, 2, 1], Ti[:, 2, 2], Ti[:, 2, 3] = st * sa, ct * sa, ca, d * ca
        else:
            Ti[:, 0, 0], Ti[:, 0, 1], Ti[:, 0, 2], Ti[:, 0, 3] = ct, -st * ca, st * sa, a * ct
            Ti[:, 1, 0], Ti[:, 1, 1], Ti[:, 1, 2], Ti[:, 1, 3] = st, ct * ca, -ct * sa, a * st
            Ti[:, 2, 1], Ti[:, 2, 2], Ti[:, 2, 3] = sa, ca, d
        T = T @ Ti[:, None]
        pts.append(T[..., :3, 3].copy())
    return np.stack(pts, axis=2)


# --------------------------------------------------------------------------------------
# Projection (a8): cv2.projectPoints restated
# --------------------------------------------------------------------------------------
def rodrigues(rvec) -> np.ndarray:
    """cv2.Rodrigues(rvec)[0]: R = cos(t) I + (1-cos t) k k^T + sin(t) [k]x."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = float(np.linalg.norm(r))
    if th < 2.220446049250313e-16:  # OpenCV: theta < DBL_EPSILON -> identity
        return np.eye(3)
    k = r / th
    c, s = math.cos(th), math.sin(th)
    Kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return c * np.eye(3) + (1.0 - c) * np.outer(k, k) + s * Kx


def project_points(X, R, t, K, dist=None) -> np.ndarray:
    """cv2.projectPoints(X, rvec, tvec, K, dist) with R = Rodrigues(rvec): pinhole +
    Brown-Conrady [k1,k2,p1,p2,k3]. X (...,3) -> (...,2), float64. Call sites:
    model/MvRoPose_FR3.py:133-141, model/Fr5_model_train.ipynb:290-305,
    visualization/Fr5_vis.ipynb:111-115, model/MV-model.ipynb:879-899."""
    X = np.asarray(X, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64).reshape(3)
    K = np.asarray(K, dtype=np.float64)
    Xc = X @ R.T + t
    z = Xc[..., 2]
    xp, yp = Xc[..., 0] / z, Xc[..., 1] / z
    if dist is not None:
        k1, k2, p1, p2, k3 = (float(v) for v in np.asarray(dist, dtype=np.float64).reshape(-1)[:5])
        r2 = xp * xp + yp * yp
        rad = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3))
        xpp = xp * rad + 2.0 * p1 * xp * yp + p2 * (r2 + 2.0 * xp * xp)
        ypp = yp * rad + p1 * (r2 + 2.0 * yp * yp) + 2.0 * p2 * xp * yp
        xp, yp = xpp, ypp
    u = K[0, 0] * xp + K[0, 2]
    v = K[1, 1] * yp + K[1, 2]
    return np.stack([u, v], axis=-1)


def projection_matrix(K, R, t) -> np.ndarray:
    """P = K [R | t], float64 (3,4)."""
    return np.asarray(K, dtype=np.float64) @ np.hstack([np.asarray(R, dtype=np.float64), np.asarray(t, dtype=np.float64).reshape(3, 1)])


# --------------------------------------------------------------------------------------
# Belief-map decode (a1-a3)
# --------------------------------------------------------------------------------------
def argmax_first(maps):
    """Flat arg-max per map with torch.argmax / np.argmax semantics: the FIRST maximal
    element wins and NaN counts as maximal. maps (...,H,W) -> (idx int64 (...), peak (...))."""
    maps = np.asarray(maps)
    flat = maps.reshape(maps.shape[:-2] + (-1,))
    idx = np.argmax(flat, axis=-1)
    peak = np.take_along_axis(flat, idx[..., None], axis=-1)[..., 0]
    return idx.astype(np.int64), peak


def sigmoid64(x):
    x = np.asarray(x, dtype=np.float64)
    return 1.0 / (1.0 + np.exp(-x))


def soft_argmax(maps, beta: float, mode: str = "global", radius: int = 3):
    """Sub-pixel soft-arg-max (absent from the reference; specification):
        weights w_i = exp(beta * (h_i - max h)) over the whole map ('global') or over the
        (2r+1)^2 window clipped to the map around the hard arg-max ('window');
        (x, y) = sum w_i (x_i, y_i) / sum w_i in map pixels (integer pixel centres, the same
        convention as the hard key-point x = idx % W, y = idx // W).
    A NaN peak gives NaN; a -inf peak (all-(-inf) map) gives the hard peak.
    maps (...,H,W) -> float64 (...,2)."""
    maps = np.asarray(maps, dtype=np.float64)
    H, W = maps.shape[-2:]
    lead = maps.shape[:-2]
    flat = maps.reshape((-1, H, W))
    out = np.zeros((flat.shape[0], 2))
    for m in range(flat.shape[0]):
        h = flat[m]
        i = int(np.argmax(h.reshape(-1)))
        py, px = divmod(i, W)
        peak = h[py, px]
        if np.isnan(peak):
            out[m] = np.nan
            continue
        if np.isneginf(peak):
            out[m] = (px, py)
            continue
        if mode == "window":
            y0, y1 = max(0, py - radius), min(H, py + radius + 1)
            x0, x1 = max(0, px - radius), min(W, px + radius + 1)
        else:
            y0, y1, x0, x1 = 0, H, 0, W
        sub = h[y0:y1, x0:x1]
        w = np.exp(beta * (sub - peak))
        s = w.sum()
        ys, xs = np.arange(y0, y1)[:, None], np.arange(x0, x1)[None, :]
        out[m] = ((w * xs).sum() / s, (w * ys).sum() / s)
    return out.reshape(lead + (2,))


def decode(maps, scale_x: float = 1.0, scale_y: float = 1.0, soft_mode: str = "none", beta: float = 1.0,
           radius: int = 3, apply_sigmoid: bool = False) -> dict:
    """Batched restatement of the decoder for maps (...,H,W):
      idx/peak: arg-max of the RAW map (inline call sites DIP_REAL.py:116-124,
      model/MvRoPose_FR3.py:299-304); kp_hard = float32(x*scale_x), float32(y*scale_y) with
      the product in Python float64 (`x * (original_w / w)` stored into a float32 array,
      model/Fr5_model_train.ipynb:4701-4703); score = sigmoid(peak) or peak (:4685,:4695)."""
    maps = np.asarray(maps)
    H, W = maps.shape[-2:]
    idx, peak = argmax_first(maps)
    x, y = idx % W, idx // W
    kp_hard = np.stack([(x * float(scale_x)), (y * float(scale_y))], axis=-1).astype(np.float32)
    peak64 = peak.astype(np.float64)
    score = sigmoid64(peak64) if apply_sigmoid else peak64
    out = dict(idx=idx, peak=peak.astype(np.float32), score=score, kp_hard=kp_hard)
    if soft_mode != "none":
        s = soft_argmax(maps, beta, soft_mode, radius)
        out["kp_soft"] = s * np.array([float(scale_x), float(scale_y)])
    return out


def extract_keypoints_from_heatmaps(heatmaps, original_image_size):
    """Call-for-call port of model/Fr5_model_train.ipynb:4674-4705 (torch CPU tensor in,
    per-key-point loop, sigmoid first). Used as the CPU baseline's decode stage and to
    state the sigmoid-then-argmax semantics exactly."""
    import torch  # local: the vectorised oracle above has no torch dependency

    num_joints, h, w = heatmaps.shape
    original_h, original_w = original_image_size
    keypoints = np.zeros((num_joints, 2), dtype=np.float32)
    scores = np.zeros(num_joints, dtype=np.float32)
    heatmaps = heatmaps.sigmoid()
    for i in range(num_joints):
        max_val, max_idx = torch.max(heatmaps[i].reshape(-1), dim=0)
        scores[i] = max_val.item()
        y, x = np.unravel_index(max_idx.cpu().numpy(), (h, w))
        keypoints[i] = [x * (original_w / w), y * (original_h / h)]
    return keypoints, scores


def decode_inline_argmax(heatmaps, frame_hw):
    """Port of the inline loop DIP_REAL.py:116-124 (torch CPU (K,h,w) in, (K,2) float64 out)."""
    import torch

    h, w = heatmaps.shape[1:]
    fh, fw = frame_hw
    kps = []
    for j in range(heatmaps.shape[0]):
        y, x = np.unravel_index(torch.argmax(heatmaps[j]).numpy(), (h, w))
        kps.append([x * (fw / w), y * (fh / h)])
    return np.array(kps)


# --------------------------------------------------------------------------------------
# Multi-view DLT triangulation (a10; specification, no reference implementation)
# --------------------------------------------------------------------------------------
def triangulate_dlt(kp, P, w=None, min_weight: float = 0.0, weighted: bool = False):
    """Homogeneous DLT in float64. kp (B,V,K,2) pixels, P (V,3,4), w (B,V,K) or None.
    Per key-point: rows u*P[2]-P[0], v*P[2]-P[1] for each valid view (weight >= min_weight
    and finite key-point; rows scaled by the weight when `weighted`), X = last right-singular
    vector de-homogenised. Fewer than two valid views -> NaN.
    Returns X (B,K,3), rms reprojection residual in px (B,K), n_views (B,K)."""
    kp = np.asarray(kp, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    B, V, K, _ = kp.shape
    ww = np.ones((B, V, K)) if w is None else np.asarray(w, dtype=np.float64)
    X = np.full((B, K, 3), np.nan)
    resid = np.full((B, K), np.nan)
    nv = np.zeros((B, K), dtype=np.int32)
    for b in range(B):
        for k in range(K):
            rows, used = [], []
            for v in range(V):
                u_, v_ = kp[b, v, k]
                wt = ww[b, v, k]
                if not (wt >= min_weight) or not np.isfinite(u_) or not np.isfinite(v_):
                    continue
                s = wt if weighted else 1.0
                rows.append(s * (u_ * P[v, 2] - P[v, 0]))
                rows.append(s * (v_ * P[v, 2] - P[v, 1]))
                used.append(v)
            nv[b, k] = len(used)
            if len(used) < 2:
                continue
            A = np.array(rows)
            x = np.linalg.svd(A)[2][-1]
            if x[3] == 0.0:
                continue
            Xp = x[:3] / x[3]
            X[b, k] = Xp
            e2 = 0.0
            for v in used:
                ph = P[v] @ np.append(Xp, 1.0)
                e2 += np.sum((ph[:2] / ph[2] - kp[b, v, k]) ** 2)
            resid[b, k] = math.sqrt(e2 / len(used))
    return X, resid, nv


# --------------------------------------------------------------------------------------
# FK + reprojection loss with autograd (a9 + backward; specification for the backward)
# --------------------------------------------------------------------------------------
def fk_reproj_loss_torch(spec: dict, q, R_view, cams: list, gt_uv, w=None, lam: float = 1.0):
    """float64 torch restatement of FK -> project -> lambda * mse(mean over B*V*K*2)
    (loss form robot_pose_loss, model/MV-model.ipynb:942-950; projection
    model/MvRoPose_FR3.py:133-141). q: torch (B,J) float64 (may require grad).
    cams: list of dict(R (3,3), t (3,), K (3,3), dist (5,) or None) per view.
    Returns (loss, X (B,V,K,3), uv (B,V,K,2)). Points with non-finite gt are skipped but
    the divisor stays B*V*K*2."""
    import torch

    B, J = q.shape
    V = len(cams)
    Rv = torch.eye(3, dtype=torch.float64).repeat(V, 1, 1) if R_view is None else torch.as_tensor(np.asarray(R_view), dtype=torch.float64)
    T = torch.zeros(B, V, 4, 4, dtype=torch.float64)
    T[..., :3, :3] = Rv
    T[..., 3, 3] = 1.0
    pts = [torch.zeros(B, V, 3, dtype=torch.float64)] if spec["emit_base"] else []
    for i in range(J):
        th = (q[:, i] + spec["theta_offset"][i]) * spec["angle_scale"]
        al = math.radians(spec["alpha_deg"][i]) if "alpha_deg" in spec else spec["alpha_rad"][i]
        ct, st = torch.cos(th), torch.sin(th)
        ca, sa = math.cos(al), math.sin(al)
        a, d = spec["a"][i], spec["d"][i]
        z, o = torch.zeros_like(ct), torch.ones_like(ct)
        if spec["convention"] == "modified":
            rows = [[ct, -st, z, a * o], [st * ca, ct * ca, -sa * o, -d * sa * o], [st * sa, ct * sa, ca * o, d * ca * o], [z, z, z, o]]
        else:
            rows = [[ct, -st * ca, st * sa, a * ct], [st, ct * ca, -ct * sa, a * st], [z, sa * o, ca * o, d * o], [z, z, z, o]]
        Ti = torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)
        T = T @ Ti[:, None]
        pts.append(T[..., :3, 3])
    X = torch.stack(pts, dim=2)  # (B,V,K,3)
    uvs = []
    for v, cam in enumerate(cams):
        R = torch.as_tensor(np.asarray(cam["R"]), dtype=torch.float64)
        t = torch.as_tensor(np.asarray(cam["t"]).reshape(3), dtype=torch.float64)
        Km = np.asarray(cam["K"], dtype=np.float64)
        Xc = X[:, v] @ R.T + t
        xp, yp = Xc[..., 0] / Xc[..., 2], Xc[..., 1] / Xc[..., 2]
        dist = cam.get("dist")
        if dist is not None:
            k1, k2, p1, p2, k3 = (float(c) for c in np.asarray(dist).reshape(-1)[:5])
            r2 = xp * xp + yp * yp
            rad = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3))
            xpp = xp * rad + 2.0 * p1 * xp * yp + p2 * (r2 + 2.0 * xp * xp)
            ypp = yp * rad + p1 * (r2 + 2.0 * yp * yp) + 2.0 * p2 * xp * yp
            xp, yp = xpp, ypp
        uvs.append(torch.stack([Km[0, 0] * xp + Km[0, 2], Km[1, 1] * yp + Km[1, 2]], dim=-1))
    uv = torch.stack(uvs, dim=1)  # (B,V,K,2)
    gt = torch.as_tensor(np.asarray(gt_uv), dtype=torch.float64)
    ok = torch.isfinite(gt).all(dim=-1, keepdim=True)
    diff = torch.where(ok, uv - torch.nan_to_num(gt, nan=0.0, posinf=0.0, neginf=0.0), torch.zeros_like(uv))
    wt = torch.ones(B, V, X.shape[2], dtype=torch.float64) if w is None else torch.as_tensor(np.asarray(w), dtype=torch.float64)
    loss = lam * (wt[..., None] * diff * diff).sum() / (B * V * X.shape[2] * 2)
    return loss, X, uv


# --------------------------------------------------------------------------------------
# GT belief-map encoder and heat-map MSE ("next" rows 1-2)
# --------------------------------------------------------------------------------------
def create_gt_heatmap(keypoint_2d, heatmap_size, sigma) -> np.ndarray:
    """model/MvRoPose_FR3.py:65-73 (= model/DREAM_Train.py:60-69). float64 (H,W)."""
    H, W = heatmap_size
    x, y = keypoint_2d
    xx, yy = np.meshgrid(np.arange(W), np.arange(H))
    dist_sq = (xx - x) ** 2 + (yy - y) ** 2
    heatmap = np.exp(-dist_sq / (2 * sigma ** 2))
    heatmap[heatmap < np.finfo(float).eps * heatmap.max()] = 0
    return heatmap


def heatmap_mse(pred, kp, sigma: float, weight: float = 1.0):
    """nn.MSELoss()(pred, gt) * weight (model/MvRoPose_FR3.py:846-847) with gt rasterised by
    create_gt_heatmap from kp (n,2); non-finite centres give all-zero targets.
    Returns (loss float64, grad float64 same shape as pred)."""
    pred = np.asarray(pred, dtype=np.float64)
    n, H, W = pred.shape
    gt = np.zeros_like(pred)
    for m in range(n):
        if np.all(np.isfinite(kp[m])):
            gt[m] = create_gt_heatmap((float(kp[m][0]), float(kp[m][1])), (H, W), sigma)
    diff = pred - gt
    loss = weight * np.mean(diff * diff)
    grad = weight * 2.0 * diff / diff.size
    return loss, grad


# --------------------------------------------------------------------------------------
# Camera-pose refinement ("next" row 3; specification, cross-checked against cv2.solvePnP)
# --------------------------------------------------------------------------------------
def rvec_from_matrix(R) -> np.ndarray:
    """cv2.Rodrigues(R)[0]: rotation matrix -> rotation vector (float64)."""
    R = np.asarray(R, dtype=np.float64)
    c = min(1.0, max(-1.0, 0.5 * (np.trace(R) - 1.0)))
    th = math.acos(c)
    a = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = 0.5 * np.linalg.norm(a)
    if s < 1e-10:
        if c > 0:
            return 0.5 * a
        d = np.sqrt(np.maximum(0.0, 0.5 * (np.diag(R) + 1.0)))
        d[1] = -d[1] if R[0, 1] + R[1, 0] < 0 else d[1]
        d[2] = -d[2] if R[0, 2] + R[2, 0] < 0 else d[2]
        return th * d
    return (0.5 * th / s) * a


def pnp_refine(X, kp, K, dist, R0, t0, w=None, min_weight: float = 0.0, max_iters: int = 50):
    """Minimise sum_k |project(R X_k + t) - kp_k|^2 over the pose by Levenberg-Marquardt from the
    prior (R0, t0), in float64 — the same cost cv2.solvePnP(SOLVEPNP_ITERATIVE,
    useExtrinsicGuess=True) minimises (call site being replaced: estimate_camera_pose,
    model/Fr5_model_train.ipynb:4707-4753). Points with w < min_weight or non-finite kp are
    dropped; fewer than 4 valid points (the reference's refusal, :4728) returns the prior.
    X (K,3), kp (K,2). Returns (rvec, tvec, rms_px, status) with status bits as the kernel."""
    X = np.asarray(X, dtype=np.float64)
    kp = np.asarray(kp, dtype=np.float64)
    R, t = np.asarray(R0, dtype=np.float64).copy(), np.asarray(t0, dtype=np.float64).reshape(3).copy()
    ok = np.isfinite(kp).all(axis=1)
    if w is not None:
        ok &= np.asarray(w, dtype=np.float64) >= min_weight
    n = int(ok.sum())
    if n < 4:
        return rvec_from_matrix(R), t, float("nan"), 0
    Xv, kv = X[ok], kp[ok]

    def residual(Rr, tt):
        return (project_points(Xv, Rr, tt, K, dist) - kv).reshape(-1)

    def jacobian(Rr, tt, eps=1e-7):
        J = np.zeros((2 * n, 6))
        for j in range(6):
            d = np.zeros(6)
            d[j] = eps
            Rp, Rm = rodrigues(d[:3]) @ Rr, rodrigues(-d[:3]) @ Rr
            J[:, j] = (residual(Rp, tt + d[3:]) - residual(Rm, tt - d[3:])) / (2 * eps)
        return J

    lam, status = 1e-3, 1
    r = residual(R, t)
    cost = float(r @ r)
    for _ in range(max_iters):
        J = jacobian(R, t)
        H, g = J.T @ J, J.T @ r
        accepted = False
        for _try in range(8):
            try:
                d = np.linalg.solve(H + lam * np.diag(np.diag(H)) + 1e-12 * np.eye(6), -g)
            except np.linalg.LinAlgError:
                lam *= 10
                continue
            Rn, tn = rodrigues(d[:3]) @ R, t + d[3:]
            rn = residual(Rn, tn)
            cn = float(rn @ rn)
            if cn <= cost:
                small = (cost - cn) <= 1e-14 * cost + 1e-20
                R, t, r, cost, lam, accepted = Rn, tn, rn, cn, max(lam * 0.1, 1e-12), True
                if small:
                    status |= 2
                break
            lam *= 10
        if not accepted or (status & 2):
            status |= 2
            break
    if 0.25 < float(t @ t) < 25.0:
        status |= 4
    return rvec_from_matrix(R), t, math.sqrt(cost / n), status


def p3p_poses(f, P):
    """Perspective-3-point (Grunert's formulation, reviewed in Haralick et al. 1994): unit bearings f (3,3)
    and world points P (3,3) -> list of (R, t) with s_i f_i = R P_i + t, all real solutions with positive
    depths. With s2 = u s1, s3 = v s1 the three law-of-cosines equations reduce to a quartic in v whose
    coefficients are formed here by polynomial arithmetic (no hand-expanded coefficient table)."""
    f = np.asarray(f, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    a2, b2, c2 = np.sum((P[1] - P[2]) ** 2), np.sum((P[0] - P[2]) ** 2), np.sum((P[0] - P[1]) ** 2)
    if min(a2, b2, c2) < 1e-18:
        return []
    ca, cb, cg = float(f[1] @ f[2]), float(f[0] @ f[2]), float(f[0] @ f[1])
    k = (a2 - c2) / b2
    # u = N(v) / D(v) from (eq.1 - eq.3 scaled by eq.2)
    N = np.array([k - 1.0, -2.0 * k * cb, 1.0 + k])          # v^2, v^1, v^0
    D = np.array([-2.0 * ca, 2.0 * cg])                       # v^1, v^0
    q = np.array([1.0, -2.0 * cb, 1.0])                       # 1 + v^2 - 2 v cos(beta)
    DD = np.polymul(D, D)
    quartic = np.polyadd(np.polyadd(np.polyadd(DD, np.polymul(N, N)), -2.0 * cg * np.polymul(N, D)),
                         -(c2 / b2) * np.polymul(q, DD))
    out = []
    for v in np.roots(quartic):
        if abs(v.imag) > 1e-7 * max(1.0, abs(v.real)) or v.real <= 0:
            continue
        v = float(v.real)
        den = 2.0 * (cg - v * ca)
        if abs(den) < 1e-12:
            continue
        u = ((k - 1.0) * v * v - 2.0 * k * cb * v + 1.0 + k) / den
        w = 1.0 + v * v - 2.0 * v * cb
        if u <= 0 or w <= 0:
            continue
        s1 = math.sqrt(b2 / w)
        Q = np.stack([s1 * f[0], u * s1 * f[1], v * s1 * f[2]])

        def frame(A):
            e1 = A[1] - A[0]
            e1 = e1 / np.linalg.norm(e1)
            e3 = np.cross(e1, A[2] - A[0])
            n3 = np.linalg.norm(e3)
            if n3 < 1e-12:
                return None
            e3 = e3 / n3
            return np.stack([e1, np.cross(e3, e1), e3], axis=1)   # columns
        Fp, Fq = frame(P), frame(Q)
        if Fp is None or Fq is None:
            continue
        R = Fq @ Fp.T
        out.append((R, Q[0] - R @ P[0]))
    return out


def pnp_solve(X, kp, K, dist=None, w=None, min_weight: float = 0.0, reproj_thresh: float = 8.0, max_iters: int = 30):
    """Camera pose WITHOUT a prior — the specification of mvgeo_pnp_solve, which replaces
    cv2.solvePnPRansac(obj, img, K, dist, flags=SOLVEPNP_EPNP) in estimate_camera_pose
    (model/Fr5_model_train.ipynb:4707-4753; defaults: reprojectionError 8 px). K <= 9 key-points make random
    sampling pointless: EVERY point triplet is a hypothesis (P3P, up to 4 poses each), every hypothesis is
    scored on all valid points (inlier count at `reproj_thresh`, then squared error of the inliers), and the
    winner is refined by Levenberg-Marquardt on its inliers (pnp_refine above: the cost
    cv2.solvePnP(SOLVEPNP_ITERATIVE) minimises). Fewer than 4 valid points, no hypothesis, or fewer than 4
    inliers -> None (the reference's refusals, :4728 and `num_inliers >= 4`).
    Returns (rvec, tvec, inlier_mask (K,), rms_px, status) or None."""
    import itertools

    X = np.asarray(X, dtype=np.float64)
    kp = np.asarray(kp, dtype=np.float64)
    Kn = len(X)
    dist = np.zeros(5) if dist is None else np.asarray(dist, dtype=np.float64).reshape(-1)[:5]
    ok = np.isfinite(kp).all(axis=1) & np.isfinite(X).all(axis=1)
    if w is not None:
        ok &= np.asarray(w, dtype=np.float64) >= min_weight
    idx = np.flatnonzero(ok)
    if len(idx) < 4:
        return None
    und = undistort_points(kp, K, dist)                      # ideal pinhole pixels
    K = np.asarray(K, dtype=np.float64)
    xn = np.stack([(und[:, 0] - K[0, 2]) / K[0, 0], (und[:, 1] - K[1, 2]) / K[1, 1], np.ones(Kn)], axis=1)
    f = xn / np.linalg.norm(xn, axis=1, keepdims=True)
    best, best_key = None, None
    for tri in itertools.combinations(idx, 3):
        for R, t in p3p_poses(f[list(tri)], X[list(tri)]):
            Xc = X[idx] @ R.T + t
            if np.any(Xc[:, 2] <= 1e-9):
                continue                                      # a valid point behind the camera
            e2 = np.sum((project_points(X[idx], R, t, K, dist) - kp[idx]) ** 2, axis=1)
            inl = e2 < reproj_thresh ** 2
            key = (int(inl.sum()), -float(e2[inl].sum()))
            if best_key is None or key > best_key:
                best_key, best = key, (R, t, inl)
    if best is None or best_key[0] < 4:
        return None
    R, t, inl = best
    mask = np.zeros(Kn, dtype=bool)
    mask[idx[inl]] = True
    rvec, tvec, rms, status = pnp_refine(X, kp, K, dist, R, t, mask.astype(np.float64), 0.5, max_iters)
    return rvec, tvec, mask, rms, status


def undistort_points(kp, K, dist, iters: int = 5) -> np.ndarray:
    """cv2.undistortPoints(kp, K, dist, P=K): OpenCV's fixed-point inversion of the Brown-Conrady
    model (5 iterations by default), float64. kp (...,2) pixels -> (...,2) pixels."""
    kp = np.asarray(kp, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    k1, k2, p1, p2, k3 = (float(c) for c in np.asarray(dist, dtype=np.float64).reshape(-1)[:5])
    x0, y0 = (kp[..., 0] - K[0, 2]) / K[0, 0], (kp[..., 1] - K[1, 2]) / K[1, 1]
    x, y = x0.copy(), y0.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x, y = (x0 - dx) * icdist, (y0 - dy) * icdist
    return np.stack([K[0, 0] * x + K[0, 2], K[1, 1] * y + K[1, 2]], axis=-1)


def average_quaternion(quaternions, weights=None) -> np.ndarray:
    """dataset/Fr5_preprocessing.py:57-65 (average_quaternion): M = sum q q^T, eigenvector of the
    largest eigenvalue (np.linalg.eigh), normalised. Sign is arbitrary in the reference."""
    M = np.zeros((4, 4))
    for i, q in enumerate(np.asarray(quaternions, dtype=np.float64)):
        q = q.reshape(4, 1)
        M += (1.0 if weights is None else float(weights[i])) * (q @ q.T)
    eigvals, eigvecs = np.linalg.eigh(M)
    avg = eigvecs[:, np.argmax(eigvals)]
    return avg / np.linalg.norm(avg)
