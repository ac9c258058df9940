"""Drop-in shims with the reference's own names, argument meaning and return types.

The reference has no package: each script / notebook re-declares these free functions
(SURVEY.md section 8b). A caller switches by replacing its local definitions with
    from mvgeo.compat import fr3      # or fr5 / meca500 / generic
    angle_to_joint_coordinate = fr3.angle_to_joint_coordinate
Single-frame NumPy / CPU-tensor inputs are staged to the GPU, run through the same sm_100a
kernels as the batched API (ops.py) and copied back, so results are the kernels' results;
there is no CPU implementation behind these names. Each call therefore costs one small H2D
and one D2H (it synchronises, as the reference's NumPy-returning functions inherently do) —
use ops.* for throughput. Do not call from fork()ed DataLoader workers (CUDA cannot be
initialised after fork); use the spawn start method or the batched on-device path.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from . import ops
from .rig import CameraRig
from .robots import Chain, view_rotation


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("mvgeo.compat needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------- decoding
def extract_keypoints_from_heatmaps(heatmaps, original_image_size):
    """model/Fr5_model_train.ipynb:4674-4705 (= Franka_research3_model_train.ipynb:3634-3665,
    DREAM_model_train.ipynb:1517-1548). heatmaps: torch (K,h,w), CPU or CUDA.
    Returns (keypoints np.float32[K,2] in image pixels, scores np.float32[K] = sigmoid(max)).
    The arg-max is taken on the RAW map; the reference takes it after a float32 sigmoid, which
    can merge neighbouring floats into a tie — in that case the two
    indices differ but carry the same sigmoid score."""
    t = heatmaps if isinstance(heatmaps, torch.Tensor) else torch.as_tensor(np.asarray(heatmaps))
    r = ops.decode_heatmaps(t.to(_device()), original_image_size, soft=None, apply_sigmoid=True)
    return r.kp_hard.cpu().numpy(), r.score.cpu().numpy()


def decode_argmax(heatmaps, frame_size):
    """The inline loop of DIP_REAL.py:116-124 / model/MvRoPose_FR3.py:299-304: (K,h,w) ->
    np.float64[K,2] of [x * (frame_w / w), y * (frame_h / h)]."""
    t = heatmaps if isinstance(heatmaps, torch.Tensor) else torch.as_tensor(np.asarray(heatmaps))
    h, w = t.shape[-2:]
    fh, fw = frame_size
    idx = ops.decode_heatmaps(t.to(_device()), None, soft=None).idx.cpu().numpy().astype(np.int64)
    return np.stack([(idx % w) * (fw / w), (idx // w) * (fh / h)], axis=-1)


# ------------------------------------------------------------------------------------- FK
def _fk_single(chain: Chain, robot: str, joint_angles, selected_view):
    q = np.asarray([float(a) for a in joint_angles][: chain.n_joints], dtype=np.float32).reshape(1, -1)
    Rv = np.asarray(view_rotation(robot, selected_view), dtype=np.float32)[None]
    X = ops.forward_kinematics(chain, torch.from_numpy(q).to(_device()), Rv)
    return X[0, 0].cpu().numpy()


def _project_single(coords_3d, rvec, tvec, camera_matrix, dist_coeffs):
    K = np.asarray(camera_matrix, dtype=np.float64).reshape(1, 3, 3)
    d = np.zeros((1, 5)) if dist_coeffs is None else np.asarray(dist_coeffs, dtype=np.float64).reshape(-1)[:5][None]
    rec = dict(rvec_x=rvec[0], rvec_y=rvec[1], rvec_z=rvec[2], tvec_x=tvec[0], tvec_y=tvec[1], tvec_z=tvec[2])
    rig = CameraRig.from_aruco([rec], K, d)
    X = torch.as_tensor(np.asarray(coords_3d, dtype=np.float32).reshape(1, -1, 3)).to(_device())
    return ops.project_points(X, rig)[0, 0].cpu().numpy()


def _estimate_camera_pose(robot, gate, predicted_angles, predicted_heatmaps, camera_matrix, dist_coeffs, selected_view,
                          original_image_size, confidence_threshold):
    """estimate_camera_pose (model/Fr5_model_train.ipynb:4707-4753; FR3 variant with the 0.5-5 m plausibility gate,
    model/Franka_research3_model_train.ipynb:3667-3708): FK of the predicted angles -> 3-D object points; decode the
    predicted heat-maps -> 2-D key-points + sigmoid scores; keep score >= threshold (refuse below 4 points, :4728);
    pose by consensus PnP (mvgeo_pnp_solve in place of cv2.solvePnPRansac(..., SOLVEPNP_EPNP), :4735-4741).
    Returns (rvec (3,1) float64, tvec (3,1) float64, object_points_3d np.float32[K,3], image_points_2d np.float32[K,2]);
    rvec = tvec = None when the reference would return None. One device round trip."""
    dev = _device()
    chain = Chain.builtin(robot)
    ang = predicted_angles.detach().cpu().numpy() if isinstance(predicted_angles, torch.Tensor) else np.asarray(predicted_angles)
    q = torch.as_tensor(np.asarray(ang, dtype=np.float32).reshape(-1)[: chain.n_joints].reshape(1, -1)).to(dev)
    Rv = np.asarray(view_rotation(robot, selected_view), dtype=np.float32)[None]
    X = ops.forward_kinematics(chain, q, Rv)[:, 0]                                   # (1,K,3)
    hm = predicted_heatmaps if isinstance(predicted_heatmaps, torch.Tensor) else torch.as_tensor(np.asarray(predicted_heatmaps))
    r = ops.decode_heatmaps(hm.detach().to(dev), original_image_size, soft=None, apply_sigmoid=True)
    Kc = np.asarray(camera_matrix, dtype=np.float64).reshape(1, 3, 3)
    d = np.zeros((1, 5)) if dist_coeffs is None else np.asarray(dist_coeffs, dtype=np.float64).reshape(-1)[:5][None]
    rig = CameraRig(Kc, d, np.eye(3)[None], np.zeros((1, 3)))
    rvec, tvec, rms, status, inl = ops.pnp_solve(X, r.kp_hard[None, None], rig, r.score[None, None],
                                                 min_weight=float(confidence_threshold))
    packed = torch.cat([X.reshape(-1), r.kp_hard.reshape(-1), rvec.reshape(-1), tvec.reshape(-1),
                        status.reshape(-1).float()]).cpu().numpy()                  # ONE device-to-host copy
    K = chain.n_points
    obj = packed[: 3 * K].reshape(K, 3).astype(np.float32)
    img = packed[3 * K: 5 * K].reshape(K, 2).astype(np.float32)
    rv, tv, st = packed[5 * K: 5 * K + 3], packed[5 * K + 3: 5 * K + 6], int(packed[5 * K + 6])
    if not (st & 1) or (gate and not (st & 4)):
        return None, None, obj, img
    return rv.astype(np.float64).reshape(3, 1), tv.astype(np.float64).reshape(3, 1), obj, img


def _make_fr3():
    chain = None

    def angle_to_joint_coordinate(joint_angles, selected_view):
        """model/MvRoPose_FR3.py:90-131: 7 angles [rad] -> np.float32[8,3] (base + 7 frames)."""
        nonlocal chain
        chain = chain or Chain.builtin("fr3")
        return _fk_single(chain, "fr3", joint_angles, selected_view)

    def joint_coordinate_to_pixel_plane(coords_3d, aruco_result, camera_matrix, dist_coeffs):
        """model/MvRoPose_FR3.py:133-141: rvec in radians."""
        r = [aruco_result["rvec_x"], aruco_result["rvec_y"], aruco_result["rvec_z"]]
        t = [aruco_result["tvec_x"], aruco_result["tvec_y"], aruco_result["tvec_z"]]
        return _project_single(coords_3d, r, t, camera_matrix, dist_coeffs)

    def estimate_camera_pose(predicted_angles, predicted_heatmaps, camera_matrix, dist_coeffs, selected_view,
                             original_image_size, confidence_threshold=0.1):
        """model/Franka_research3_model_train.ipynb:3667-3708 (with the 0.5 m < |t| < 5 m gate, :3696-3701)."""
        return _estimate_camera_pose("fr3", True, predicted_angles, predicted_heatmaps, camera_matrix, dist_coeffs,
                                     selected_view, original_image_size, confidence_threshold)

    return SimpleNamespace(angle_to_joint_coordinate=angle_to_joint_coordinate,
                           joint_coordinate_to_pixel_plane=joint_coordinate_to_pixel_plane,
                           estimate_camera_pose=estimate_camera_pose)


def _make_fr5():
    chain = None

    def angle_to_joint_coordinate(joint_angles, selected_view):
        """model/Fr5_model_train.ipynb:256-288: 6 angles [DEGREES] -> np.float32[7,3]."""
        nonlocal chain
        chain = chain or Chain.builtin("fr5")
        return _fk_single(chain, "fr5", joint_angles, selected_view)

    def joint_coordinate_to_pixel_plane(joint_coords, aruco_result, camera_matrix, dist_coeffs):
        """model/Fr5_model_train.ipynb:290-305: rvec stored in DEGREES, cast to float32."""
        r = np.array([math.radians(aruco_result[k]) for k in ("rvec_x", "rvec_y", "rvec_z")], dtype=np.float32)
        t = np.array([aruco_result[k] for k in ("tvec_x", "tvec_y", "tvec_z")], dtype=np.float32)
        return _project_single(joint_coords, r, t, camera_matrix, dist_coeffs)

    def estimate_camera_pose(predicted_angles, predicted_heatmaps, camera_matrix, dist_coeffs, selected_view,
                             original_image_size, confidence_threshold=0.1):
        """model/Fr5_model_train.ipynb:4707-4753 (no distance gate in this variant)."""
        return _estimate_camera_pose("fr5", False, predicted_angles, predicted_heatmaps, camera_matrix, dist_coeffs,
                                     selected_view, original_image_size, confidence_threshold)

    return SimpleNamespace(angle_to_joint_coordinate=angle_to_joint_coordinate,
                           joint_coordinate_to_pixel_plane=joint_coordinate_to_pixel_plane,
                           estimate_camera_pose=estimate_camera_pose)


def _make_meca500():
    chain = None

    def forward_kinematics(joint_angles):
        """visualization/Meca500_vis.ipynb:62-82: 6 angles [DEGREES] -> np.float32[7,3]."""
        nonlocal chain
        chain = chain or Chain.builtin("meca500")
        return _fk_single(chain, "meca500", joint_angles, None)

    def project_to_pixel(coords_3d, rvec, tvec, camera_matrix, dist_coeffs):
        """visualization/Meca500_vis.ipynb:84-87 (= visualization/Fr5_vis.ipynb:111-115): rvec in radians."""
        return _project_single(coords_3d, np.asarray(rvec).reshape(3), np.asarray(tvec).reshape(3), camera_matrix,
                               dist_coeffs)

    return SimpleNamespace(forward_kinematics=forward_kinematics, project_to_pixel=project_to_pixel)


fr3 = _make_fr3()
fr5 = _make_fr5()
meca500 = _make_meca500()


# ------------------------------------------------------------- MV-model.ipynb prototype API
class ForwardKinematics:
    """model/MV-model.ipynb:841-874. dh_params: list of (theta0, d, a, alpha) in radians.
    forward(angles (B,J) tensor) -> torch.float32 (B,J,3) on the CPU (joints only, no base)."""

    def __init__(self, dh_params):
        self.dh_params = dh_params
        th0, d, a, al = (list(c) for c in zip(*dh_params))
        self._chain = Chain.from_dh(a, d, al, th0, convention="standard", angle_scale=1.0, emit_base=False)

    def forward(self, angles):
        q = torch.as_tensor(angles, dtype=torch.float32).to(_device())
        return ops.forward_kinematics(self._chain, q)[:, 0].cpu()


def project_3d_to_2d(joint_3d, camera_matrix, dist_coeffs=None, rvec=None, tvec=None):
    """model/MV-model.ipynb:879-899: (B,J,3) -> torch.float32 (B,J,2); rvec / tvec are per-batch
    lists (None = zero rotation / translation). The per-item poses become the "views" of one rig, so the whole
    batch is ONE launch (chunks of 16 items: MVGEO_MAX_VIEWS) instead of one launch + sync per item."""
    X = torch.as_tensor(np.asarray(joint_3d), dtype=torch.float32)
    Bn = int(X.shape[0])
    K = np.asarray(camera_matrix, dtype=np.float64).reshape(3, 3)
    d = np.zeros(5) if dist_coeffs is None else np.asarray(dist_coeffs, dtype=np.float64).reshape(-1)[:5]
    dev = _device()
    Xd = X.to(dev)
    out = torch.empty((Bn, X.shape[1], 2), dtype=torch.float32, device=dev)
    for b0 in range(0, Bn, 16):
        n = min(16, Bn - b0)
        recs = []
        for b in range(b0, b0 + n):
            r = np.zeros(3) if rvec is None else np.asarray(rvec[b], dtype=np.float64).reshape(3)
            t = np.zeros(3) if tvec is None else np.asarray(tvec[b], dtype=np.float64).reshape(3)
            recs.append(dict(rvec_x=r[0], rvec_y=r[1], rvec_z=r[2], tvec_x=t[0], tvec_y=t[1], tvec_z=t[2]))
        rig = CameraRig.from_aruco(recs, np.broadcast_to(K, (n, 3, 3)), np.broadcast_to(d, (n, 5)))
        out[b0:b0 + n] = ops.project_points(Xd[b0:b0 + n].unsqueeze(0).contiguous(), rig)[0]   # X per "view"
    return out.cpu()


def robot_pose_loss(pred, gt_keypoints=None, gt_angles=None, lambda_kp=1.0, lambda_angle=1.0, lambda_fk=1.0, *,
                    chain=None, cams=None, R_view=None):
    """model/MV-model.ipynb:942-950, unchanged in form: key-point MSE + angle MSE + FK-consistency MSE on
    pred['proj_2d'] (which the reference computes detached, MV-model.ipynb:874,899).
    Extension (keyword-only, absent in the reference): when pred has no 'proj_2d' and a `chain` plus one camera
    (`cams`: CameraRig with a single view) are given, the FK term is computed from pred['angles'] by
    mvgeo_fk_reproj_fwd/bwd — same value, but DIFFERENTIABLE in the angles."""
    import torch.nn.functional as F

    loss = 0.0
    if gt_keypoints is not None:
        loss = loss + lambda_kp * F.mse_loss(pred["keypoints_2d"], gt_keypoints)
    if gt_angles is not None:
        loss = loss + lambda_angle * F.mse_loss(pred["angles"], gt_angles)
    if gt_keypoints is not None and pred.get("proj_2d") is not None:
        loss = loss + lambda_fk * F.mse_loss(pred["proj_2d"], gt_keypoints)
    elif gt_keypoints is not None and chain is not None and cams is not None:
        dev = _device()
        q = pred["angles"].to(dev, dtype=torch.float32)
        gt = gt_keypoints.to(dev, dtype=torch.float32).unsqueeze(1)                   # (B,1,J,2)
        fk, _, _, _ = ops.fk_reproj_loss(chain, q, cams, gt.contiguous(), R_view, lam=float(lambda_fk))
        loss = loss + fk.to(pred["angles"].device)
    return loss


def create_gt_heatmap(keypoint_2d, HEATMAP_SIZE, sigma):
    """model/MvRoPose_FR3.py:65-73: (x, y) in map pixels -> float64 (H,W) array (the kernel
    evaluates in float32; the reference's float64 values are matched to ~1e-6 absolute)."""
    kp = torch.tensor([[float(keypoint_2d[0]), float(keypoint_2d[1])]], dtype=torch.float32, device=_device())
    return ops.encode_gaussian(kp, HEATMAP_SIZE, float(sigma))[0].cpu().numpy().astype(np.float64)
