"""Batched, tensor-native front end of the geometry hot path.

PyTorch is plumbing only (device memory, streams, autograd hooks): every function hands raw
device pointers and the current CUDA stream to libmvgeo.so (include/mvgeo.h) through ctypes.
Nothing here computes on the CPU and nothing falls back to torch ops; CPU tensors are
rejected (use `compat` for the reference's single-frame, NumPy-returning signatures, which
stage to the GPU explicitly).
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .rig import CameraRig
from .robots import Chain

_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
_SOFT = {"none": _lib.SOFT_NONE, None: _lib.SOFT_NONE, "global": _lib.SOFT_GLOBAL, "window": _lib.SOFT_WINDOW}

DecodeResult = namedtuple("DecodeResult", "idx peak score kp_hard kp_soft")


def _need_cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: this package has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _scales(image_size, H, W):
    if image_size is None:
        return 1.0, 1.0
    h_img, w_img = image_size
    return w_img / W, h_img / H  # Python float division, as `original_w / w` in the reference


def _view_ptrs(views: Sequence[torch.Tensor]):
    """Host array of V device pointers (mvgeo_decode_views / mvgeo_pipeline_views)."""
    return (C.c_void_p * len(views))(*[v.data_ptr() for v in views])


def _as_view_list(maps):
    """dict view -> (B,K,H,W) (the reference network's output, model/MvRoPose_FR3.py:625; insertion
    order = view order), list / tuple of such tensors, or None when `maps` is one stacked tensor."""
    if isinstance(maps, torch.Tensor):
        return None
    views = list(maps.values()) if isinstance(maps, dict) else list(maps)
    views = [_need_cuda(v, "maps[view]") for v in views]
    if not views or any(v.shape != views[0].shape or v.dtype != views[0].dtype or v.dim() != 4 or v.device != views[0].device
                        for v in views):
        raise ValueError("a view list must hold V tensors of identical shape (B, K, H, W), dtype and device")
    return views


def cameras_to_device(rig: Union[CameraRig, torch.Tensor], device) -> torch.Tensor:
    """(V,24) float32 device tensor of mvgeo_camera records."""
    if isinstance(rig, torch.Tensor):
        return _need_cuda(rig, "cams", torch.float32)
    return torch.from_numpy(rig.packed()).to(device)


# ---------------------------------------------------------------------------------- decode
def decode_heatmaps(maps, image_size=None, *, soft: Optional[str] = "global", beta: float = 100.0,
                    window_radius: int = 3, apply_sigmoid: bool = False) -> DecodeResult:
    """Arg-max + sub-pixel soft-arg-max of belief maps.

    maps: CUDA tensor (..., H, W) in fp32 / bf16 / fp16, or a dict / sequence of V tensors (B, K, H, W)
    (the reference's dict view -> heat-maps, model/MvRoPose_FR3.py:625), decoded in ONE launch
    without a stack copy, in which case results are (B, V, K). image_size = (H_img, W_img) scales key-points to image pixels
    like extract_keypoints_from_heatmaps (model/Fr5_model_train.ipynb:4701-4702).
    Returns DecodeResult(idx int32, peak, score, kp_hard (...,2), kp_soft (...,2))."""
    lib = _lib.load()
    mode = _SOFT[soft]
    views: Sequence[torch.Tensor]
    if isinstance(maps, torch.Tensor):
        t = _need_cuda(maps, "maps")
        if t.dim() < 2:
            raise ValueError("maps must have at least 2 dimensions (H, W)")
        views, lead, strided = [t], tuple(t.shape[:-2]), False
    else:
        views = _as_view_list(maps)
        Bv, Kv = views[0].shape[:2]
        lead, strided = (Bv, len(views), Kv), True
    t0 = views[0]
    if t0.dtype not in _DTYPES:
        raise TypeError(f"unsupported belief-map dtype {t0.dtype}")
    H, W = int(t0.shape[-2]), int(t0.shape[-1])
    sx, sy = _scales(image_size, H, W)
    dev = t0.device
    n_out = int(np.prod(lead)) if lead else 1
    idx = torch.empty(lead, dtype=torch.int32, device=dev)
    peak = torch.empty(lead, dtype=torch.float32, device=dev)
    score = torch.empty(lead, dtype=torch.float32, device=dev)
    kp_hard = torch.empty(lead + (2,), dtype=torch.float32, device=dev)
    kp_soft = torch.empty(lead + (2,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _stream(dev)
        if strided:  # V per-view tensors: ONE launch walks all of them, results land in (B, V, K) order
            _lib.check(lib.mvgeo_decode_views(_view_ptrs(views), len(views), _DTYPES[t0.dtype], lead[0], lead[2], H, W,
                                              sx, sy, mode, float(beta), int(window_radius), int(bool(apply_sigmoid)),
                                              idx.data_ptr(), peak.data_ptr(), score.data_ptr(), kp_hard.data_ptr(),
                                              kp_soft.data_ptr(), st), "mvgeo_decode_views")
        else:
            _lib.check(lib.mvgeo_decode(t0.data_ptr(), _DTYPES[t0.dtype], t0.numel() // (H * W), H, W, sx, sy, mode,
                                        float(beta), int(window_radius), int(bool(apply_sigmoid)), 1, 1, 0,
                                        idx.data_ptr(), peak.data_ptr(), score.data_ptr(), kp_hard.data_ptr(),
                                        kp_soft.data_ptr(), st), "mvgeo_decode")
    assert n_out == idx.numel()
    return DecodeResult(idx, peak, score, kp_hard, kp_soft)


# ---------------------------------------------------------------------------- triangulation
def triangulate(kp: torch.Tensor, P: torch.Tensor, w: Optional[torch.Tensor] = None, *, min_weight: float = 0.0,
                weighted: bool = False):
    """Batched DLT. kp (B,V,K,2) pixels, P (V,3,4), w (B,V,K) or None ->
    X (B,K,3) [NaN where fewer than two views are valid], resid (B,K) px RMS, n_views (B,K) int32."""
    lib = _lib.load()
    kp = _need_cuda(kp, "kp", torch.float32)
    P = _need_cuda(P, "P", torch.float32)
    if kp.dim() != 4 or kp.shape[-1] != 2:
        raise ValueError("kp must be (B, V, K, 2)")
    B, V, K, _ = kp.shape
    if tuple(P.shape) != (V, 3, 4):
        raise ValueError(f"P must be ({V}, 3, 4)")
    if w is not None:
        w = _need_cuda(w, "w", torch.float32)
        if tuple(w.shape) != (B, V, K):
            raise ValueError("w must be (B, V, K)")
    dev = kp.device
    X = torch.empty((B, K, 3), dtype=torch.float32, device=dev)
    resid = torch.empty((B, K), dtype=torch.float32, device=dev)
    nv = torch.empty((B, K), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mvgeo_triangulate(kp.data_ptr(), _ptr(w), P.data_ptr(), B, V, K, float(min_weight),
                                         int(bool(weighted)), X.data_ptr(), resid.data_ptr(), nv.data_ptr(),
                                         _stream(dev)), "mvgeo_triangulate")
    return X, resid, nv


def quat_mean(q: torch.Tensor, w: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched average_quaternion (dataset/Fr5_preprocessing.py:57-65): q (G, N, 4) -> (G, 4) unit
    quaternions (largest eigenvector of sum w q q^T), sign in the hemisphere of q[:, 0]."""
    lib = _lib.load()
    q = _need_cuda(q, "q", torch.float32)
    if q.dim() != 3 or q.shape[-1] != 4:
        raise ValueError("q must be (G, N, 4)")
    if w is not None:
        w = _need_cuda(w, "w", torch.float32)
        if tuple(w.shape) != tuple(q.shape[:2]):
            raise ValueError("w must be (G, N)")
    out = torch.empty((q.shape[0], 4), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.mvgeo_quat_mean(q.data_ptr(), _ptr(w), q.shape[0], q.shape[1], out.data_ptr(), _stream(q.device)),
                   "mvgeo_quat_mean")
    return out


# ------------------------------------------------------------------------ FK and projection
def _view_rot(R_view, dev, V_expected=None):
    if R_view is None:
        return None
    if not isinstance(R_view, torch.Tensor):
        R_view = torch.as_tensor(np.asarray(R_view, dtype=np.float32))
    R_view = R_view.to(device=dev, dtype=torch.float32).contiguous()
    if R_view.dim() != 3 or R_view.shape[1:] != (3, 3):
        raise ValueError("R_view must be (V, 3, 3)")
    if V_expected is not None and R_view.shape[0] != V_expected:
        raise ValueError(f"R_view must hold {V_expected} rotations")
    return R_view


def forward_kinematics(chain: Chain, q: torch.Tensor, R_view=None) -> torch.Tensor:
    """q (B, J) in the chain's native unit (radians for FR3, DEGREES for Fr5 / Meca500) ->
    X (B, V, K, 3) with the per-view base rotation applied (V = 1 when R_view is None)."""
    lib = _lib.load()
    q = _need_cuda(q, "q", torch.float32)
    if q.dim() != 2 or q.shape[1] != chain.n_joints:
        raise ValueError(f"q must be (B, {chain.n_joints})")
    dev = q.device
    Rv = _view_rot(R_view, dev)
    V = 1 if Rv is None else int(Rv.shape[0])
    X = torch.empty((q.shape[0], V, chain.n_points, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mvgeo_fk(C.byref(chain.struct), q.data_ptr(), q.shape[0], _ptr(Rv), V, X.data_ptr(),
                                _stream(dev)), "mvgeo_fk")
    return X


def project_points(X: torch.Tensor, cams) -> torch.Tensor:
    """cv2.projectPoints for a rig. X (B,K,3) [same points for every camera] or (B,V,K,3);
    cams: CameraRig or packed (V,24) device tensor -> uv (B,V,K,2)."""
    lib = _lib.load()
    X = _need_cuda(X, "X", torch.float32)
    dev = X.device
    cams_t = cameras_to_device(cams, dev)
    V = int(cams_t.shape[0])
    if X.dim() == 3:
        per_view, B, K = 0, X.shape[0], X.shape[1]
    elif X.dim() == 4 and X.shape[1] == V:
        per_view, B, K = 1, X.shape[0], X.shape[2]
    else:
        raise ValueError("X must be (B,K,3) or (B,V,K,3) with V matching the rig")
    uv = torch.empty((B, V, K, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mvgeo_project(X.data_ptr(), per_view, cams_t.data_ptr(), B, V, K, uv.data_ptr(), _stream(dev)),
                   "mvgeo_project")
    return uv


def undistort_points(kp: torch.Tensor, cams, iters: int = 5) -> torch.Tensor:
    """cv2.undistortPoints(kp, K, dist, P=K) per view: kp (B,V,K,2) pixels of the distorted image ->
    pixels of the ideal pinhole camera (what cv2.undistort does to whole images, MvRoPose_FR3.py:212)."""
    lib = _lib.load()
    kp = _need_cuda(kp, "kp", torch.float32)
    cams_t = cameras_to_device(cams, kp.device)
    V = int(cams_t.shape[0])
    if kp.dim() != 4 or kp.shape[1] != V or kp.shape[-1] != 2:
        raise ValueError("kp must be (B, V, K, 2) with V matching the rig")
    out = torch.empty_like(kp)
    with torch.cuda.device(kp.device):
        _lib.check(lib.mvgeo_undistort_points(kp.data_ptr(), cams_t.data_ptr(), kp.shape[0], V, kp.shape[2], int(iters),
                                              out.data_ptr(), _stream(kp.device)), "mvgeo_undistort_points")
    return out


class _FKReprojLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, chain, cams_t, gt_uv, R_view, w, lam):
        lib = _lib.load()
        dev = q.device
        B, V, K = q.shape[0], int(cams_t.shape[0]), chain.n_points
        X = torch.empty((B, V, K, 3), dtype=torch.float32, device=dev)
        uv = torch.empty((B, V, K, 2), dtype=torch.float32, device=dev)
        frame_loss = torch.empty((B,), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mvgeo_fk_reproj_fwd(C.byref(chain.struct), q.data_ptr(), B, _ptr(R_view), cams_t.data_ptr(),
                                               V, gt_uv.data_ptr(), _ptr(w), float(lam), X.data_ptr(), uv.data_ptr(),
                                               frame_loss.data_ptr(), loss.data_ptr(), _stream(dev)),
                       "mvgeo_fk_reproj_fwd")
        ctx.save_for_backward(q, cams_t, gt_uv, R_view, w)
        ctx.chain, ctx.lam = chain, float(lam)
        ctx.mark_non_differentiable(X, uv, frame_loss)
        return loss, X, uv, frame_loss

    @staticmethod
    def backward(ctx, g_loss, _gx, _guv, _gfl):
        lib = _lib.load()
        q, cams_t, gt_uv, R_view, w = ctx.saved_tensors
        dev = q.device
        dq = torch.empty_like(q)
        g = g_loss.to(dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.mvgeo_fk_reproj_bwd(C.byref(ctx.chain.struct), q.data_ptr(), q.shape[0], _ptr(R_view),
                                               cams_t.data_ptr(), int(cams_t.shape[0]), gt_uv.data_ptr(), _ptr(w),
                                               ctx.lam, g.data_ptr(), dq.data_ptr(), _stream(dev)),
                       "mvgeo_fk_reproj_bwd")
        return dq, None, None, None, None, None, None


def fk_reproj_loss(chain: Chain, q: torch.Tensor, cams, gt_uv: torch.Tensor, R_view=None,
                   w: Optional[torch.Tensor] = None, lam: float = 1.0):
    """Differentiable FK-consistency loss: lam * mean_{b,v,k,c} w (project(FK(q)) - gt_uv)^2
    (robot_pose_loss FK term, model/MV-model.ipynb:949, made differentiable in q).
    Returns (loss [scalar, differentiable w.r.t. q], X (B,V,K,3), uv (B,V,K,2), frame_loss (B,))."""
    q = _need_cuda(q, "q", torch.float32)
    dev = q.device
    cams_t = cameras_to_device(cams, dev)
    V, K = int(cams_t.shape[0]), chain.n_points
    gt_uv = _need_cuda(gt_uv, "gt_uv", torch.float32)
    if q.dim() != 2 or q.shape[1] != chain.n_joints:
        raise ValueError(f"q must be (B, {chain.n_joints})")
    if tuple(gt_uv.shape) != (q.shape[0], V, K, 2):
        raise ValueError(f"gt_uv must be ({q.shape[0]}, {V}, {K}, 2)")
    if w is not None:
        w = _need_cuda(w, "w", torch.float32)
        if tuple(w.shape) != (q.shape[0], V, K):
            raise ValueError("w must be (B, V, K)")
    Rv = _view_rot(R_view, dev, V)
    return _FKReprojLoss.apply(q, chain, cams_t, gt_uv, Rv, w, lam)


# ------------------------------------------------------------------- camera-pose refinement
def pnp_refine(X: torch.Tensor, kp: torch.Tensor, cams, w: Optional[torch.Tensor] = None, *, min_weight: float = 0.0,
               max_iters: int = 20):
    """Batched camera-pose refinement (Levenberg-Marquardt on the reprojection error) per
    (frame, view), started from the prior pose held in `cams` — the batched stand-in for
    estimate_camera_pose's cv2.solvePnPRansac + fallback-to-prior (model/Fr5_model_train.ipynb:4707-4753).
    X (B,K,3) or (B,V,K,3) object points; kp (B,V,K,2) image points; w (B,V,K) scores or None.
    Returns rvec (B,V,3), tvec (B,V,3), rms (B,V) px, status (B,V) int32
    (bit 0 solved with >= 4 points, bit 1 converged, bit 2 plausible 0.5 m < |t| < 5 m)."""
    lib = _lib.load()
    X = _need_cuda(X, "X", torch.float32)
    kp = _need_cuda(kp, "kp", torch.float32)
    dev = X.device
    cams_t = cameras_to_device(cams, dev)
    V = int(cams_t.shape[0])
    if kp.dim() != 4 or kp.shape[1] != V or kp.shape[-1] != 2:
        raise ValueError("kp must be (B, V, K, 2) with V matching the rig")
    B, _, K, _ = kp.shape
    if X.dim() == 3 and tuple(X.shape) == (B, K, 3):
        per_view = 0
    elif X.dim() == 4 and tuple(X.shape) == (B, V, K, 3):
        per_view = 1
    else:
        raise ValueError("X must be (B,K,3) or (B,V,K,3)")
    if w is not None:
        w = _need_cuda(w, "w", torch.float32)
        if tuple(w.shape) != (B, V, K):
            raise ValueError("w must be (B, V, K)")
    rvec = torch.empty((B, V, 3), dtype=torch.float32, device=dev)
    tvec = torch.empty((B, V, 3), dtype=torch.float32, device=dev)
    rms = torch.empty((B, V), dtype=torch.float32, device=dev)
    status = torch.empty((B, V), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mvgeo_pnp_refine(X.data_ptr(), per_view, kp.data_ptr(), _ptr(w), cams_t.data_ptr(), B, V, K,
                                        float(min_weight), int(max_iters), rvec.data_ptr(), tvec.data_ptr(),
                                        rms.data_ptr(), status.data_ptr(), _stream(dev)), "mvgeo_pnp_refine")
    return rvec, tvec, rms, status


def pnp_solve(X: torch.Tensor, kp: torch.Tensor, cams, w: Optional[torch.Tensor] = None, *, min_weight: float = 0.0,
              reproj_thresh: float = 8.0, max_iters: int = 30):
    """Batched camera pose WITHOUT a prior per (frame, view): every key-point triplet is a P3P hypothesis,
    consensus at `reproj_thresh` pixels, Levenberg-Marquardt on the inliers — the batched replacement of
    cv2.solvePnPRansac(..., flags=cv2.SOLVEPNP_EPNP) in estimate_camera_pose (model/Fr5_model_train.ipynb:4735-4741).
    Only the intrinsics / distortion of `cams` are used. Shapes as pnp_refine.
    Returns rvec (B,V,3), tvec (B,V,3), rms (B,V), status (B,V) int32 (0 = refused: < 4 valid points or < 4
    inliers, pose NaN), inliers (B,V) int32 bit mask."""
    lib = _lib.load()
    X = _need_cuda(X, "X", torch.float32)
    kp = _need_cuda(kp, "kp", torch.float32)
    dev = X.device
    cams_t = cameras_to_device(cams, dev)
    V = int(cams_t.shape[0])
    if kp.dim() != 4 or kp.shape[1] != V or kp.shape[-1] != 2:
        raise ValueError("kp must be (B, V, K, 2) with V matching the rig")
    B, _, K, _ = kp.shape
    if X.dim() == 3 and tuple(X.shape) == (B, K, 3):
        per_view = 0
    elif X.dim() == 4 and tuple(X.shape) == (B, V, K, 3):
        per_view = 1
    else:
        raise ValueError("X must be (B,K,3) or (B,V,K,3)")
    if w is not None:
        w = _need_cuda(w, "w", torch.float32)
        if tuple(w.shape) != (B, V, K):
            raise ValueError("w must be (B, V, K)")
    rvec = torch.empty((B, V, 3), dtype=torch.float32, device=dev)
    tvec = torch.empty((B, V, 3), dtype=torch.float32, device=dev)
    rms = torch.empty((B, V), dtype=torch.float32, device=dev)
    status = torch.empty((B, V), dtype=torch.int32, device=dev)
    inliers = torch.empty((B, V), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mvgeo_pnp_solve(X.data_ptr(), per_view, kp.data_ptr(), _ptr(w), cams_t.data_ptr(), B, V, K,
                                       float(min_weight), float(reproj_thresh), int(max_iters), rvec.data_ptr(),
                                       tvec.data_ptr(), rms.data_ptr(), status.data_ptr(), inliers.data_ptr(),
                                       _stream(dev)), "mvgeo_pnp_solve")
    return rvec, tvec, rms, status, inliers


# --------------------------------------------------------------- GT encoder and heat-map MSE
def encode_gaussian(kp: torch.Tensor, heatmap_size, sigma: float, dtype=torch.float32) -> torch.Tensor:
    """Batched create_gt_heatmap (model/MvRoPose_FR3.py:65-73): kp (..., 2) in MAP pixels ->
    maps (..., H, W) of `dtype`."""
    lib = _lib.load()
    kp = _need_cuda(kp, "kp", torch.float32)
    H, W = int(heatmap_size[0]), int(heatmap_size[1])
    lead = tuple(kp.shape[:-1])
    maps = torch.empty(lead + (H, W), dtype=dtype, device=kp.device)
    with torch.cuda.device(kp.device):
        _lib.check(lib.mvgeo_encode_gaussian(kp.data_ptr(), kp.numel() // 2, H, W, float(sigma), _DTYPES[dtype],
                                             maps.data_ptr(), _stream(kp.device)), "mvgeo_encode_gaussian")
    return maps


class _HeatmapMSE(torch.autograd.Function):
    """Forward reads the maps once (loss only); backward reads them again and writes
    upstream * d loss / d pred in the same kernel — no materialised targets, no extra scaling pass."""

    @staticmethod
    def _call(pred, kp, sigma, weight, dloss, grad):
        lib = _lib.load()
        dev = pred.device
        H, W = int(pred.shape[-2]), int(pred.shape[-1])
        n_maps = pred.numel() // (H * W)
        partial = torch.empty((n_maps,), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mvgeo_heatmap_mse(pred.data_ptr(), _DTYPES[pred.dtype], kp.data_ptr(), n_maps, H, W,
                                             float(sigma), float(weight), _ptr(dloss), partial.data_ptr(),
                                             loss.data_ptr(), _ptr(grad), _stream(dev)), "mvgeo_heatmap_mse")
        return loss

    @staticmethod
    def forward(ctx, pred, kp, sigma, weight):
        ctx.save_for_backward(pred, kp)
        ctx.sigma, ctx.weight = float(sigma), float(weight)
        return _HeatmapMSE._call(pred, kp, sigma, weight, None, None)

    @staticmethod
    def backward(ctx, g):
        pred, kp = ctx.saved_tensors
        grad = torch.empty_like(pred)
        _HeatmapMSE._call(pred, kp, ctx.sigma, ctx.weight, g.to(dtype=torch.float32).contiguous(), grad)
        return grad, None, None, None


def heatmap_mse_loss(pred: torch.Tensor, kp: torch.Tensor, sigma: float, weight: float = 1.0) -> torch.Tensor:
    """nn.MSELoss()(pred, gaussian_targets(kp)) * weight (model/MvRoPose_FR3.py:846-847) without
    materialising the targets. pred (..., H, W); kp (..., 2) in map pixels (NaN -> zero target)."""
    pred = _need_cuda(pred, "pred")
    kp = _need_cuda(kp, "kp", torch.float32)
    if pred.dtype not in _DTYPES:
        raise TypeError(f"unsupported dtype {pred.dtype}")
    if tuple(kp.shape) != tuple(pred.shape[:-2]) + (2,):
        raise ValueError("kp must be pred.shape[:-2] + (2,)")
    return _HeatmapMSE.apply(pred, kp, sigma, weight)


class _DecodeMSE(torch.autograd.Function):
    """Forward: ONE read of the prediction gives the loss and the hard decode (mvgeo_decode_mse). Backward: the
    gradient pass of mvgeo_heatmap_mse (reads the prediction, writes the gradient, upstream scalar applied on the device)."""

    @staticmethod
    def forward(ctx, pred, kp, sigma, weight, sx, sy, apply_sigmoid):
        lib = _lib.load()
        dev = pred.device
        H, W = int(pred.shape[-2]), int(pred.shape[-1])
        lead = tuple(pred.shape[:-2])
        n_maps = pred.numel() // (H * W)
        idx = torch.empty(lead, dtype=torch.int32, device=dev)
        peak = torch.empty(lead, dtype=torch.float32, device=dev)
        score = torch.empty(lead, dtype=torch.float32, device=dev)
        kp_hard = torch.empty(lead + (2,), dtype=torch.float32, device=dev)
        partial = torch.empty((n_maps,), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mvgeo_decode_mse(pred.data_ptr(), _DTYPES[pred.dtype], n_maps, H, W, sx, sy, int(bool(apply_sigmoid)),
                                            kp.data_ptr(), float(sigma), float(weight), idx.data_ptr(), peak.data_ptr(),
                                            score.data_ptr(), kp_hard.data_ptr(), partial.data_ptr(), loss.data_ptr(),
                                            _stream(dev)), "mvgeo_decode_mse")
        ctx.save_for_backward(pred, kp)
        ctx.sigma, ctx.weight = float(sigma), float(weight)
        ctx.mark_non_differentiable(idx, peak, score, kp_hard)
        return loss, idx, peak, score, kp_hard

    @staticmethod
    def backward(ctx, g, *_unused):
        pred, kp = ctx.saved_tensors
        grad = torch.empty_like(pred)
        _HeatmapMSE._call(pred, kp, ctx.sigma, ctx.weight, g.to(dtype=torch.float32).contiguous(), grad)
        return grad, None, None, None, None, None, None


def decode_and_mse(pred: torch.Tensor, kp_target: torch.Tensor, sigma: float, weight: float = 1.0, image_size=None,
                   apply_sigmoid: bool = False):
    """One-read training step: heat-map MSE loss against Gaussian targets centred at
    kp_target (map pixels) AND the hard decode of the same prediction, the maps read from HBM once.
    Returns (loss [differentiable w.r.t. pred], DecodeResult with kp_soft = kp_hard).
    Falls back to nothing: maps the fused kernel cannot take raise ValueError (use heatmap_mse_loss + decode_heatmaps)."""
    pred = _need_cuda(pred, "pred")
    kp_target = _need_cuda(kp_target, "kp_target", torch.float32)
    if pred.dtype not in _DTYPES:
        raise TypeError(f"unsupported dtype {pred.dtype}")
    if tuple(kp_target.shape) != tuple(pred.shape[:-2]) + (2,):
        raise ValueError("kp_target must be pred.shape[:-2] + (2,)")
    H, W = int(pred.shape[-2]), int(pred.shape[-1])
    sx, sy = _scales(image_size, H, W)
    loss, idx, peak, score, kp_hard = _DecodeMSE.apply(pred, kp_target, sigma, weight, sx, sy, apply_sigmoid)
    return loss, DecodeResult(idx, peak, score, kp_hard, kp_hard)


# --------------------------------------------------------------------------- fused pipeline
def _make_cfg(dtype, H, W, V, K, image_size, soft, beta, window_radius, apply_sigmoid, tri_use_soft, tri_weighted,
              min_score, lam) -> _lib.PipelineCfg:
    sx, sy = _scales(image_size, H, W)
    return _lib.PipelineCfg(_DTYPES[dtype], H, W, V, K, _SOFT[soft], int(window_radius), int(bool(apply_sigmoid)),
                            int(bool(tri_use_soft)), int(bool(tri_weighted)), float(beta), float(min_score),
                            float(lam), sx, sy)


_OUT_SPECS = (  # name, dtype, shape builder (B, V, K)
    ("idx", torch.int32, lambda B, V, K: (B, V, K)),
    ("peak", torch.float32, lambda B, V, K: (B, V, K)),
    ("score", torch.float32, lambda B, V, K: (B, V, K)),
    ("kp_hard", torch.float32, lambda B, V, K: (B, V, K, 2)),
    ("kp_soft", torch.float32, lambda B, V, K: (B, V, K, 2)),
    ("X_tri", torch.float32, lambda B, V, K: (B, K, 3)),
    ("tri_resid", torch.float32, lambda B, V, K: (B, K)),
    ("tri_views", torch.int32, lambda B, V, K: (B, K)),
    ("X_fk", torch.float32, lambda B, V, K: (B, V, K, 3)),
    ("uv_fk", torch.float32, lambda B, V, K: (B, V, K, 2)),
    ("frame_loss", torch.float32, lambda B, V, K: (B,)),
    ("loss", torch.float32, lambda B, V, K: ()),
    ("ticket", torch.int32, lambda B, V, K: (1,)),  # zero on entry / exit: geometry tail in one launch
)


def alloc_outputs(B: int, V: int, K: int, device, pin: bool = False) -> dict:
    if pin:
        return {n: torch.zeros(f(B, V, K), dtype=dt).pin_memory() for n, dt, f in _OUT_SPECS}
    out = {n: torch.empty(f(B, V, K), dtype=dt, device=device) for n, dt, f in _OUT_SPECS}
    out["ticket"].zero_()
    return out


def _out_struct(out: dict) -> _lib.PipelineOut:
    return _lib.PipelineOut(*[out[n].data_ptr() if out.get(n) is not None else None for n, _, _ in _OUT_SPECS])


def pipeline(maps, P: torch.Tensor, chain: Chain, q: torch.Tensor, cams, R_view=None, *,
             image_size=None, soft: Optional[str] = "global", beta: float = 100.0, window_radius: int = 3,
             apply_sigmoid: bool = False, tri_use_soft: bool = True, tri_weighted: bool = False,
             min_score: float = 0.0, lam: float = 1.0, out: Optional[dict] = None) -> dict:
    """decode -> triangulate -> FK -> reprojection consistency in two launches on the current
    stream, no host synchronisation. maps: (B,V,K,H,W), or the reference's dict view -> (B,K,H,W) /
    a list of V such tensors (read in place through a pointer array: no torch.stack copy);
    P (V,3,4); q (B,J). Returns a dict of device tensors (see alloc_outputs); pass `out` to reuse
    buffers (e.g. under CUDA-graph capture)."""
    lib = _lib.load()
    views = _as_view_list(maps)
    if views is None:
        maps = _need_cuda(maps, "maps")
        if maps.dim() != 5:
            raise ValueError("maps must be (B, V, K, H, W)")
        B, V, K, H, W = (int(s) for s in maps.shape)
        dev, mdtype = maps.device, maps.dtype
    else:
        V = len(views)
        B, K, H, W = (int(s) for s in views[0].shape)
        dev, mdtype = views[0].device, views[0].dtype
    if mdtype not in _DTYPES:
        raise TypeError(f"unsupported belief-map dtype {mdtype}")
    P = _need_cuda(P, "P", torch.float32)
    q = _need_cuda(q, "q", torch.float32)
    if tuple(P.shape) != (V, 3, 4) or tuple(q.shape) != (B, chain.n_joints) or K != chain.n_points:
        raise ValueError("shape mismatch between maps, P, q and the chain")
    cams_t = cameras_to_device(cams, dev)
    Rv = _view_rot(R_view, dev, V)
    cfg = _make_cfg(mdtype, H, W, V, K, image_size, soft, beta, window_radius, apply_sigmoid, tri_use_soft,
                    tri_weighted, min_score, lam)
    if out is None:
        out = alloc_outputs(B, V, K, dev)
    o = _out_struct(out)
    with torch.cuda.device(dev):
        if views is None:
            _lib.check(lib.mvgeo_pipeline(C.byref(cfg), maps.data_ptr(), B, P.data_ptr(), C.byref(chain.struct),
                                          q.data_ptr(), _ptr(Rv), cams_t.data_ptr(), C.byref(o), _stream(dev)),
                       "mvgeo_pipeline")
        else:
            _lib.check(lib.mvgeo_pipeline_views(C.byref(cfg), _view_ptrs(views), B, P.data_ptr(),
                                                C.byref(chain.struct), q.data_ptr(), _ptr(Rv), cams_t.data_ptr(),
                                                C.byref(o), _stream(dev)), "mvgeo_pipeline_views")
    out["_keepalive"] = (cams_t, Rv, views)
    return out


class HostPipeline:
    """Host-buffer front end (mvgeo_ctx): pinned host tensors in, pinned host tensors out, H2D of
    frame chunk i+1 overlapped with the kernels of chunk i. This is the call a CPU-tensor
    caller makes (the reference decodes after `.cpu()`, DIP_REAL.py:113)."""

    def __init__(self, chain: Chain, rig: CameraRig, R_view, *, dtype, H: int, W: int, image_size=None,
                 soft: Optional[str] = "global", beta: float = 100.0, window_radius: int = 3,
                 apply_sigmoid: bool = False, tri_use_soft: bool = True, tri_weighted: bool = False,
                 min_score: float = 0.0, lam: float = 1.0, chunk_frames: int = 64, device: int = 0):
        lib = _lib.load()
        self.chain, self.V, self.K, self.H, self.W, self.dtype = chain, rig.n_views, chain.n_points, H, W, dtype
        self.cfg = _make_cfg(dtype, H, W, self.V, self.K, image_size, soft, beta, window_radius, apply_sigmoid,
                             tri_use_soft, tri_weighted, min_score, lam)
        self._cams = np.ascontiguousarray(rig.packed())
        Rv = None if R_view is None else np.ascontiguousarray(np.asarray(R_view, dtype=np.float32))
        self._Rv = Rv
        self._P = np.ascontiguousarray(rig.projection_matrices(None if Rv is None else Rv.astype(np.float64)))
        self._ctx = C.c_void_p()
        _lib.check(lib.mvgeo_ctx_create(C.byref(self._ctx), int(device), C.byref(self.cfg), C.byref(chain.struct),
                                        int(chunk_frames)), "mvgeo_ctx_create")

    def run(self, maps_host: torch.Tensor, q_host: torch.Tensor, out: Optional[dict] = None) -> dict:
        if maps_host.is_cuda or q_host.is_cuda:
            raise ValueError("HostPipeline takes host tensors; use ops.pipeline for device tensors")
        B = int(maps_host.shape[0])
        if tuple(maps_host.shape) != (B, self.V, self.K, self.H, self.W) or maps_host.dtype != self.dtype:
            raise ValueError("maps_host shape / dtype does not match the context")
        if tuple(q_host.shape) != (B, self.chain.n_joints) or q_host.dtype != torch.float32:
            raise ValueError("q_host must be float32 (B, J)")
        maps_host, q_host = maps_host.contiguous(), q_host.contiguous()
        if out is None:
            out = alloc_outputs(B, self.V, self.K, None, pin=True)
        o = _out_struct(out)
        _lib.check(_lib.load().mvgeo_pipeline_host(
            self._ctx, maps_host.data_ptr(), B, self._P.ctypes.data, q_host.data_ptr(),
            None if self._Rv is None else self._Rv.ctypes.data, self._cams.ctypes.data, C.byref(o)),
            "mvgeo_pipeline_host")
        return out

    def close(self):
        if self._ctx:
            _lib.load().mvgeo_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
