"""ctypes binding of libmvgeo.so (include/mvgeo.h). No torch types cross this boundary: only
raw pointers, sizes and a cudaStream_t. There is NO fallback: if the library is missing the
import of any compute entry point raises, loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmvgeo.so")

MAX_JOINTS = 8
MAX_VIEWS = 16
MAX_WINDOW_RADIUS = 15

F32, BF16, F16 = 0, 1, 2
SOFT_NONE, SOFT_GLOBAL, SOFT_WINDOW = 0, 1, 2
ROBOT_FR3, ROBOT_FR5, ROBOT_MECA500 = 0, 1, 2
DH_STANDARD, DH_MODIFIED = 0, 1


class MvgeoError(RuntimeError):
    pass


class ChainStruct(C.Structure):
    _fields_ = [
        ("n_joints", C.c_int32),
        ("convention", C.c_int32),
        ("emit_base", C.c_int32),
        ("angle_scale", C.c_float),
        ("a", C.c_float * MAX_JOINTS),
        ("d", C.c_float * MAX_JOINTS),
        ("cos_alpha", C.c_float * MAX_JOINTS),
        ("sin_alpha", C.c_float * MAX_JOINTS),
        ("theta_offset", C.c_float * MAX_JOINTS),
    ]


class PipelineCfg(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("V", C.c_int32), ("K", C.c_int32),
        ("soft_mode", C.c_int32), ("window_radius", C.c_int32), ("apply_sigmoid", C.c_int32),
        ("tri_use_soft", C.c_int32), ("tri_weighted", C.c_int32),
        ("beta", C.c_float), ("min_score", C.c_float), ("lam", C.c_float),
        ("scale_x", C.c_double), ("scale_y", C.c_double),
    ]


class PipelineOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "idx", "peak", "score", "kp_hard", "kp_soft", "X_tri", "tri_resid", "tri_views", "X_fk", "uv_fk",
        "frame_loss", "loss", "ticket")]


_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

_SIGNATURES = {
    "mvgeo_version": ([], C.c_int),
    "mvgeo_error_string": ([_i], C.c_char_p),
    "mvgeo_chain_builtin": ([_i, C.POINTER(ChainStruct)], _i),
    "mvgeo_decode": ([_vp, _i, _i64, _i, _i, _d, _d, _i, _f, _i, _i, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mvgeo_decode_views": ([C.POINTER(_vp), _i, _i, _i64, _i, _i, _i, _d, _d, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mvgeo_triangulate": ([_vp, _vp, _vp, _i64, _i, _i, _f, _i, _vp, _vp, _vp, _vp], _i),
    "mvgeo_quat_mean": ([_vp, _vp, _i64, _i, _vp, _vp], _i),
    "mvgeo_fk": ([C.POINTER(ChainStruct), _vp, _i64, _vp, _i, _vp, _vp], _i),
    "mvgeo_project": ([_vp, _i, _vp, _i64, _i, _i, _vp, _vp], _i),
    "mvgeo_undistort_points": ([_vp, _vp, _i64, _i, _i, _i, _vp, _vp], _i),
    "mvgeo_fk_reproj_fwd": ([C.POINTER(ChainStruct), _vp, _i64, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp], _i),
    "mvgeo_fk_reproj_bwd": ([C.POINTER(ChainStruct), _vp, _i64, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp], _i),
    "mvgeo_geometry": ([_vp, _vp, _vp, C.POINTER(ChainStruct), _vp, _i64, _vp, _vp, _i, _i, _f, _i, _f, _vp, _vp, _vp, _vp,
                        _vp, _vp, _vp, _vp, _vp], _i),
    "mvgeo_pnp_refine": ([_vp, _i, _vp, _vp, _vp, _i64, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "mvgeo_pnp_solve": ([_vp, _i, _vp, _vp, _vp, _i64, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mvgeo_encode_gaussian": ([_vp, _i64, _i, _i, _f, _i, _vp, _vp], _i),
    "mvgeo_heatmap_mse": ([_vp, _i, _vp, _i64, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp], _i),
    "mvgeo_decode_mse": ([_vp, _i, _i64, _i, _i, _d, _d, _i, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mvgeo_pipeline": ([C.POINTER(PipelineCfg), _vp, _i64, _vp, C.POINTER(ChainStruct), _vp, _vp, _vp,
                        C.POINTER(PipelineOut), _vp], _i),
    "mvgeo_pipeline_views": ([C.POINTER(PipelineCfg), C.POINTER(_vp), _i64, _vp, C.POINTER(ChainStruct), _vp, _vp, _vp,
                              C.POINTER(PipelineOut), _vp], _i),
    "mvgeo_ctx_create": ([C.POINTER(_vp), _i, C.POINTER(PipelineCfg), C.POINTER(ChainStruct), _i64], _i),
    "mvgeo_ctx_destroy": ([_vp], _i),
    "mvgeo_pipeline_host": ([_vp, _vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(PipelineOut)], _i),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load libmvgeo.so (built in-tree by `make` / __graft_entry__.build()). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MvgeoError(
            f"{LIB_PATH} not found: build it with `make` (nvcc, sm_100a). This package has no CPU or "
            "PyTorch fallback by design.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = load().mvgeo_error_string(rc).decode()
    if rc < 0:
        raise ValueError(f"{what}: {msg} (mvgeo status {rc})")
    raise MvgeoError(f"{what}: CUDA error {rc}: {msg}")
