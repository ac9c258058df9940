"""Serial-chain descriptions (host side). The numeric tables live in the C library
(mvgeo_chain_builtin, csrc/fk.cu) — this module only wraps them and builds custom chains."""
from __future__ import annotations

import math
from typing import Sequence

from . import _lib

_ROBOT_IDS = {"fr3": _lib.ROBOT_FR3, "fr5": _lib.ROBOT_FR5, "meca500": _lib.ROBOT_MECA500}

#: scipy `R.from_euler('zyx', angles, degrees=True)` base corrections per view:
#: FR3 model/MvRoPose_FR3.py:105-110, Fr5 model/Fr5_model_train.ipynb:269-273, Meca500 none.
VIEW_EULER_ZYX_DEG = {
    "fr3": {"view1": (90, 180, 0), "view2": (90, 180, 0), "view3": (90, 180, 0), "view4": (90, 180, 0)},
    "fr5": {"top": (-85, 0, 180), "left": (180, 0, 90), "right": (0, 0, 90)},
    "meca500": {},
}


class Chain:
    """A DH chain handed to the kernels by value (mvgeo_chain, include/mvgeo.h)."""

    def __init__(self, struct: _lib.ChainStruct, name: str = "custom"):
        self.struct = struct
        self.name = name

    @property
    def n_joints(self) -> int:
        return int(self.struct.n_joints)

    @property
    def n_points(self) -> int:
        return self.n_joints + (1 if self.struct.emit_base else 0)

    @staticmethod
    def builtin(name: str) -> "Chain":
        s = _lib.ChainStruct()
        _lib.check(_lib.load().mvgeo_chain_builtin(_ROBOT_IDS[name.lower()], s), "mvgeo_chain_builtin")
        return Chain(s, name.lower())

    @staticmethod
    def from_dh(a: Sequence[float], d: Sequence[float], alpha_rad: Sequence[float], theta_offset: Sequence[float],
                convention: str = "standard", angle_scale: float = 1.0, emit_base: bool = True) -> "Chain":
        """Custom chain. `ForwardKinematics(dh_params)` of model/MV-model.ipynb:841-874 is
        convention='standard', angle_scale=1, emit_base=False with tuples (theta0, d, a, alpha)."""
        n = len(a)
        if not (1 <= n <= _lib.MAX_JOINTS and len(d) == n and len(alpha_rad) == n and len(theta_offset) == n):
            raise ValueError(f"chain must have 1..{_lib.MAX_JOINTS} joints with matching table lengths")
        s = _lib.ChainStruct()
        s.n_joints = n
        s.convention = {"standard": _lib.DH_STANDARD, "modified": _lib.DH_MODIFIED}[convention]
        s.emit_base = 1 if emit_base else 0
        s.angle_scale = angle_scale
        for i in range(n):
            s.a[i], s.d[i] = a[i], d[i]
            s.cos_alpha[i], s.sin_alpha[i] = math.cos(alpha_rad[i]), math.sin(alpha_rad[i])
            s.theta_offset[i] = theta_offset[i]
        return Chain(s)


def euler_zyx_extrinsic(angles_deg):
    """3x3 (nested lists, float64) of scipy Rotation.from_euler('zyx', [a,b,c], degrees=True):
    R = Rx(c) Ry(b) Rz(a)."""
    a, b, c = (math.radians(float(v)) for v in angles_deg)
    ca, sa, cb, sb, cc, sc = math.cos(a), math.sin(a), math.cos(b), math.sin(b), math.cos(c), math.sin(c)
    Rz = [[ca, -sa, 0.0], [sa, ca, 0.0], [0.0, 0.0, 1.0]]
    Ry = [[cb, 0.0, sb], [0.0, 1.0, 0.0], [-sb, 0.0, cb]]
    Rx = [[1.0, 0.0, 0.0], [0.0, cc, -sc], [0.0, sc, cc]]
    mm = lambda A, Bm: [[sum(A[i][k] * Bm[k][j] for k in range(3)) for j in range(3)] for i in range(3)]
    return mm(Rx, mm(Ry, Rz))


def view_rotation(robot: str, view) -> list:
    """Base correction applied by the reference's angle_to_joint_coordinate; unknown view ->
    identity (`if selected_view in view_rotations`, model/MvRoPose_FR3.py:113-114)."""
    table = VIEW_EULER_ZYX_DEG[robot.lower()]
    if view in table:
        return euler_zyx_extrinsic(table[view])
    return [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
