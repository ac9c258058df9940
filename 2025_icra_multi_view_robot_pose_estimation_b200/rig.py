"""Calibrated camera rigs (host side): intrinsics, extrinsics, projection matrices.

Calibration front-end of the path: ZED `.conf` intrinsics
(dataset/3_Calib_cam_save.py:17-50, dataset/4_Calib_cam_save.py:35-59), ArUco extrinsic
records {view, cam, tvec_x..z, rvec_x..z} (dataset/Franka_research3_preprocessing.py:285-291),
and P = K [R|t] for triangulation. This is set-up code that runs once per rig in float64 numpy,
exactly where the reference does it (host, offline); the per-frame work is in csrc/.
"""
from __future__ import annotations

import configparser
import math
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

#: (fx, fy, cx, cy, k1, k2, p1, p2, k3): section {LEFT,RIGHT}_CAM_FHD1200 (1920x1200) of
#: dataset/All_camera_conf/SN{serial}.conf, ZED-X; identical to what the reference's
#: load_fhd_calibration returns (pinned in tests/golden/reference_golden.npz: zedx_K, zedx_dist).
ZEDX_FHD1200 = {
    "41182735_left": (737.118, 737.085, 974.584, 552.68, -0.005643106304680097, -0.04613633865985787, 0.00013427180750958065, 0.000311206091784389, 0.014788022640489918),
    "49429257_left": (740.294, 740.666, 970.826, 555.388, -0.004324792836863234, -0.044924273051576244, -0.00015804802069500418, 8.446305090471108e-05, 0.014128446052889002),
    "44377151_left": (742.696, 742.894, 973.838, 535.853, -0.0017388454017429765, -0.05135136385205364, -4.901526708441068e-05, 7.529832735358121e-05, 0.018002525767767445),
    "49045152_left": (739.738, 739.836, 967.386, 561.871, -0.006431271312612757, -0.040943310019251125, 0.00018106696256736828, 0.00010602951664724419, 0.011960269432213276),
    "41182735_right": (737.599, 737.411, 952.203, 519.56, -0.011644211970573087, -0.03446171529477812, -0.00025158205901685437, -1.601647386337792e-05, 0.009449363842105223),
    "49429257_right": (739.865, 739.852, 973.085, 544.875, -0.004010291097564468, -0.04379502094552287, -0.00021885868880579663, 0.000343442064376711, 0.013218838686316696),
    "44377151_right": (738.193, 738.258, 964.927, 542.498, -0.0005965345717048424, -0.050415878106735264, -0.0001392298076543885, -0.0001475244168605054, 0.01627959228038135),
    "49045152_right": (741.345, 741.305, 964.97, 557.332, -0.004702661745079126, -0.043820573013860074, -0.00013491302983219014, -9.248572284493359e-05, 0.013343884779934726),
}
ZEDX_IMAGE_SIZE = (1200, 1920)  # (H, W)
#: serial -> view name, model/MvRoPose_FR3.py:169-172
FR3_SERIAL_TO_VIEW = {"41182735": "view1", "49429257": "view2", "44377151": "view3", "49045152": "view4"}
#: model/Fr5_model_train.ipynb:344-348
FR5_SERIAL_TO_VIEW = {"38007749": "left", "34850673": "right", "30779426": "top"}


def load_conf_calibration(conf_path: str, side: str, section_suffix: str = "FHD1200"):
    """Parse a ZED `.conf` (BOM-prefixed for ZED-X) like load_fhd_calibration
    (dataset/4_Calib_cam_save.py:35-59; `section_suffix='FHD'` gives dataset/3_Calib_cam_save.py:17-50).
    Returns (camera_matrix 3x3 list, distortion [k1,k2,p1,p2,k3], advanced-distortion dict)."""
    config = configparser.ConfigParser()
    with open(conf_path, "r", encoding="utf-8-sig") as f:
        config.read_file(f)
    cam = config[f"{side.upper()}_CAM_{section_suffix}"]
    fx, fy, cx, cy = (float(cam[k]) for k in ("fx", "fy", "cx", "cy"))
    k1, k2, k3, p1, p2 = (float(cam[k]) for k in ("k1", "k2", "k3", "p1", "p2"))
    adv = {}
    adv_section = f"{side.upper()}_DISTO"
    if adv_section in config:
        adv = {k: float(v) for k, v in config[adv_section].items()}
    return [[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], [k1, k2, p1, p2, k3], adv


def rodrigues(rvec) -> np.ndarray:
    """cv2.Rodrigues(rvec)[0] in float64."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = float(np.linalg.norm(r))
    if th < 2.220446049250313e-16:
        return np.eye(3)
    k = r / th
    c, s = math.cos(th), math.sin(th)
    Kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return c * np.eye(3) + (1.0 - c) * np.outer(k, k) + s * Kx


@dataclass
class CameraRig:
    """V calibrated cameras. K (V,3,3), dist (V,5), R (V,3,3) world->camera, t (V,3); float64."""
    K: np.ndarray
    dist: np.ndarray
    R: np.ndarray
    t: np.ndarray
    image_size: tuple = ZEDX_IMAGE_SIZE

    @property
    def n_views(self) -> int:
        return int(self.K.shape[0])

    @staticmethod
    def from_aruco(records: Sequence[dict], Ks, dists, rvec_in_degrees: bool = False, image_size=ZEDX_IMAGE_SIZE):
        """records: ArUco pose dicts (rvec_x.., tvec_x..), one per view. rvec is radians for
        FR3 (model/Franka_research3_model_train.ipynb:275-279) and DEGREES for Fr5 / Meca500
        (model/Fr5_model_train.ipynb:291-295, visualization/Meca500_vis.ipynb:133-138)."""
        R, t = [], []
        for rec in records:
            rv = np.array([rec["rvec_x"], rec["rvec_y"], rec["rvec_z"]], dtype=np.float64)
            if rvec_in_degrees:
                rv = np.radians(rv)
            R.append(rodrigues(rv))
            t.append([rec["tvec_x"], rec["tvec_y"], rec["tvec_z"]])
        Ks = np.asarray(Ks, dtype=np.float64).reshape(-1, 3, 3)
        dists = np.zeros((len(R), 5)) if dists is None else np.asarray(dists, dtype=np.float64).reshape(-1, 5)
        return CameraRig(Ks, dists, np.array(R), np.array(t, dtype=np.float64), tuple(image_size))

    @staticmethod
    def synthetic_ring(n_views: int, radius: float = 1.5, height: float = 0.8, target=(0.0, 0.0, 0.4),
                       distortion: bool = False, phase: float = 0.3) -> "CameraRig":
        """The bench rig of SURVEY.md section 8d: V cameras on a circle looking at `target`, with
        the real ZED-X intrinsics taken in order (4 left, then 4 right, then cycling)."""
        names = list(ZEDX_FHD1200)
        K, D, Rs, ts = [], [], [], []
        tgt = np.asarray(target, dtype=np.float64)
        for v in range(n_views):
            fx, fy, cx, cy, k1, k2, p1, p2, k3 = ZEDX_FHD1200[names[v % len(names)]]
            K.append([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])
            D.append([k1, k2, p1, p2, k3] if distortion else [0.0] * 5)
            ang = 2.0 * math.pi * v / n_views + phase
            c = np.array([radius * math.cos(ang), radius * math.sin(ang), height])
            z = tgt - c
            z /= np.linalg.norm(z)
            x = np.cross(z, [0.0, 0.0, 1.0])
            x /= np.linalg.norm(x)
            y = np.cross(z, x)
            Rm = np.stack([x, y, z])
            Rs.append(Rm)
            ts.append(-Rm @ c)
        return CameraRig(np.array(K), np.array(D), np.array(Rs), np.array(ts))

    #: ring placement that keeps the whole arm in view, in the (view-rotated) frame the robot's FK
    #: emits: FR3's base correction flips z (model/MvRoPose_FR3.py:105-110), so its workspace is z < 0
    _RING_FOR = {"fr3": dict(radius=2.2, height=-1.2, target=(0.0, 0.0, -0.45)),
                 "fr5": dict(radius=2.4, height=0.6, target=(0.0, 0.0, 0.0)),
                 "meca500": dict(radius=0.9, height=0.5, target=(0.0, 0.0, 0.2))}

    @staticmethod
    def synthetic_ring_for(robot: str, n_views: int, distortion: bool = False) -> "CameraRig":
        """synthetic_ring aimed at `robot`'s workspace (closed-loop benchmarks and examples)."""
        return CameraRig.synthetic_ring(n_views, distortion=distortion, **CameraRig._RING_FOR[robot.lower()])

    def packed(self) -> np.ndarray:
        """(V,24) float32 rows laid out as mvgeo_camera (include/mvgeo.h)."""
        V = self.n_views
        out = np.zeros((V, 24), dtype=np.float32)
        out[:, 0:9] = self.R.reshape(V, 9)
        out[:, 9:12] = self.t
        out[:, 12], out[:, 13], out[:, 14], out[:, 15] = self.K[:, 0, 0], self.K[:, 1, 1], self.K[:, 0, 2], self.K[:, 1, 2]
        out[:, 16:21] = self.dist
        return out

    def projection_matrices(self, R_view: Optional[np.ndarray] = None) -> np.ndarray:
        """P_v = K_v [R_v R_view_v | t_v], (V,3,4) float32 (formed in float64). With R_view the
        triangulated points come out in the robot BASE frame although every view's FK output
        carries its own base rotation (SURVEY.md section 7.3 'World frame for triangulation')."""
        V = self.n_views
        P = np.zeros((V, 3, 4))
        for v in range(V):
            Rv = self.R[v] if R_view is None else self.R[v] @ np.asarray(R_view[v], dtype=np.float64)
            P[v] = self.K[v] @ np.hstack([Rv, self.t[v].reshape(3, 1)])
        return P.astype(np.float32)

    def subset(self, views: Iterable[int]) -> "CameraRig":
        idx = list(views)
        return CameraRig(self.K[idx], self.dist[idx], self.R[idx], self.t[idx], self.image_size)
