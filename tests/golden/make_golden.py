"""
tests/golden/make_golden.py — freeze golden vectors from the UNMODIFIED reference.

Run once in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
It imports the reference's own functions through oracle/ref_loader.py, evaluates them on
seeded inputs and writes tests/golden/reference_golden.npz. Inputs are regenerated in the
tests from the same seeds (numpy PCG64 streams are stable across platforms); a SHA-256 of
every generated input is stored so RNG drift is detected instead of silently mis-compared.
"""
from __future__ import annotations

import glob
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

SEED = 20251018


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_inputs():
    """Deterministic inputs shared with tests/test_oracle_golden.py."""
    rng = np.random.default_rng(SEED)
    inp = {}
    inp["fr3_q"] = rng.uniform(-2.8, 2.8, size=(48, 7))
    inp["fr3_q"][0] = [0.6480, -0.1083, 0.2098, -1.9201, 0.8858, 3.1025, -2.3876]  # SURVEY 8c pose
    inp["fr3_q"][1] = 0.0
    inp["fr5_q_rand"] = rng.uniform(-175.0, 175.0, size=(24, 6))
    inp["fr5_q_rand"][0] = 0.0
    inp["meca_q"] = rng.uniform(-170.0, 170.0, size=(32, 6))
    inp["meca_q"][0] = 0.0
    inp["meca_q"][1] = [10, -20, 30, -40, 50, -60]
    inp["generic_angles"] = rng.uniform(-3.0, 3.0, size=(6, 5)).astype(np.float32)
    inp["generic_dh"] = np.array(  # (theta0, d, a, alpha) radians, arbitrary 5-joint arm
        [[0.1, 0.30, 0.00, 1.5707963], [-0.4, 0.00, 0.25, 0.0], [0.0, 0.05, 0.20, -1.5707963],
         [0.7, 0.18, 0.00, 1.5707963], [0.0, 0.00, 0.06, 0.3]])
    inp["proj_rvec"] = rng.uniform(-0.6, 0.6, size=(6, 3))
    inp["proj_rvec"][0] = [0.1, -0.2, 0.3]
    inp["proj_tvec"] = np.stack([rng.uniform(-0.3, 0.3, 6), rng.uniform(-0.3, 0.3, 6), rng.uniform(1.2, 2.5, 6)], axis=1)
    inp["proj_tvec"][0] = [0.1, 0.0, 1.5]
    inp["maps_small"] = rng.normal(0.0, 1.0, size=(5, 7, 24, 40)).astype(np.float32)
    inp["maps_native"] = rng.uniform(0.0, 1.0, size=(7, 128, 128)).astype(np.float32)
    # coarse values: many exact ties, exercises the first-maximum rule
    inp["maps_ties"] = (rng.integers(0, 4, size=(3, 8, 30, 40)) * 0.25).astype(np.float32)
    inp["gt_kp"] = np.array([[40.3, 77.8], [0.0, 0.0], [127.0, 127.0], [-3.5, 60.2], [63.5, 63.5]])
    inp["loss_pred"] = rng.normal(500.0, 100.0, size=(4, 7, 2)).astype(np.float32)
    inp["loss_gt"] = (inp["loss_pred"] + rng.normal(0.0, 5.0, size=(4, 7, 2))).astype(np.float32)
    # camera-pose case (Fr5, view "top"): joint angles in degrees and the true camera pose; no random draws
    inp["pose_q"] = np.array([10.0, -80.0, 100.0, -110.0, -85.0, 20.0])
    inp["pose_rt"] = np.array([[1.9, 0.3, -0.2], [0.1, -0.05, 1.8]])
    return inp


POSE_MAP_HW, POSE_IMAGE_HW, POSE_THRESHOLD = (270, 480), (1200, 1920), 0.5


def pose_case_maps(uv, low=()):
    """Belief maps (7, 270, 480) float32 in logit scale for the camera-pose case: a sigma = 2 px blob of amplitude 6 on
    a -3 floor at every projected key-point uv (image px); key-points in `low` get no blob (score sigmoid(-3) = 0.047,
    below the 0.5 threshold). Shared by the generator and the tests."""
    H, W = POSE_MAP_HW
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.empty((len(uv), H, W), dtype=np.float32)
    for k, (u, v) in enumerate(uv):
        cx, cy = u * W / POSE_IMAGE_HW[1], v * H / POSE_IMAGE_HW[0]
        amp = 0.0 if k in low else 6.0
        m[k] = amp * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 2.0 ** 2)) - 3.0
    return m


def main():
    assert ref_loader.available(), "reference checkout not found"
    import torch

    inp = gen_inputs()
    out = {f"sha_{k}": np.array(sha(v)) for k, v in inp.items()}

    fr3 = ref_loader.load_fr3()
    fr5 = ref_loader.load_fr5()
    meca = ref_loader.load_meca500()
    mv = ref_loader.load_mv_model()
    calib = ref_loader.load_calib()

    # ---- camera intrinsics from the reference's own .conf parser (ZED-X, FHD1200) ----
    serials = ["41182735", "49429257", "44377151", "49045152"]
    Ks, dists = [], []
    for side in ("left", "right"):
        for sn in serials:
            K, dist, _ = calib["load_fhd_calibration"](
                os.path.join(ref_loader.REF_ROOT, "dataset", "All_camera_conf", f"SN{sn}.conf"), side)
            Ks.append(K)
            dists.append(dist)
    out["zedx_K"] = np.array(Ks, dtype=np.float64)        # (8,3,3): 4 left then 4 right
    out["zedx_dist"] = np.array(dists, dtype=np.float64)  # (8,5)

    # ---- FK ----
    out["fr3_fk_view1"] = np.stack([fr3["angle_to_joint_coordinate"](q, "view1") for q in inp["fr3_q"]])
    out["fr3_fk_noview"] = np.stack([fr3["angle_to_joint_coordinate"](q, "none") for q in inp["fr3_q"]])
    # real Fr5 joint rows (degrees) from the repo's data fixture
    import pandas as pd

    csvs = sorted(glob.glob(os.path.join(ref_loader.REF_ROOT, "dataset", "Fr5", "*", "matched_index.csv")))
    df = pd.read_csv(csvs[0])
    cols = [f"joint.{i}" for i in range(6)]
    real = df[cols].to_numpy(dtype=np.float64)[:: max(1, len(df) // 40)][:40]
    out["fr5_q_real"] = real
    fr5_q = np.concatenate([inp["fr5_q_rand"], real], axis=0)
    for view in ("top", "left", "right", "none"):
        out[f"fr5_fk_{view}"] = np.stack([fr5["angle_to_joint_coordinate"](q, view) for q in fr5_q])
    out["meca_fk"] = np.stack([meca["forward_kinematics"](q) for q in inp["meca_q"]])
    fk = mv["ForwardKinematics"]([tuple(r) for r in inp["generic_dh"]])
    out["generic_fk"] = fk.forward(torch.from_numpy(inp["generic_angles"])).numpy()

    # ---- projection (cv2.projectPoints through the reference wrappers) ----
    K0 = np.array(Ks[0], dtype=np.float32)
    d_real = np.array(dists[0], dtype=np.float32)
    d_zero = np.zeros(5, dtype=np.float32)
    uv_zero, uv_real, uv_fr5 = [], [], []
    for i in range(6):
        X = out["fr3_fk_view1"][i]
        ar = dict(rvec_x=inp["proj_rvec"][i, 0], rvec_y=inp["proj_rvec"][i, 1], rvec_z=inp["proj_rvec"][i, 2],
                  tvec_x=inp["proj_tvec"][i, 0], tvec_y=inp["proj_tvec"][i, 1], tvec_z=inp["proj_tvec"][i, 2])
        uv_zero.append(fr3["joint_coordinate_to_pixel_plane"](X, ar, K0, d_zero))
        uv_real.append(fr3["joint_coordinate_to_pixel_plane"](X, ar, K0, d_real))
        ar_deg = dict(ar)
        for ax in "xyz":  # the Fr5 wrapper takes rvec in DEGREES (Fr5_model_train.ipynb:291-295)
            ar_deg[f"rvec_{ax}"] = float(np.degrees(ar[f"rvec_{ax}"]))
        uv_fr5.append(fr5["joint_coordinate_to_pixel_plane"](out["fr5_fk_top"][i], ar_deg, K0, d_real))
    out["proj_fr3_zero"] = np.stack(uv_zero)
    out["proj_fr3_real"] = np.stack(uv_real)
    out["proj_fr5_real_degrvec"] = np.stack(uv_fr5)
    uv_meca = meca["project_to_pixel"](out["meca_fk"][1], np.deg2rad(np.array([96, 98, -45], dtype=np.float32)),
                                       np.array([0, -0.01, 0.75], dtype=np.float32), K0, d_real)
    out["proj_meca_prior"] = uv_meca
    j3 = torch.from_numpy(out["generic_fk"])
    out["proj_generic"] = mv["project_3d_to_2d"](
        j3, K0.astype(np.float64), None,
        rvec=[inp["proj_rvec"][i].reshape(3, 1) for i in range(6)],
        tvec=[inp["proj_tvec"][i].reshape(3, 1) for i in range(6)]).numpy()

    # ---- decoder ----
    kp, sc = [], []
    for f in range(inp["maps_small"].shape[0]):
        k_, s_ = fr5["extract_keypoints_from_heatmaps"](torch.from_numpy(inp["maps_small"][f]), (1200, 1920))
        kp.append(k_)
        sc.append(s_)
    out["dec_small_kp"], out["dec_small_score"] = np.stack(kp), np.stack(sc)
    k_, s_ = fr5["extract_keypoints_from_heatmaps"](torch.from_numpy(inp["maps_native"]), (1080, 1920))
    out["dec_native_kp"], out["dec_native_score"] = k_, s_
    # plain arg-max loop exactly as DIP_REAL.py:116-124 / MvRoPose_FR3.py:299-304 (re-typed: it is
    # inline code, not a function, so it cannot be imported)
    tie_idx = []
    for f in range(inp["maps_ties"].shape[0]):
        t = torch.from_numpy(inp["maps_ties"][f])
        tie_idx.append([int(torch.argmax(t[j])) for j in range(t.shape[0])])
    out["dec_ties_idx"] = np.array(tie_idx, dtype=np.int64)
    out["dec_small_rawidx"] = np.array(
        [[int(torch.argmax(torch.from_numpy(inp["maps_small"][f, j]))) for j in range(7)] for f in range(5)], dtype=np.int64)

    # ---- GT belief maps ----
    out["gt_maps_128"] = np.stack([fr3["create_gt_heatmap"](tuple(k), (128, 128), 5.0) for k in inp["gt_kp"]])
    out["gt_maps_rect"] = np.stack([fr3["create_gt_heatmap"]((k[0] * 0.3, k[1] * 0.2), (24, 40), 3.0) for k in inp["gt_kp"]])

    # ---- loss ----
    pred = dict(keypoints_2d=torch.from_numpy(inp["loss_pred"]), angles=torch.zeros(4, 7),
                proj_2d=torch.from_numpy(inp["loss_pred"]) + 1.5)
    out["loss_fk_only"] = np.array(float(mv["robot_pose_loss"](
        dict(pred, keypoints_2d=torch.from_numpy(inp["loss_gt"])), gt_keypoints=torch.from_numpy(inp["loss_gt"]),
        lambda_kp=0.0, lambda_fk=2.5)))

    # ---- camera pose without a prior: estimate_camera_pose (FK -> decode -> confidence filter -> cv2.solvePnPRansac EPNP,
    #      model/Fr5_model_train.ipynb:4707-4753) on belief maps rendered at the projections of the true pose ----
    import contextlib
    import io

    import cv2

    obj = fr5["angle_to_joint_coordinate"](inp["pose_q"], "top")
    Kp, dp = out["zedx_K"][0], out["zedx_dist"][0]
    uv, _ = cv2.projectPoints(obj.astype(np.float64), inp["pose_rt"][0], inp["pose_rt"][1], Kp, dp)
    uv = uv.reshape(-1, 2)
    out["pose_uv_true"] = uv
    for name, low in (("ok", (6,)), ("refused", (1, 3, 5, 6))):
        cv2.setRNGSeed(7)
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints the scores
            rvec, tvec, obj3, img2 = fr5["estimate_camera_pose"](
                torch.tensor(inp["pose_q"]), torch.from_numpy(pose_case_maps(uv, low)), Kp, dp, "top",
                POSE_IMAGE_HW, POSE_THRESHOLD)
        out[f"pose_{name}_obj"], out[f"pose_{name}_img"] = obj3, img2
        out[f"pose_{name}_rt"] = (np.full((2, 3), np.nan) if rvec is None else
                                  np.stack([np.asarray(rvec).ravel(), np.asarray(tvec).ravel()]))

    import scipy

    out["versions"] = np.array(f"cv2 {cv2.__version__}; scipy {scipy.__version__}; numpy {np.__version__}; torch {torch.__version__}")
    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
