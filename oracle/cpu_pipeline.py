"""
oracle/cpu_pipeline.py — the reference's CPU execution of the hot path, restated call for call.

TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs).
It runs the path the way the reference runs it: Python loops over frames and views, one
torch-CPU decode per (frame, view) (extract_keypoints_from_heatmaps,
model/Fr5_model_train.ipynb:4674-4705), one NumPy float64 FK and one projection per
(frame, view) (model/MvRoPose_FR3.py:90-141), plus — because the reference has no
triangulation — the float64 SVD-DLT restatement per (frame, key-point), labelled
"restatement". Frames are spread over worker processes the way the reference spreads
samples over DataLoader workers (num_workers=8, model/MvRoPose_FR3.py:969).
"""
from __future__ import annotations

import math
import os
import time

import numpy as np

from . import mvgeo_oracle as O

_ROBOTS = {
    "fr3": (O.fk_fr3, ["view1", "view2", "view3", "view4"], 7, (-2.0, 2.0)),
    "fr5": (O.fk_fr5, ["top", "left", "right"], 6, (-120.0, 120.0)),
    "meca500": (lambda q, view=None: O.fk_meca500(q), [None], 6, (-120.0, 120.0)),
}


def ring_rig(V: int, K_list, radius=1.5, height=0.8, target=(0.0, 0.0, 0.4), phase=0.3):
    """Same synthetic rig as the product's CameraRig.synthetic_ring (zero distortion)."""
    cams = []
    tgt = np.asarray(target, dtype=np.float64)
    for v in range(V):
        ang = 2.0 * math.pi * v / V + phase
        c = np.array([radius * math.cos(ang), radius * math.sin(ang), height])
        z = tgt - c
        z /= np.linalg.norm(z)
        x = np.cross(z, [0.0, 0.0, 1.0])
        x /= np.linalg.norm(x)
        R = np.stack([x, np.cross(z, x), z])
        cams.append(dict(R=R, t=-R @ c, K=np.asarray(K_list[v % len(K_list)], dtype=np.float64)))
    return cams


def make_frames(robot: str, V: int, n_frames: int, H: int, W: int, image_size, K_list, seed: int):
    """Synthetic closed-loop sample: FK -> project -> Gaussian blobs (sigma 3 px) + N(0, 0.01)
    noise, as float32 CPU tensors [n, V, K, H, W] (the reference decodes float32 CPU maps)."""
    import torch

    fk, views, J, (lo, hi) = _ROBOTS[robot]
    rng = np.random.default_rng(seed)
    cams = ring_rig(V, K_list)
    q = rng.uniform(lo, hi, size=(n_frames, J))
    Hi, Wi = image_size
    ys, xs = np.arange(H, dtype=np.float32)[:, None], np.arange(W, dtype=np.float32)[None, :]
    Kp = J + 1
    maps = np.empty((n_frames, V, Kp, H, W), dtype=np.float32)
    for f in range(n_frames):
        for v in range(V):
            X = fk(q[f], views[v % len(views)])
            uv = O.project_points(X, cams[v]["R"], cams[v]["t"], cams[v]["K"])
            for k in range(Kp):
                cx, cy = uv[k, 0] * W / Wi, uv[k, 1] * H / Hi
                maps[f, v, k] = np.exp(-((xs - cx) ** 2) / 18.0) * np.exp(-((ys - cy) ** 2) / 18.0)
    maps += rng.normal(0.0, 0.01, size=maps.shape).astype(np.float32)
    return torch.from_numpy(maps), q, cams


def run_frames(robot: str, maps, q, cams, image_size, min_score: float = 0.5):
    """One pass of the path over the sample, exactly as the reference would loop it.
    Returns per-frame results so the work cannot be optimised away."""
    fk, views, J, _ = _ROBOTS[robot]
    n, V, Kp = maps.shape[:3]
    P = np.stack([O.projection_matrix(c["K"], c["R"] @ O.view_rotation(robot, views[v % len(views)]), c["t"])
                  for v, c in enumerate(cams)])
    out = []
    for f in range(n):
        kps = np.zeros((1, V, Kp, 2))
        scores = np.zeros((1, V, Kp))
        err = 0.0
        for v in range(V):
            kp, sc = O.extract_keypoints_from_heatmaps(maps[f, v], image_size)      # decode
            X = fk(q[f], views[v % len(views)])                                       # FK
            uv = O.project_points(X, cams[v]["R"], cams[v]["t"], cams[v]["K"])        # reprojection
            err += float(np.mean((uv - kp) ** 2))
            kps[0, v], scores[0, v] = kp, sc
        Xt, _, _ = O.triangulate_dlt(kps, P, scores, min_weight=min_score)            # DLT (restatement)
        out.append((Xt[0], err / V))
    return out


def _worker(args):
    robot, V, n_frames, H, W, image_size, K_list, seed, min_seconds = args
    import torch

    torch.set_num_threads(1)
    maps, q, cams = make_frames(robot, V, n_frames, H, W, image_size, K_list, seed)
    run_frames(robot, maps[:1], q[:1], cams, image_size)  # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        run_frames(robot, maps, q, cams, image_size)
        done += n_frames
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            return done, dt


def timed_throughput(robot: str, V: int, H: int, W: int, image_size, K_list, frames_per_worker: int = 4,
                     min_seconds: float = 8.0, workers: int | None = None, timeout: float = 600.0):
    """frames/s of the CPU path with `workers` processes (default: every host core), each
    looping over its own `frames_per_worker` synthetic frames for at least `min_seconds`.
    Workers are plain subprocesses (`python -m oracle.cpu_pipeline ...`): no fork after CUDA
    initialisation, no dependence on the parent's __main__, hard timeout.
    Returns (frames_per_second, workers, total_frames, description)."""
    import json
    import subprocess
    import sys

    workers = workers or (os.cpu_count() or 1)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = dict(robot=robot, V=V, n=frames_per_worker, H=H, W=W, image_size=list(image_size),
                K_list=[np.asarray(k).tolist() for k in K_list], min_seconds=min_seconds)
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, "-m", "oracle.cpu_pipeline", json.dumps(dict(spec, seed=1234 + w))],
                              cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for w in range(workers)]
    res = []
    deadline = time.time() + timeout
    for p in procs:
        try:
            out, err = p.communicate(timeout=max(1.0, deadline - time.time()))
        except subprocess.TimeoutExpired:
            for q_ in procs:
                q_.kill()
            raise RuntimeError("cpu baseline worker timed out")
        if p.returncode != 0:
            raise RuntimeError("cpu baseline worker failed: " + err[-2000:])
        r = json.loads(out.strip().splitlines()[-1])
        res.append((r["frames"], r["seconds"]))
    fps = sum(d / t for d, t in res)  # workers run concurrently on their own cores
    total = sum(d for d, _ in res)
    desc = (f"{workers} processes x {frames_per_worker} synthetic frames (V={V}, {H}x{W} float32 maps), looped "
            f">= {min_seconds:.0f} s each; decode+FK+projection as the reference loops them, DLT = float64 SVD restatement")
    return fps, workers, total, desc


if __name__ == "__main__":
    import json
    import sys

    a = json.loads(sys.argv[1])
    frames, seconds = _worker((a["robot"], a["V"], a["n"], a["H"], a["W"], tuple(a["image_size"]), a["K_list"], a["seed"],
                               a["min_seconds"]))
    print(json.dumps(dict(frames=frames, seconds=seconds)))
