"""Per-call latency of the reference-named shims (mvgeo.compat) against the oracle port on ONE CPU core
(the reference's own single-frame call pattern): INTEGRATION.md section 5. Run on a GPU box."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.set_num_threads(1)
import mvgeo
from mvgeo import compat
from oracle import mvgeo_oracle as O

def bench(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

rng = np.random.default_rng(0)
hm_cpu = torch.from_numpy(rng.normal(size=(8, 128, 128)).astype(np.float32))
hm_gpu = hm_cpu.cuda()
q7 = [0.648, -0.108, 0.21, -1.92, 0.886, 3.10, -2.39]
K = np.array([[737.118, 0, 974.584], [0, 737.085, 552.68], [0, 0, 1.0]])
dist = np.array([-0.0056, -0.0461, 1.3e-4, 3.1e-4, 0.0148])
aruco = dict(rvec_x=0.1, rvec_y=-0.2, rvec_z=0.3, tvec_x=0.1, tvec_y=0.0, tvec_z=1.5)
X8 = compat.fr3.angle_to_joint_coordinate(q7, "view1")
Rm, tv = O.rodrigues([0.1, -0.2, 0.3]), np.array([0.1, 0.0, 1.5])
dh = [(0.0, 0.3, 0.1, 0.5)] * 6
fkc = compat.ForwardKinematics(dh)
ang8 = torch.from_numpy(rng.uniform(-1, 1, (8, 6)).astype(np.float32))
X83 = rng.uniform(-0.3, 0.3, (8, 6, 3)).astype(np.float32) + np.float32([0, 0, 2.0])
uv = O.project_points(X8, Rm, tv, K, dist)
hm_pose = (mvgeo.encode_gaussian(torch.tensor(uv * [128 / 1920, 128 / 1200], dtype=torch.float32, device="cuda"), (128, 128), 2.0) * 6 - 3).cpu()
import cv2
rows = []
def row(name, shim, ref, note=""):
    a, b = bench(shim), bench(ref, n=50, warm=5)
    rows.append({"call": name, "shim_us": round(a, 1), "reference_port_us_1core": round(b, 1), "note": note})
    print(rows[-1], flush=True)
row("extract_keypoints_from_heatmaps((8,128,128) CPU tensor)", lambda: compat.extract_keypoints_from_heatmaps(hm_cpu, (1200, 1920)),
    lambda: O.extract_keypoints_from_heatmaps(hm_cpu, (1200, 1920)), "shim = H2D 512 KB + 1 launch + D2H")
row("extract_keypoints_from_heatmaps((8,128,128) CUDA tensor)", lambda: compat.extract_keypoints_from_heatmaps(hm_gpu, (1200, 1920)),
    lambda: O.extract_keypoints_from_heatmaps(hm_cpu, (1200, 1920)), "maps already on the GPU (every reference call site)")
row("decode_argmax((8,128,128) CPU tensor)", lambda: compat.decode_argmax(hm_cpu, (1200, 1920)),
    lambda: O.decode_inline_argmax(hm_cpu, (1200, 1920)))
row("fr3.angle_to_joint_coordinate(7 angles)", lambda: compat.fr3.angle_to_joint_coordinate(q7, "view1"), lambda: O.fk_fr3(q7, "view1"))
row("fr3.joint_coordinate_to_pixel_plane(8 points)", lambda: compat.fr3.joint_coordinate_to_pixel_plane(X8, aruco, K, dist),
    lambda: O.project_points(X8, Rm, tv, K, dist))
row("ForwardKinematics.forward((8,6))", lambda: fkc.forward(ang8), lambda: O.fk_generic(dh, ang8.numpy()))
row("project_3d_to_2d((8,6,3))", lambda: compat.project_3d_to_2d(X83, K, dist, [np.zeros(3)] * 8, [np.zeros(3)] * 8),
    lambda: [O.project_points(X83[b], np.eye(3), np.zeros(3), K, dist) for b in range(8)], "one launch for the batch")
row("create_gt_heatmap((x,y),(128,128),5)", lambda: compat.create_gt_heatmap((40.3, 77.8), (128, 128), 5.0),
    lambda: O.create_gt_heatmap((40.3, 77.8), (128, 128), 5.0))
row("fr3.estimate_camera_pose (FK + decode + PnP)", lambda: compat.fr3.estimate_camera_pose(torch.tensor(q7), hm_pose, K, dist, "view1", (1200, 1920), 0.5),
    lambda: cv2.solvePnPRansac(X8.astype(np.float64), uv, K, dist, flags=cv2.SOLVEPNP_EPNP), "reference column = cv2.solvePnPRansac alone")
# batched API for scale: 4096 frames x 4 views in one call
chain = mvgeo.Chain.builtin("fr3")
qB = torch.from_numpy(rng.uniform(-2, 2, (4096, 7)).astype(np.float32)).cuda()
Rv = np.stack([np.asarray(mvgeo.view_rotation("fr3", "view1"))] * 4).astype(np.float32)
us = bench(lambda: mvgeo.forward_kinematics(chain, qB, Rv), n=100)
rows.append({"call": "ops.forward_kinematics(4096 frames x 4 views), per frame-view", "shim_us": round(us / (4096 * 4), 4),
             "reference_port_us_1core": rows[3]["reference_port_us_1core"], "note": f"one launch {us:.1f} us"})
print(rows[-1])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/shim_latency.json", "w"), indent=1)
