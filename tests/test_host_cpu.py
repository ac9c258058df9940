"""CPU tests of the host side: the C-ABI library loads and exports every symbol declared in
include/mvgeo.h, host-only entry points work, argument validation, frame sharding (incl. a
world_size-2 gloo gather), rig / calibration front-end, and the no-CPU-fallback rule."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import mvgeo_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="module")
def mv():
    import mvgeo

    if not os.path.isfile(mvgeo.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return mvgeo


def test_library_exports_every_declared_symbol(mv):
    header = open(os.path.join(ROOT, "include", "mvgeo.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(mvgeo_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 15
    lib = ctypes.CDLL(mv.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mvgeo.h but not exported"
    assert declared == set(mv._lib.EXPORTED_SYMBOLS)  # the ctypes binding covers the whole header
    assert mv._lib.load().mvgeo_version() == 100
    assert mv._lib.load().mvgeo_error_string(-1).decode().startswith("invalid argument")


def test_struct_layouts_match_header(mv):
    assert ctypes.sizeof(mv._lib.ChainStruct) == 16 + 5 * 8 * 4
    assert ctypes.sizeof(mv._lib.PipelineOut) == 13 * 8
    assert ctypes.sizeof(mv._lib.PipelineCfg) == 13 * 4 + 4 + 16  # 13 x 32-bit, pad to 8, two doubles
    assert mv.CameraRig.synthetic_ring(3).packed().shape == (3, 24)


def test_builtin_chains_match_oracle_tables(mv):
    for robot in ("fr3", "fr5", "meca500"):
        c, spec = mv.Chain.builtin(robot), O.chain_spec(robot)
        n = len(spec["a"])
        assert c.n_joints == n and c.n_points == n + 1
        assert c.struct.convention == (1 if spec["convention"] == "modified" else 0)
        np.testing.assert_allclose(list(c.struct.a)[:n], spec["a"], rtol=1e-7)
        np.testing.assert_allclose(list(c.struct.d)[:n], spec["d"], rtol=1e-7)
        np.testing.assert_allclose(list(c.struct.cos_alpha)[:n], np.cos(np.radians(spec["alpha_deg"])), atol=1e-7)
        np.testing.assert_allclose(list(c.struct.sin_alpha)[:n], np.sin(np.radians(spec["alpha_deg"])), atol=1e-7)
        np.testing.assert_allclose(list(c.struct.theta_offset)[:n], spec["theta_offset"])
        assert abs(c.struct.angle_scale - spec["angle_scale"]) < 1e-8
        for view in mv.VIEW_EULER_ZYX_DEG[robot]:
            np.testing.assert_allclose(mv.view_rotation(robot, view), O.view_rotation(robot, view), atol=1e-15)
    with pytest.raises(ValueError):
        mv._lib.check(mv._lib.load().mvgeo_chain_builtin(7, mv._lib.ChainStruct()), "x")
    with pytest.raises(ValueError):
        mv.Chain.from_dh([0.0] * 9, [0.0] * 9, [0.0] * 9, [0.0] * 9)


def test_argument_validation_without_a_gpu(mv):
    lib = mv._lib.load()
    z = ctypes.c_void_p(0)
    # bad sizes / enums are rejected before anything touches the device
    assert lib.mvgeo_decode(z, 0, -1, 8, 8, 1.0, 1.0, 0, 1.0, 0, 0, 1, 1, 0, z, z, z, z, z, z) == -1
    assert lib.mvgeo_decode(z, 9, 4, 8, 8, 1.0, 1.0, 0, 1.0, 0, 0, 1, 1, 0, z, z, z, z, z, z) == -1
    assert lib.mvgeo_decode(z, 0, 4, 8, 8, 1.0, 1.0, 1, 0.0, 0, 0, 1, 1, 0, z, z, z, z, z, z) == -1   # beta <= 0
    assert lib.mvgeo_decode(z, 0, 4, 8, 8, 1.0, 1.0, 2, 1.0, 16, 0, 1, 1, 0, z, z, z, z, z, z) == -1  # radius > 15
    assert lib.mvgeo_decode(z, 0, 4, 8, 8, 1.0, 1.0, 0, 1.0, 0, 0, 1, 1, 0, z, z, z, z, z, z) == -2   # NULL maps
    assert lib.mvgeo_decode(z, 0, 0, 8, 8, 1.0, 1.0, 0, 1.0, 0, 0, 1, 1, 0, z, z, z, z, z, z) == 0    # empty is fine
    assert lib.mvgeo_triangulate(z, z, z, 4, 1 + mv._lib.MAX_VIEWS, 8, 0.0, 0, z, z, z, z) == -1
    assert lib.mvgeo_triangulate(z, z, z, 4, 4, 8, 0.0, 0, z, z, z, z) == -2
    c = mv.Chain.builtin("fr3")
    assert lib.mvgeo_fk(ctypes.byref(c.struct), z, 4, z, 3, z, z) == -1     # V > 1 needs R_view
    assert lib.mvgeo_fk(ctypes.byref(c.struct), z, 0, z, 1, z, z) == 0
    bad = mv._lib.ChainStruct()
    assert lib.mvgeo_fk(ctypes.byref(bad), z, 4, z, 1, z, z) == -1
    assert lib.mvgeo_encode_gaussian(z, 4, 8, 8, 0.0, 0, z, z) == -1        # sigma <= 0
    with pytest.raises(ValueError, match="no CPU path"):
        mv.decode_heatmaps(torch.zeros(2, 8, 8))
    with pytest.raises(ValueError, match="no CPU path"):
        mv.triangulate(torch.zeros(1, 2, 3, 2), torch.zeros(2, 3, 4))


def test_no_cpu_fallback_and_oracle_isolation(mv):
    """The product package never imports oracle/ and fails loudly without its library."""
    pkg = os.path.join(ROOT, "2025_icra_multi_view_robot_pose_estimation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), fn
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            mv.compat.fr3.angle_to_joint_coordinate([0.0] * 7, "view1")
    code = ("import sys; sys.path.insert(0, %r); import mvgeo; mvgeo._lib.LIB_PATH = '/nonexistent/libmvgeo.so'; "
            "mvgeo._lib._lib = None\ntry:\n    mvgeo._lib.load()\nexcept mvgeo.MvgeoError as e:\n    print('LOUD', e)" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "LOUD" in out.stdout and "no CPU" in out.stdout, out.stdout + out.stderr


# ------------------------------------------------------------------------------ rig
def test_rig_intrinsics_match_reference_parser(mv):
    names = [f"{sn}_{side}" for side in ("left", "right") for sn in ("41182735", "49429257", "44377151", "49045152")]
    for i, n in enumerate(names):
        fx, fy, cx, cy, *dist = mv.ZEDX_FHD1200[n]
        np.testing.assert_array_equal([fx, fy, cx, cy], [G["zedx_K"][i][0, 0], G["zedx_K"][i][1, 1], G["zedx_K"][i][0, 2], G["zedx_K"][i][1, 2]])
        np.testing.assert_array_equal(dist, G["zedx_dist"][i])


def test_conf_parser(mv, tmp_path):
    p = tmp_path / "SN1.conf"
    p.write_text("\ufeff[LEFT_CAM_FHD1200]\nfx = 700.5\nfy = 701.5\ncx = 960.25\ncy = 600.75\nk1 = -0.01\nk2 = 0.02\nk3 = 0.003\n"
                 "p1 = 1e-4\np2 = -2e-4\n\n[LEFT_CAM_FHD]\nfx=1066.51\nfy=1066.89\ncx=989.51\ncy=578.779\nk1=-0.05\nk2=0.02\nk3=0\np1=0\np2=0\n"
                 "[LEFT_DISTO]\nk4 = 0.5\n", encoding="utf-8")
    K, dist, adv = mv.load_conf_calibration(str(p), "left")
    assert K == [[700.5, 0.0, 960.25], [0.0, 701.5, 600.75], [0.0, 0.0, 1.0]]
    assert dist == [-0.01, 0.02, 1e-4, -2e-4, 0.003] and adv == {"k4": 0.5}  # OpenCV order k1 k2 p1 p2 k3
    K2, _, _ = mv.load_conf_calibration(str(p), "left", "FHD")
    assert K2[0][0] == 1066.51
    ref_conf = "/root/reference/dataset/All_camera_conf/SN41182735.conf"
    if os.path.isfile(ref_conf):
        K3, d3, _ = mv.load_conf_calibration(ref_conf, "left")
        np.testing.assert_array_equal(K3, G["zedx_K"][0])
        np.testing.assert_array_equal(d3, G["zedx_dist"][0])


def test_rig_projection_matrices_and_aruco(mv):
    rig = mv.CameraRig.synthetic_ring(4, distortion=True)
    P = rig.projection_matrices()
    X = np.array([[0.1, -0.2, 0.5], [0.0, 0.0, 0.0]])
    for v in range(4):
        np.testing.assert_allclose(P[v], O.projection_matrix(rig.K[v], rig.R[v], rig.t[v]), rtol=1e-6)
        h = P[v].astype(np.float64) @ np.append(X[0], 1.0)
        np.testing.assert_allclose(h[:2] / h[2], O.project_points(X[0], rig.R[v], rig.t[v], rig.K[v]), rtol=1e-5)
        assert (rig.R[v] @ X[1] + rig.t[v])[2] > 0.5  # cameras look at the workspace
        np.testing.assert_allclose(rig.R[v] @ rig.R[v].T, np.eye(3), atol=1e-12)
    Rv = np.stack([O.view_rotation("fr5", v) for v in ("top", "left", "right")] + [np.eye(3)])
    Pv = rig.projection_matrices(Rv)
    np.testing.assert_allclose(Pv[0][:, :3], (rig.K[0] @ rig.R[0] @ Rv[0]), rtol=1e-6, atol=1e-4)
    rec = [dict(rvec_x=0.1, rvec_y=-0.2, rvec_z=0.3, tvec_x=0.1, tvec_y=0.0, tvec_z=1.5)]
    r = mv.CameraRig.from_aruco(rec, G["zedx_K"][:1], None)
    np.testing.assert_allclose(r.R[0], O.rodrigues([0.1, -0.2, 0.3]), atol=1e-15)
    rd = mv.CameraRig.from_aruco([dict(rec[0], rvec_x=np.degrees(0.1), rvec_y=np.degrees(-0.2), rvec_z=np.degrees(0.3))],
                                 G["zedx_K"][:1], None, rvec_in_degrees=True)
    np.testing.assert_allclose(rd.R[0], r.R[0], atol=1e-14)
    pk = r.packed()[0]
    np.testing.assert_allclose(pk[12:16], [737.118, 737.085, 974.584, 552.68], rtol=1e-7)
    np.testing.assert_allclose(mv.rodrigues([0.3, 0.2, -0.9]), O.rodrigues([0.3, 0.2, -0.9]), atol=1e-15)


# ------------------------------------------------------------------------- sharding
def test_frame_range_partition(mv):
    fr = mv.sharding.frame_range
    for n in (0, 1, 7, 8, 65536, 1000003):
        for ws in (1, 2, 3, 4, 8):
            r = [fr(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes)  # remainder on the last ranks
    with pytest.raises(ValueError):
        fr(10, 3, 2)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import mvgeo
rank, ws, n = int(sys.argv[1]), 2, 11
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT={port!r}, RANK=str(rank), WORLD_SIZE=str(ws))
dist.init_process_group("gloo", rank=rank, world_size=ws)
a, b = mvgeo.sharding.frame_range(n, rank, ws)
full = dict(X=torch.arange(n * 6, dtype=torch.float32).reshape(n, 2, 3), idx=torch.arange(n, dtype=torch.int32) * 7)
local = {{k: v[a:b].clone() for k, v in full.items()}}
local["loss"] = torch.tensor(1.0)
out = mvgeo.sharding.gather_frames(local, n)
assert set(out) == {{"X", "idx"}}
for k in full:
    assert torch.equal(out[k], full[k]), k
pr = mvgeo.sharding.PackedResults({{"X": ((3, 2), torch.float32), "i": ((3,), torch.int32)}}, "cpu")
pr["X"].copy_(torch.arange(6, dtype=torch.float32).reshape(3, 2) + 100 * rank); pr["i"].copy_(torch.arange(3, dtype=torch.int32) - rank)
g = pr.all_gather()
assert g["X"].shape == (2, 3, 2) and g["i"].shape == (2, 3) and g["i"].dtype == torch.int32
for r in range(ws):
    assert torch.equal(g["X"][r], torch.arange(6, dtype=torch.float32).reshape(3, 2) + 100 * r)
    assert torch.equal(g["i"][r], torch.arange(3, dtype=torch.int32) - r)
ring = mvgeo.sharding.ResultRing({{"X": ((3, 2), torch.float32), "n": ((3,), torch.int32)}}, 4, "cpu")
for k in range(3):  # a 3-batch job
    ring.slot[k]["X"].fill_(10.0 * rank + k); ring.slot[k]["n"].fill_(100 * rank + k)
allr = ring.final_gather(3)
assert allr.shape == (ws, 3, 9)
for r in range(ws):
    for k in range(3):
        assert torch.all(allr[r, k, :6] == 10.0 * r + k) and torch.all(allr[r, k, 6:].view(torch.int32) == 100 * r + k)
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
"""


def test_gather_frames_gloo_world_size_2(mv, tmp_path):
    port = str(29500 + os.getpid() % 2000)
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for r, p in enumerate(procs):
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0 and f"OK {r}" in out, err[-2000:]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU port of the path) prints the contract's JSON line."""
    import json

    env = dict(os.environ, MVGEO_BENCH_REF_SECONDS="0.5")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"] == "frames/s decode+triangulate+FK" and d["config"]["workload"].startswith("C2")
    assert d["config"]["views"] == 4 and d["config"]["map"] == [240, 320] and d["config"]["map_dtype"] == "bf16"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    # other ranks of a torchrun launch exit 0 without work
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, env=dict(env, RANK="3", WORLD_SIZE="8"), cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_header_is_plain_c_and_library_is_callable_from_c(mv, tmp_path):
    """include/mvgeo.h compiles as C99 (no C++-isms) and a C program can dlopen the library and call the
    host-only entry points — the boundary really is a C ABI, not a Python/C++ one."""
    import shutil

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "mvgeo.h"
typedef int (*version_fn)(void);
typedef const char* (*errstr_fn)(int);
typedef int (*chain_fn)(int, mvgeo_chain*);
typedef int (*decode_fn)(const void*, int, int64_t, int, int, double, double, int, float, int, int, int64_t, int64_t,
                         int64_t, int32_t*, float*, float*, float*, float*, void*);
int main(int argc, char** argv) {
  if (argc < 2) return 1;
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  version_fn ver = (version_fn)dlsym(h, "mvgeo_version");
  errstr_fn es = (errstr_fn)dlsym(h, "mvgeo_error_string");
  chain_fn cb = (chain_fn)dlsym(h, "mvgeo_chain_builtin");
  decode_fn dec = (decode_fn)dlsym(h, "mvgeo_decode");
  if (!ver || !es || !cb || !dec) return 3;
  mvgeo_chain c;
  memset(&c, 0, sizeof c);
  if (ver() != MVGEO_VERSION) return 4;
  if (cb(MVGEO_ROBOT_FR3, &c) != MVGEO_OK || c.n_joints != 7 || c.convention != MVGEO_DH_MODIFIED || !c.emit_base) return 5;
  if (cb(MVGEO_ROBOT_MECA500, &c) != MVGEO_OK || c.n_joints != 6 || c.theta_offset[1] != -90.0f) return 6;
  if (cb(99, &c) != MVGEO_EINVAL) return 7;
  if (dec(NULL, MVGEO_F32, 4, 8, 8, 1.0, 1.0, MVGEO_SOFT_NONE, 1.0f, 0, 0, 1, 1, 0, NULL, NULL, NULL, NULL, NULL, NULL) != MVGEO_ENULL) return 8;
  if (dec(NULL, MVGEO_F32, 0, 8, 8, 1.0, 1.0, MVGEO_SOFT_NONE, 1.0f, 0, 0, 1, 1, 0, NULL, NULL, NULL, NULL, NULL, NULL) != MVGEO_OK) return 9;
  printf("%d %s | sizeof(mvgeo_chain)=%zu sizeof(mvgeo_camera)=%zu sizeof(mvgeo_pipeline_out)=%zu\n", ver(), es(MVGEO_EALIGN),
         sizeof(mvgeo_chain), sizeof(mvgeo_camera), sizeof(mvgeo_pipeline_out));
  return 0;
}
''')
    exe = tmp_path / "abi"
    hdr_only = tmp_path / "hdr.c"
    hdr_only.write_text('#include "mvgeo.h"\nint mvgeo_header_ok;\n')
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I",
                         os.path.join(ROOT, "include"), str(hdr_only)], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr  # the header itself is strictly conforming C99
    # (the program is built without -pedantic only because ISO C has no object->function pointer cast for dlsym)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                         str(src), "-o", str(exe), "-ldl"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe), mv.LIB_PATH], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert run.stdout.startswith("100 pointer is not aligned")
    assert f"sizeof(mvgeo_chain)={ctypes.sizeof(mv._lib.ChainStruct)} " in run.stdout  # ctypes mirror == C layout
    assert "sizeof(mvgeo_camera)=96 " in run.stdout and f"sizeof(mvgeo_pipeline_out)={ctypes.sizeof(mv._lib.PipelineOut)}" in run.stdout


def test_numa_binding_helper_is_safe_without_a_gpu(mv):
    """sharding.bind_to_gpu_numa never raises: unknown topology (no NVML / no GPU) returns None and leaves the
    affinity untouched."""
    before = os.sched_getaffinity(0)
    r = mv.sharding.bind_to_gpu_numa(0)
    assert r is None or set(r) <= set(before)
    if r is None:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)
    assert isinstance(mv.sharding.gpu_local_cpus(0), list)


def test_new_entry_points_validate_arguments_without_a_gpu(mv):
    lib = mv._lib.load()
    z = ctypes.c_void_p(0)
    views = (ctypes.c_void_p * 2)(None, None)
    assert lib.mvgeo_decode_views(views, 0, 0, 4, 8, 8, 8, 1.0, 1.0, 0, 1.0, 0, 0, z, z, z, z, z, z) == -1          # no views
    assert lib.mvgeo_decode_views(views, 2, 0, 4, 8, 8, 8, 1.0, 1.0, 0, 1.0, 0, 0, z, z, z, z, z, z) == -2          # NULL view
    assert lib.mvgeo_decode_views(views, 2, 0, 0, 8, 8, 8, 1.0, 1.0, 0, 1.0, 0, 0, z, z, z, z, z, z) == 0           # empty
    assert lib.mvgeo_decode_mse(z, 0, 4, 8, 8, 1.0, 1.0, 0, z, 0.0, 1.0, z, z, z, z, z, z, z) == -1                 # sigma <= 0
    assert lib.mvgeo_decode_mse(z, 0, 4, 8, 8, 1.0, 1.0, 0, z, 2.0, 1.0, z, z, z, z, z, z, z) == -2                 # NULL maps
    assert lib.mvgeo_pnp_solve(z, 0, z, z, z, 4, 2, 17, 0.0, 8.0, 10, z, z, z, z, z, z) == -1                       # K > 16
    assert lib.mvgeo_pnp_solve(z, 0, z, z, z, 4, 2, 8, 0.0, 0.0, 10, z, z, z, z, z, z) == -1                        # threshold <= 0
    assert lib.mvgeo_pnp_solve(z, 0, z, z, z, 4, 2, 8, 0.0, 8.0, 10, z, z, z, z, z, z) == -2
    assert lib.mvgeo_pnp_solve(z, 0, z, z, z, 0, 2, 8, 0.0, 8.0, 10, z, z, z, z, z, z) == 0
