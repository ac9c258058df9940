"""CPU: pin the oracle (oracle/mvgeo_oracle.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py -> reference_golden.npz), against the
known-answer values of SURVEY.md section 8c, and against cv2 / scipy where they import."""
import hashlib
import importlib.util
import math
import os

import numpy as np
import pytest

from oracle import mvgeo_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))

spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
_mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(_mg)
INP = _mg.gen_inputs()


def test_inputs_regenerate_bit_exactly():
    for k, v in INP.items():
        assert hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() == str(G[f"sha_{k}"]), k


# ------------------------------------------------------------------ FK
def test_fk_fr3_golden():
    for view, key in (("view1", "fr3_fk_view1"), ("none", "fr3_fk_noview")):
        got = np.stack([O.fk_fr3(q, view) for q in INP["fr3_q"]])
        assert got.dtype == np.float32 and got.shape == (48, 8, 3)
        np.testing.assert_allclose(got, G[key], rtol=0, atol=6e-8)  # <= 1 ulp of a float32 metre


def test_fk_fr3_known_answers():
    # SURVEY.md 8c
    p = O.fk_fr3([0.6480, -0.1083, 0.2098, -1.9201, 0.8858, 3.1025, -2.3876], "view1")
    np.testing.assert_allclose(p[3], [-0.0206163, -0.0272323, -0.6471487], atol=2e-7)
    np.testing.assert_allclose(p[7], [0.3067923, 0.3452037, -0.6980957], atol=2e-7)
    np.testing.assert_allclose(O.view_rotation("fr3", "view1"), [[0, 1, 0], [1, 0, 0], [0, 0, -1]], atol=1e-15)
    z = O.fk_fr3([0.0] * 7, None)
    np.testing.assert_allclose(z[:, 2], [0, 0.333, 0.333, 0.649, 0.649, 1.033, 1.033, 1.033], atol=1e-6)
    np.testing.assert_allclose(z[[4, 7], 0], [0.0825, 0.088], atol=1e-6)


def test_fk_fr5_golden_and_known():
    q = np.concatenate([INP["fr5_q_rand"], G["fr5_q_real"]], axis=0)
    for view in ("top", "left", "right", "none"):
        got = np.stack([O.fk_fr5(a, view) for a in q])
        np.testing.assert_allclose(got, G[f"fr5_fk_{view}"], rtol=0, atol=6e-8)
    z = O.fk_fr5([0.0] * 6, None)
    np.testing.assert_allclose(z[1:], [[0, 0, .152], [-.425, 0, .152], [-.82, 0, .152], [-.82, -.102, .152],
                                       [-.82, -.102, .05], [-.82, -.202, .05]], atol=1e-6)
    last = O.fk_fr5([-60.6619353341584, -95.85930248298267, 117.4375386757425, -111.5669235380569,
                     -90.00021755105197, 29.3358891553], "top")[-1]
    np.testing.assert_allclose(last, [0.2941491, -0.3244689, -0.3294898], atol=5e-7)


def test_fk_meca500_golden_and_known():
    got = np.stack([O.fk_meca500(a) for a in INP["meca_q"]])
    np.testing.assert_allclose(got, G["meca_fk"], rtol=0, atol=6e-8)
    np.testing.assert_allclose(got[0][1:], [[0, 0, .135], [0, 0, .27], [0, 0, .308], [.12, 0, .308], [.12, 0, .308],
                                            [.19, 0, .308]], atol=1e-6)
    np.testing.assert_allclose(got[1][-1], [0.1200077, -0.0138394, 0.2301765], atol=2e-7)


def test_fk_generic_class_golden():
    got = O.fk_generic([tuple(r) for r in INP["generic_dh"]], INP["generic_angles"])
    np.testing.assert_allclose(got, G["generic_fk"], rtol=0, atol=1e-6)  # reference accumulates f32 link matrices


def test_fk_chain_vectorised_matches_per_robot():
    Rv = np.stack([O.view_rotation("fr5", v) for v in ("top", "left", "right")])
    q = np.concatenate([INP["fr5_q_rand"], G["fr5_q_real"]], axis=0)
    X = O.fk_chain(O.chain_spec("fr5"), q, Rv)
    for vi, view in enumerate(("top", "left", "right")):
        np.testing.assert_allclose(X[:, vi], G[f"fr5_fk_{view}"], rtol=0, atol=2e-7)
    X = O.fk_chain(O.chain_spec("fr3"), INP["fr3_q"], O.view_rotation("fr3", "view1")[None])
    np.testing.assert_allclose(X[:, 0], G["fr3_fk_view1"], rtol=0, atol=2e-7)
    X = O.fk_chain(O.chain_spec("meca500"), INP["meca_q"])
    np.testing.assert_allclose(X[:, 0], G["meca_fk"], rtol=0, atol=2e-7)
    gen = dict(convention="standard", emit_base=False, angle_scale=1.0, a=list(INP["generic_dh"][:, 2]),
               d=list(INP["generic_dh"][:, 1]), alpha_rad=list(INP["generic_dh"][:, 3]),
               theta_offset=list(INP["generic_dh"][:, 0]))
    X = O.fk_chain(gen, INP["generic_angles"])
    np.testing.assert_allclose(X[:, 0], G["generic_fk"], rtol=0, atol=1e-6)


def test_view_rotations_match_scipy():
    R = pytest.importorskip("scipy.spatial.transform").Rotation
    for robot, table in O.VIEW_EULER_ZYX_DEG.items():
        for view, ang in table.items():
            np.testing.assert_allclose(O.view_rotation(robot, view), R.from_euler("zyx", list(ang), degrees=True).as_matrix(),
                                       atol=1e-15)


# ------------------------------------------------------------------ projection
def _cam(i, dist):
    return O.rodrigues(INP["proj_rvec"][i]), INP["proj_tvec"][i], G["zedx_K"][0].astype(np.float32), dist


def test_projection_golden():
    d_real = G["zedx_dist"][0].astype(np.float32)
    for i in range(6):
        R, t, K, _ = _cam(i, None)
        X = G["fr3_fk_view1"][i]
        np.testing.assert_allclose(O.project_points(X, R, t, K, None), G["proj_fr3_zero"][i], rtol=0, atol=2e-4)
        np.testing.assert_allclose(O.project_points(X, R, t, K, d_real), G["proj_fr3_real"][i], rtol=0, atol=2e-4)
        # the Fr5 wrapper casts rvec (deg->rad) and tvec to float32 before cv2 (Fr5_model_train.ipynb:291-300)
        r32 = np.array([math.radians(float(np.degrees(v))) for v in INP["proj_rvec"][i]], dtype=np.float32)
        t32 = INP["proj_tvec"][i].astype(np.float32)
        got = O.project_points(G["fr5_fk_top"][i], O.rodrigues(r32), t32, K, d_real)
        np.testing.assert_allclose(got, G["proj_fr5_real_degrvec"][i], rtol=0, atol=2e-4)
    # known answer: first / last point of the FR3 pose, zero distortion
    np.testing.assert_allclose(G["proj_fr3_zero"][0][0], [1023.72516, 552.68], atol=1e-3)
    np.testing.assert_allclose(G["proj_fr3_zero"][0][-1], [1306.5613, 962.1317], atol=1e-3)
    R = O.rodrigues(np.deg2rad(np.array([96, 98, -45], dtype=np.float32)))
    got = O.project_points(G["meca_fk"][1], R, np.array([0, -0.01, 0.75], dtype=np.float32), G["zedx_K"][0].astype(np.float32), d_real)
    np.testing.assert_allclose(got, G["proj_meca_prior"], rtol=0, atol=2e-4)
    for i in range(6):
        R, t, K, _ = _cam(i, None)
        np.testing.assert_allclose(O.project_points(G["generic_fk"][i], R, t, K, None), G["proj_generic"][i], rtol=0, atol=3e-4)


def test_projection_and_rodrigues_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for _ in range(20):
        rvec = rng.uniform(-2.5, 2.5, 3)
        np.testing.assert_allclose(O.rodrigues(rvec), cv2.Rodrigues(rvec)[0], atol=1e-13)
        tvec = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(1.0, 3.0)])
        X = rng.uniform(-0.7, 0.7, size=(9, 3))
        K = G["zedx_K"][rng.integers(0, 8)]
        dist = G["zedx_dist"][rng.integers(0, 8)]
        ref = cv2.projectPoints(X, rvec, tvec, K, dist)[0].reshape(-1, 2)
        np.testing.assert_allclose(O.project_points(X, O.rodrigues(rvec), tvec, K, dist), ref, rtol=0, atol=1e-9)
    np.testing.assert_array_equal(O.rodrigues(np.zeros(3)), np.eye(3))


# ------------------------------------------------------------------ decode
def test_decode_golden_sigmoid_variant():
    for f in range(5):
        d = O.decode(INP["maps_small"][f], 1920 / 40, 1200 / 24, apply_sigmoid=True)
        # raw arg-max == sigmoid arg-max on these maps (no sigmoid tie collapse; checked, not assumed)
        np.testing.assert_array_equal(d["idx"], G["dec_small_rawidx"][f])
        np.testing.assert_array_equal(d["kp_hard"], G["dec_small_kp"][f])
        np.testing.assert_allclose(d["score"], G["dec_small_score"][f], rtol=2e-7)
    d = O.decode(INP["maps_native"], 1920 / 128, 1080 / 128, apply_sigmoid=True)
    ref_kp = G["dec_native_kp"]
    # tie class: the reference argmaxes sigmoid(h); where the indices differ the sigmoid values tie
    same = np.all(d["kp_hard"] == ref_kp, axis=-1)
    np.testing.assert_allclose(d["score"], G["dec_native_score"], rtol=2e-7)
    assert same.sum() >= 6


def test_decode_port_matches_reference_bitwise():
    torch = pytest.importorskip("torch")
    for f in range(5):
        kp, sc = O.extract_keypoints_from_heatmaps(torch.from_numpy(INP["maps_small"][f]), (1200, 1920))
        np.testing.assert_array_equal(kp, G["dec_small_kp"][f])
        np.testing.assert_array_equal(sc, G["dec_small_score"][f])
    kp, sc = O.extract_keypoints_from_heatmaps(torch.from_numpy(INP["maps_native"]), (1080, 1920))
    np.testing.assert_array_equal(kp, G["dec_native_kp"])
    np.testing.assert_array_equal(sc, G["dec_native_score"])


def test_decode_first_maximum_tie_rule():
    idx, peak = O.argmax_first(INP["maps_ties"])
    np.testing.assert_array_equal(idx, G["dec_ties_idx"])
    assert np.all(peak == 0.75)
    a = np.array([[[0, 5, 5], [5, 1, 5]]], dtype=np.float32)
    assert O.argmax_first(a)[0][0] == 1
    a = np.array([[[0, np.nan, 9], [np.nan, 1, 5]]], dtype=np.float32)
    assert O.argmax_first(a)[0][0] == 1  # NaN is maximal, first NaN wins (torch.argmax semantics)
    torch = pytest.importorskip("torch")
    assert int(torch.argmax(torch.from_numpy(a[0]))) == 1
    z = np.array([[[-0.0, 0.0, -1.0]]], dtype=np.float32)
    assert O.argmax_first(z)[0][0] == 0 and int(torch.argmax(torch.from_numpy(z[0]))) == 0  # -0 == +0: first wins


def test_inline_argmax_port():
    torch = pytest.importorskip("torch")
    kps = O.decode_inline_argmax(torch.from_numpy(INP["maps_ties"][0]), (1200, 1920))
    idx = G["dec_ties_idx"][0]
    np.testing.assert_allclose(kps[:, 0], (idx % 40) * (1920 / 40))
    np.testing.assert_allclose(kps[:, 1], (idx // 40) * (1200 / 30))


def test_soft_argmax_properties():
    H, W = 48, 64
    yy, xx = np.mgrid[0:H, 0:W]
    c = (23.3, 17.8)
    g = np.exp(-((xx - c[0]) ** 2 + (yy - c[1]) ** 2) / (2 * 2.5 ** 2))
    s = O.soft_argmax(g, beta=25.0, mode="global")
    np.testing.assert_allclose(s, c, atol=0.05)  # symmetric blob: centroid at the true sub-pixel centre
    s = O.soft_argmax(g, beta=25.0, mode="window", radius=5)
    np.testing.assert_allclose(s, c, atol=0.05)
    hard = O.soft_argmax(g, beta=1e4, mode="global")
    np.testing.assert_allclose(hard, (23, 18), atol=1e-6)  # beta -> inf collapses to the hard peak
    assert np.all(np.isnan(O.soft_argmax(np.full((4, 4), np.nan), 1.0)))
    np.testing.assert_array_equal(O.soft_argmax(np.full((4, 4), -np.inf), 1.0), [0, 0])


# ------------------------------------------------------------------ triangulation
def _rig(V, rng):
    cams, P = [], []
    for v in range(V):
        ang = 2 * np.pi * v / V + 0.3
        c = np.array([1.5 * np.cos(ang), 1.5 * np.sin(ang), 0.8])
        z = np.array([0, 0, 0.4]) - c
        z /= np.linalg.norm(z)
        x = np.cross(z, [0, 0, 1.0])
        x /= np.linalg.norm(x)
        R = np.stack([x, np.cross(z, x), z])
        t = -R @ c
        K = G["zedx_K"][v % 8]
        cams.append((R, t, K))
        P.append(O.projection_matrix(K, R, t))
    return cams, np.array(P)


def test_triangulate_closed_loop_and_cv2():
    rng = np.random.default_rng(3)
    cams, P = _rig(4, rng)
    X = rng.uniform(-0.5, 0.5, size=(5, 6, 3)) + [0, 0, 0.4]
    kp = np.stack([O.project_points(X, R, t, K) for R, t, K in cams], axis=1)  # (B,V,K,2)
    Xt, resid, nv = O.triangulate_dlt(kp, P)
    np.testing.assert_allclose(Xt, X, atol=1e-9)
    assert np.all(nv == 4) and np.all(resid < 1e-8)
    # fewer than two valid views -> NaN; weights below the threshold drop a view
    w = np.ones((5, 4, 6))
    w[0, 1:, 2] = 0.1
    Xt2, _, nv2 = O.triangulate_dlt(kp, P, w, min_weight=0.5)
    assert nv2[0, 2] == 1 and np.all(np.isnan(Xt2[0, 2]))
    w[1, 2:, 3] = 0.1
    Xt3, _, nv3 = O.triangulate_dlt(kp, P, w, min_weight=0.5)
    assert nv3[1, 3] == 2
    np.testing.assert_allclose(Xt3[1, 3], X[1, 3], atol=1e-9)
    cv2 = pytest.importorskip("cv2")
    noisy = kp + rng.normal(0, 0.7, kp.shape)
    Xn, _, _ = O.triangulate_dlt(noisy[:, :2], P[:2])
    for b in range(5):
        h = cv2.triangulatePoints(P[0], P[1], noisy[b, 0].T.copy(), noisy[b, 1].T.copy())
        np.testing.assert_allclose(Xn[b], (h[:3] / h[3]).T, rtol=1e-7, atol=1e-9)


# ------------------------------------------------------------------ loss / gradient
def test_reprojection_loss_matches_reference_form_and_gradcheck():
    torch = pytest.importorskip("torch")
    # reference robot_pose_loss with only the FK term: lambda_fk * mse(proj, gt)
    pred, gt = INP["loss_pred"].astype(np.float64) + 1.5, INP["loss_gt"].astype(np.float64)
    assert abs(2.5 * np.mean((np.float32(pred) - np.float32(gt)) ** 2) - float(G["loss_fk_only"])) < 1e-3
    rng = np.random.default_rng(5)
    cams_, _ = _rig(3, rng)
    for robot, lo, hi in (("fr3", -2.0, 2.0), ("fr5", -170, 170), ("meca500", -120, 120)):
        spec_ = O.chain_spec(robot)
        J = len(spec_["a"])
        cams = [dict(R=R, t=t, K=K, dist=G["zedx_dist"][i]) for i, (R, t, K) in enumerate(cams_)]
        Rv = np.stack([O.view_rotation(robot, v) for v in (list(O.VIEW_EULER_ZYX_DEG[robot]) + [None] * 3)[:3]])
        q = torch.tensor(rng.uniform(lo, hi, size=(2, J)), dtype=torch.float64, requires_grad=True)
        with torch.no_grad():
            _, _, uv = O.fk_reproj_loss_torch(spec_, q, Rv, cams, np.zeros((2, 3, J + 1, 2)))
        gt_uv = uv.numpy() + rng.normal(0, 3.0, uv.shape)
        f = lambda qq: O.fk_reproj_loss_torch(spec_, qq, Rv, cams, gt_uv, None, 0.7)[0]
        assert torch.autograd.gradcheck(f, (q,), eps=1e-6, atol=1e-5, rtol=1e-5)
        # forward of the autograd restatement equals the numpy FK + projection restatements
        X = O.fk_chain(spec_, q.detach().numpy(), Rv)
        _, Xt, uvt = O.fk_reproj_loss_torch(spec_, q, Rv, cams, gt_uv)
        np.testing.assert_allclose(Xt.detach().numpy(), X, atol=1e-12)
        for v, cam in enumerate(cams):
            np.testing.assert_allclose(uvt[:, v].detach().numpy(), O.project_points(X[:, v], cam["R"], cam["t"], cam["K"], cam["dist"]),
                                       atol=1e-9)


# ------------------------------------------------------------------ GT maps
def test_gt_heatmap_golden():
    got = np.stack([O.create_gt_heatmap(tuple(k), (128, 128), 5.0) for k in INP["gt_kp"]])
    np.testing.assert_array_equal(got, G["gt_maps_128"])
    m = got[0]
    assert np.unravel_index(np.argmax(m), m.shape) == (78, 40)
    assert abs(m.max() - 0.9974034) < 1e-7 and int((m > 0).sum()) == 5638
    got = np.stack([O.create_gt_heatmap((k[0] * 0.3, k[1] * 0.2), (24, 40), 3.0) for k in INP["gt_kp"]])
    np.testing.assert_array_equal(got, G["gt_maps_rect"])


def test_heatmap_mse_restatement():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    pred = rng.normal(0, 0.3, size=(3, 24, 40))
    kp = np.array([[10.2, 7.7], [np.nan, 3.0], [39.0, 23.0]])
    loss, grad = O.heatmap_mse(pred, kp, 3.0, 100.0)
    gt = np.stack([O.create_gt_heatmap(tuple(kp[0]), (24, 40), 3.0), np.zeros((24, 40)), O.create_gt_heatmap(tuple(kp[2]), (24, 40), 3.0)])
    p = torch.tensor(pred, requires_grad=True)
    l = torch.nn.MSELoss()(p, torch.tensor(gt)) * 100.0
    l.backward()
    assert abs(loss - float(l)) < 1e-12
    np.testing.assert_allclose(grad, p.grad.numpy(), atol=1e-15)


# ------------------------------------------------------------------ camera-pose refinement
def test_pnp_refine_oracle_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    for trial in range(6):
        Kc, dist = G["zedx_K"][trial % 8], G["zedx_dist"][trial % 8]
        rv_true = rng.uniform(-0.8, 0.8, 3)
        t_true = np.array([rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), rng.uniform(1.2, 2.5)])
        X = O.fk_fr3(rng.uniform(-2, 2, 7), "view1").astype(np.float64) + rng.normal(0, 0.02, (8, 3))
        kp = O.project_points(X, O.rodrigues(rv_true), t_true, Kc, dist) + rng.normal(0, 0.5, (8, 2))
        rv0, t0 = rv_true + rng.normal(0, 0.05, 3), t_true + rng.normal(0, 0.05, 3)
        rv, tv, rms, st = O.pnp_refine(X, kp, Kc, dist, O.rodrigues(rv0), t0)
        ok, rv_cv, t_cv = cv2.solvePnP(X, kp, Kc, dist, rv0.reshape(3, 1).copy(), t0.reshape(3, 1).copy(), True, cv2.SOLVEPNP_ITERATIVE)
        assert ok and (st & 3) == 3 and (st & 4)
        # same minimum of the same cost: cv2's LM stops at its own tolerance, ours goes to machine precision
        np.testing.assert_allclose(rv, rv_cv.ravel(), atol=2e-4)
        np.testing.assert_allclose(tv, t_cv.ravel(), atol=2e-4)
        e_cv = O.project_points(X, O.rodrigues(rv_cv.ravel()), t_cv.ravel(), Kc, dist) - kp
        assert rms <= math.sqrt(np.mean(np.sum(e_cv ** 2, axis=1))) + 1e-9
        np.testing.assert_allclose(O.rvec_from_matrix(O.rodrigues(rv_true)), rv_true, atol=1e-12)
    rv, tv, rms, st = O.pnp_refine(X[:3], kp[:3], Kc, dist, O.rodrigues(rv0), t0)   # < 4 points: the prior comes back
    assert st == 0 and np.isnan(rms)
    np.testing.assert_allclose(rv, rv0, atol=1e-12)
    np.testing.assert_allclose(tv, t0, atol=1e-12)


def test_undistort_points_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    for i in range(8):
        Kc, dist = G["zedx_K"][i], G["zedx_dist"][i]
        kp = np.stack([rng.uniform(0, 1920, 50), rng.uniform(0, 1200, 50)], axis=1)
        ref = cv2.undistortPoints(kp.reshape(-1, 1, 2), Kc, dist, P=Kc).reshape(-1, 2)
        np.testing.assert_allclose(O.undistort_points(kp, Kc, dist), ref, rtol=0, atol=1e-9)
        # round trip: distort(undistort(p)) == p (what makes raw-image key-points triangulable)
        und = O.undistort_points(kp, Kc, dist, iters=20)
        xn = np.concatenate([(und - [Kc[0, 2], Kc[1, 2]]) / [Kc[0, 0], Kc[1, 1]], np.ones((50, 1))], axis=1)
        np.testing.assert_allclose(O.project_points(xn, np.eye(3), np.zeros(3), Kc, dist), kp, atol=1e-6)


def test_average_quaternion_matches_reference_function():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference checkout not present")
    import importlib.util, types
    src = open(os.path.join(ref_loader.REF_ROOT, "dataset", "Fr5_preprocessing.py")).read()
    ns = dict(np=np)
    exec(ref_loader._extract_defs(src, ["average_quaternion"]), ns)
    rng = np.random.default_rng(5)
    q = rng.normal(size=(9, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    ref = ns["average_quaternion"](q)
    got = O.average_quaternion(q)
    assert abs(abs(ref @ got) - 1.0) < 1e-12


def test_pnp_solve_oracle_p3p_consensus_vs_cv2():
    """The specification of mvgeo_pnp_solve (every-triplet P3P consensus + LM) pinned against OpenCV: P3P returns the
    true pose among its solutions; with 0-2 gross outliers of 8 points the consensus set is exactly the clean points and
    the refined pose equals cv2.solvePnP(SOLVEPNP_ITERATIVE) on them; cv2.solvePnPRansac(SOLVEPNP_EPNP) — the call being
    replaced (model/Fr5_model_train.ipynb:4735-4741) — lands within its own (non-optimal, random) accuracy of it."""
    import cv2

    rng = np.random.default_rng(0)
    K = np.array([[1066.5, 0, 989.5], [0, 1066.9, 578.8], [0, 0, 1.0]])
    dist = np.array([-0.005, -0.046, 1e-4, 3e-4, 0.0148])
    # P3P alone, noise-free: one of the solutions is the truth
    for _ in range(20):
        rv = rng.normal(size=3)
        R, t = O.rodrigues(rv), np.array([0.1, -0.2, rng.uniform(1.0, 3.0)])
        P = rng.uniform(-0.4, 0.4, size=(3, 3))
        Xc = P @ R.T + t
        f = Xc / np.linalg.norm(Xc, axis=1, keepdims=True)
        sols = O.p3p_poses(f, P)
        assert 1 <= len(sols) <= 4
        assert min(np.abs(Rs - R).max() + np.abs(ts - t).max() for Rs, ts in sols) < 1e-7
    worst_r = worst_t = 0.0
    for trial in range(30):
        rv = rng.normal(size=3)
        rv *= rng.uniform(0.2, 2.5) / np.linalg.norm(rv)
        R = O.rodrigues(rv)
        X = rng.uniform(-0.4, 0.4, size=(8, 3))
        t = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(1.0, 3.0)])
        kp = O.project_points(X, R, t, K, dist) + rng.normal(0, 0.3, size=(8, 2))
        out_idx = rng.choice(8, trial % 3, replace=False)
        kp[out_idx] += rng.uniform(40, 120, size=(len(out_idx), 2)) * rng.choice([-1, 1], size=(len(out_idx), 2))
        rvec, tvec, mask, rms, st = O.pnp_solve(X, kp, K, dist)
        clean = np.ones(8, dtype=bool)
        clean[out_idx] = False
        np.testing.assert_array_equal(mask, clean)
        assert st & 1 and rms < 1.0
        _, r2, t2 = cv2.solvePnP(X[clean], kp[clean], K, dist, rvec=rv.reshape(3, 1).copy(), tvec=t.reshape(3, 1).copy(),
                                 useExtrinsicGuess=True, flags=cv2.SOLVEPNP_ITERATIVE)
        dR = O.rodrigues(rvec) @ O.rodrigues(r2.reshape(3)).T
        worst_r = max(worst_r, float(np.linalg.norm(O.rvec_from_matrix(dR))))
        worst_t = max(worst_t, float(np.linalg.norm(tvec - t2.reshape(3))))
        ok, r3, t3, inl = cv2.solvePnPRansac(X, kp, K, dist, flags=cv2.SOLVEPNP_EPNP)
        if ok:
            assert np.linalg.norm(tvec - t3.reshape(3)) < 0.05
    assert worst_r < 1e-6 and worst_t < 1e-6, (worst_r, worst_t)
    # the reference's refusals: fewer than 4 confident points -> None
    assert O.pnp_solve(X, kp, K, dist, w=np.array([1, 1, 1, 0, 0, 0, 0, 0.0]), min_weight=0.5) is None


def test_estimate_camera_pose_golden_from_the_reference():
    """The reference's own estimate_camera_pose (FK -> sigmoid/arg-max decode -> confidence filter ->
    cv2.solvePnPRansac EPNP, model/Fr5_model_train.ipynb:4707-4753), run unmodified by make_golden.py on belief maps
    rendered at the projections of a known camera pose. The oracle must reproduce its intermediate results exactly
    (FK points, decoded key-points) and its pose up to the estimators' difference: EPnP + RANSAC has no refinement,
    the oracle ends in an LM optimum; both sit within the 4 px cell quantisation of the true pose, so 1e-2 rad /
    5 mm between them (measured 3.1e-3 rad / 0.95 mm) and 2e-2 rad / 1 cm to the truth."""
    q, (rv_true, tv_true) = INP["pose_q"], INP["pose_rt"]
    K, dist = G["zedx_K"][0], G["zedx_dist"][0]
    obj = O.fk_fr5(q, "top")
    np.testing.assert_allclose(obj, G["pose_ok_obj"], rtol=0, atol=6e-8)
    uv = O.project_points(obj.astype(np.float64), O.rodrigues(rv_true), tv_true, K, dist)
    np.testing.assert_allclose(uv, G["pose_uv_true"], atol=1e-6)
    for name, low in (("ok", (6,)), ("refused", (1, 3, 5, 6))):
        maps = _mg.pose_case_maps(G["pose_uv_true"], low)
        d = O.decode(maps, _mg.POSE_IMAGE_HW[1] / maps.shape[2], _mg.POSE_IMAGE_HW[0] / maps.shape[1], apply_sigmoid=True)
        np.testing.assert_array_equal(d["kp_hard"], G[f"pose_{name}_img"])          # bit-exact decode
        res = O.pnp_solve(obj.astype(np.float64), d["kp_hard"].astype(np.float64), K, dist, w=d["score"],
                          min_weight=_mg.POSE_THRESHOLD)
        ref = G[f"pose_{name}_rt"]
        if name == "refused":
            assert res is None and np.isnan(ref).all()                              # < 4 confident points: both refuse
            continue
        rvec, tvec = res[0], res[1]

        def rot_diff(a, b):
            return math.acos(min(1.0, max(-1.0, (np.trace(O.rodrigues(a).T @ O.rodrigues(b)) - 1) / 2)))
        assert rot_diff(rvec, ref[0]) < 1e-2 and np.abs(tvec - ref[1]).max() < 5e-3
        assert rot_diff(rvec, rv_true) < 2e-2 and np.abs(tvec - tv_true).max() < 1e-2
        assert rot_diff(ref[0], rv_true) < 2e-2 and np.abs(ref[1] - tv_true).max() < 1e-2
