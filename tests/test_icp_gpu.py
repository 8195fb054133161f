"""Parity of the CUDA ICP (through the C-ABI) with the oracle and the golden float64 reference outputs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from onepiece_b200 import registration as reg
from onepiece_b200 import scenes
from oracle import oracleapi

pytestmark = pytest.mark.gpu


def pose_delta(A, B):
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    R = A[:3, :3].T @ B[:3, :3]
    ang = np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return float(np.linalg.norm(A[:3, 3] - B[:3, 3])), float(ang)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "icp_small.npz"))


@pytest.mark.parametrize("mode", ["plane", "point"])
def test_golden_float64_reference(g, mode):
    src, tgt = reg.PointCloud(g["src"]), reg.PointCloud(g["tgt"], g["nrm"] if mode == "plane" else None)
    par = reg.ICPParameter(int(g["max_iter"]), float(g["threshold"]), 1.0)
    r = (reg.PointToPlane if mode == "plane" else reg.PointToPoint)(src, tgt, np.eye(4), par)
    dt, dr = pose_delta(r.T, g[f"{mode}_T64"])
    ft, fr = pose_delta(g[f"{mode}_T32"], g[f"{mode}_T64"])
    # north_star tolerance 1e-5 m / 1e-4 rad; result.T is float32, so allow its quantisation (1e-7 at ~1 m)
    assert dt < 1e-5 and dr < 1e-4, (dt, dr, "float32 reference's own deviation:", ft, fr)
    assert np.array_equal(r.correspondence_set_index, g[f"{mode}_pairs64"])
    assert abs(r.rmse - float(g[f"{mode}_rmse64"])) < 1e-7
    assert r.correspondence_set[0].shape == (len(r.correspondence_set_index), 3)


def float_transform(T, pts):
    """geometry::TransformPoints in float32, as the search does it (w = 1 for a rigid pose)"""
    T = np.asarray(T, np.float32)
    return np.stack([((T[k, 0] * pts[:, 0] + T[k, 1] * pts[:, 1]) + T[k, 2] * pts[:, 2]) + T[k, 3] for k in range(3)], 1).astype(np.float32)


def test_nearest_neighbour_is_exact(g):
    """max_iteration = 1: one search under init_T.  Every neighbour the call keeps must be the exact nearest neighbour the
    oracle's brute force finds under init_T, and a point it reports as 'none' has no neighbour within the threshold (the
    closing inlier test cannot pass for it)."""
    T0 = np.eye(4, dtype=np.float32)
    T0[:3, 3] = [0.01, -0.02, 0.015]
    thr = 0.03
    src = g["src"]
    reg.PointToPoint(reg.PointCloud(src), reg.PointCloud(g["tgt"]), T0, reg.ICPParameter(1, thr, 1.0))
    nn = reg.last_nn(len(src))
    assert np.array_equal(reg.last_prev_pose(), T0)
    moved = float_transform(T0, src)
    ref = oracleapi.nearest(moved, g["tgt"])
    dist = np.linalg.norm(moved - g["tgt"][ref], axis=1)
    found = nn >= 0
    assert np.array_equal(nn[found], ref[found])
    assert np.all(dist[~found] > thr * 0.999)
    assert found.sum() > 0.5 * len(src)


def test_zero_iterations_is_the_reference_degenerate_result(g):
    """max_iteration = 0: corresponding_index stays -1 (ICP.cpp:58,174), so there are no inliers, rmse = sqrt(0/0) and result.T is
    the Kabsch fit of nothing"""
    r = reg.PointToPoint(reg.PointCloud(g["src"]), reg.PointCloud(g["tgt"]), np.eye(4), reg.ICPParameter(0, 0.05, 1.0))
    assert len(r.correspondence_set_index) == 0 and np.isnan(r.rmse) and np.all(np.isnan(r.T))
    assert np.array_equal(r.T_iterated, np.eye(4, dtype=np.float32))


@pytest.mark.parametrize("mode", ["plane", "point"])
@pytest.mark.parametrize("max_iter", [1, 2, 3])
@pytest.mark.parametrize("persistent", ["1", "0"])
def test_unconverged_registration_matches_the_reference(g, mode, max_iter, persistent, tmp_path):
    """Few iterations from a large offset (DenseSlam calls PointToPoint with max_iteration = 1, DenseSlam.cpp:72-90): the closing
    CountInliers re-tests the neighbours of the last iteration -- found under the previous pose, at whatever distance -- under
    the final pose.  Pairs identical to the oracle's (pinned to the compiled reference for exactly these settings in
    tests/test_oracle_icp.py) and to the compiled float64 reference itself where it travelled; both launch forms."""
    import subprocess
    import sys

    from conftest import ROOT
    from oracle import refapi
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from onepiece_b200 import registration as reg\n"
        "g = np.load(sys.argv[2]); mode = sys.argv[3]; it = int(sys.argv[4])\n"
        "T0 = np.eye(4); T0[:3, 3] = [0.03, -0.01, 0.015]\n"
        "tgt = reg.PointCloud(g['tgt'], g['nrm'] if mode == 'plane' else None)\n"
        "f = reg.PointToPlane if mode == 'plane' else reg.PointToPoint\n"
        "r = f(reg.PointCloud(g['src']), tgt, T0, reg.ICPParameter(it, 0.04, 1.0))\n"
        "np.savez(sys.argv[1], T=r.T, Ti=r.T_iterated, pairs=r.correspondence_set_index, rmse=r.rmse)\n" % ROOT)
    path = str(tmp_path / "out.npz")
    env = dict(os.environ, OPB_ICP_PERSISTENT=persistent)
    subprocess.run([sys.executable, "-c", code, path, os.path.join(GOLDEN, "icp_small.npz"), mode, str(max_iter)], check=True, env=env,
                   timeout=300)
    r = np.load(path)
    nrm = g["nrm"] if mode == "plane" else None
    T0 = np.eye(4)
    T0[:3, 3] = [0.03, -0.01, 0.015]
    o = oracleapi.icp(g["src"], g["tgt"], nrm, T0, max_iter, 0.04)
    assert np.array_equal(r["pairs"], o["pairs"])
    dt, dr = pose_delta(r["T"], o["T"])
    assert dt < 1e-5 and dr < 1e-4, (dt, dr)
    assert abs(float(r["rmse"]) - o["rmse"]) < 1e-6
    if refapi.available("f64"):
        f = refapi.icp(g["src"], g["tgt"], nrm, T0, max_iter, 0.04, "f64")
        assert np.array_equal(r["pairs"], f["pairs"])
        dt, dr = pose_delta(r["T"], f["T"])
        assert dt < 1e-5 and dr < 1e-4, (dt, dr)


@pytest.mark.parametrize("mode", ["plane", "point"])
def test_vs_oracle_with_initial_guess_and_clutter(g, mode):
    rng = np.random.default_rng(3)
    src = g["src"].copy()
    src[::50] += rng.normal(0, 0.5, src[::50].shape).astype(np.float32)  # gross outliers
    src[7] = np.nan  # a non-finite point never matches
    T0 = scenes.se3_exp([0.004, -0.003, 0.002, 0.002, -0.001, 0.003]).astype(np.float32)
    nrm = g["nrm"] if mode == "plane" else None
    par = reg.ICPParameter(8, 0.04, 1.0)
    r = (reg.PointToPlane if mode == "plane" else reg.PointToPoint)(reg.PointCloud(src), reg.PointCloud(g["tgt"], nrm), T0, par)
    o = oracleapi.icp(src, g["tgt"], nrm, T0, 8, 0.04)
    dt, dr = pose_delta(r.T, o["T"])
    assert dt < 1e-5 and dr < 1e-4, (dt, dr)
    dti, dri = pose_delta(r.T_iterated, o["T_iterated"])
    assert dti < 1e-5 and dri < 1e-4
    a, b = r.correspondence_set_index, o["pairs"]
    # pairs may differ only where a float32-level difference of the iterated pose flips a near-tie
    common = len(set(map(tuple, a)) & set(map(tuple, b)))
    assert common >= 0.999 * max(len(a), len(b))
    assert abs(r.rmse - o["rmse"]) < 1e-6


def test_reference_error_paths():
    rng = np.random.default_rng(4)
    t = rng.normal(0, 1, (2000, 3)).astype(np.float32)
    s = (t + 0.01).astype(np.float32)
    n = np.tile(np.array([0, 0, 1], np.float32), (2000, 1))
    # target without normals / scaling != 1: the reference prints an error and returns a default result
    r = reg.PointToPlane(reg.PointCloud(s), reg.PointCloud(t), np.eye(4), reg.ICPParameter(3, 0.5, 1.0))
    assert not r.ok and len(r.correspondence_set_index) == 0
    r = reg.PointToPlane(reg.PointCloud(s), reg.PointCloud(t, n), np.eye(4), reg.ICPParameter(3, 0.5, 2.0))
    assert not r.ok
    # PointToPoint with scaling solves the same problem in scaled units
    a = reg.PointToPoint(reg.PointCloud(s), reg.PointCloud(t), np.eye(4), reg.ICPParameter(3, 0.5, 1.0))
    b = reg.PointToPoint(reg.PointCloud(s), reg.PointCloud(t), np.eye(4), reg.ICPParameter(3, 1.0, 2.0))
    o = oracleapi.icp(s, t, None, np.eye(4), 3, 1.0, scaling=2.0)
    assert pose_delta(a.T, b.T)[0] < 1e-5 and pose_delta(b.T, o["T"])[0] < 1e-5
    # clouds farther apart than the threshold: no inliers, result.T is NaN like the reference's 0/0
    far = reg.PointToPoint(reg.PointCloud(s + 100), reg.PointCloud(t), np.eye(4), reg.ICPParameter(2, 0.05, 1.0))
    assert len(far.correspondence_set_index) == 0 and np.all(np.isnan(far.T))


def test_full_size_frame_pair_recovers_the_motion():
    """640x480 (307,200 points), the bench's ICP workload: deterministic, and the recovered motion agrees with the
    analytic camera motion to sensor quantisation."""
    cam = scenes.Camera()
    d0, _, T0, n0 = scenes.room(cam, 0, with_normals=True)
    d1, _, T1 = scenes.room(cam, 1)
    tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)
    nrm = n0.reshape(-1, 3)
    par = reg.ICPParameter(30, 0.05, 1.0)
    r1 = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), par)
    r2 = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), par)
    assert np.array_equal(r1.T, r2.T) and np.array_equal(r1.correspondence_set_index, r2.correspondence_set_index)
    truth = np.linalg.inv(T0.astype(np.float64)) @ T1.astype(np.float64)
    dt, dr = pose_delta(r1.T_iterated, truth)
    assert dt < 2e-3 and dr < 2e-3
    assert len(r1.correspondence_set_index) > 0.99 * len(src)


@pytest.mark.parametrize("iters", [1, 2, 3, 6, 12])
def test_certified_neighbours_are_the_exact_ones(iters):
    """After a few iterations most queries keep the neighbour of an earlier full search because they provably moved less than
    that search's budget (icp_certify_kernel).  The neighbours of the LAST ITERATION -- certified or searched -- must be the exact
    nearest neighbours the oracle's brute force finds for that iteration's pose."""
    import ctypes as C

    from onepiece_b200 import capi
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 2, c0.fy / 2, c0.cx / 2, c0.cy / 2, 320, 240, 1000.0)
    d0, _, _, n0 = scenes.room(cam, 0, with_normals=True)
    d1, _, _ = scenes.room(cam, 4)
    tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)
    nrm = np.ascontiguousarray(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)])
    thr = 0.05
    r = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), reg.ICPParameter(iters, thr, 1.0))
    nn = reg.last_nn(len(src))
    n_search = C.c_uint64(0)
    capi.check(capi.lib.opb_icp_last_search_count(reg._Workspace.get(0), C.byref(n_search)))
    total = len(src) * (iters + 1)
    assert 0 < n_search.value <= total
    if iters >= 6:
        assert n_search.value < 0.6 * total, (n_search.value, total)   # the converged passes are answered from certificates
    # the neighbours the call ends with are those of the LAST ITERATION's search, under the pose before its update
    moved = float_transform(reg.last_prev_pose(), src)
    ref = oracleapi.nearest(moved, tgt)
    dist = np.linalg.norm(moved.astype(np.float64) - tgt[ref], axis=1)
    found = nn >= 0
    assert np.array_equal(nn[found], ref[found])
    assert np.all(dist[~found] > thr * 0.999)   # 'none' = nothing that could pass the closing inlier test
    assert found.mean() > 0.9


def test_certification_does_not_change_the_result(tmp_path):
    """The same registration with certificates switched off (every pass searches every point): identical pairs and pose bits."""
    import subprocess
    import sys

    from conftest import ROOT
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from onepiece_b200 import registration as reg, scenes\n"
        "cam = scenes.Camera()\n"
        "d0, _, _, n0 = scenes.room(cam, 0, with_normals=True); d1, _, _ = scenes.room(cam, 2)\n"
        "tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)\n"
        "nrm = np.ascontiguousarray(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)])\n"
        "r = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), reg.ICPParameter(30, 0.05, 1.0))\n"
        "np.savez(sys.argv[1], T=r.T, Ti=r.T_iterated, pairs=r.correspondence_set_index, rmse=r.rmse)\n" % ROOT)
    out = {}
    for flag in ("1", "0"):
        path = str(tmp_path / f"icp_{flag}.npz")
        env = dict(os.environ, OPB_ICP_CERTIFY=flag)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
        out[flag] = np.load(path)
    assert np.array_equal(out["1"]["pairs"], out["0"]["pairs"])
    assert np.array_equal(out["1"]["Ti"].view(np.uint32), out["0"]["Ti"].view(np.uint32))
    assert np.array_equal(out["1"]["T"].view(np.uint32), out["0"]["T"].view(np.uint32))
    assert float(out["1"]["rmse"]) == float(out["0"]["rmse"])


def test_estimate_normals_matches_the_oracle_bit_for_bit():
    """PointCloud::EstimateNormals on the device: the reference's own k-d tree walk (same tree, same visiting order) + FitPlane +
    the JacobiSVD restatement, so even the tie-ridden raw cloud must agree bit for bit with the oracle (itself bit-identical to
    the compiled reference there).  The uniform-grid search (developer knob) orders ties by index: identical on tie-free clouds."""
    from onepiece_b200 import capi
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0)
    d, _, _, n_true = scenes.room(cam, 0, with_normals=True)
    pts = scenes.backproject(d, cam)
    jit = (pts + np.random.default_rng(0).normal(0, 1e-4, pts.shape)).astype(np.float32)
    for cloud, what in ((jit, "tie-free cloud"), (pts, "raw cloud")):
        pc = reg.PointCloud(cloud)
        pc.EstimateNormals()
        ref = oracleapi.estimate_normals(cloud)
        same = (pc.normals.view(np.uint32) == ref.view(np.uint32)).all(1)
        assert same.all(), f"{what}: {np.count_nonzero(~same)} of {len(cloud)} normals differ, first at {np.argmax(~same)}"
    os.environ["OPB_NORMALS_KNN"] = "grid"
    try:
        pg = reg.PointCloud(jit)
        pg.EstimateNormals()
        assert np.array_equal(pg.normals.view(np.uint32), oracleapi.estimate_normals(jit).view(np.uint32)), "grid search, tie-free cloud"
    finally:
        del os.environ["OPB_NORMALS_KNN"]
    # the estimated normals are the surface normals up to sign (away from the room's edges)
    nt = n_true.reshape(-1, 3)[(d > 0).reshape(-1)]
    assert np.median(np.abs((pc.normals * nt).sum(1))) > 0.999
    # other parameters, and the reference's corner cases
    pc = reg.PointCloud(jit)
    pc.EstimateNormals(0.0004, 12)                 # squared-distance cut at 2 cm
    assert np.array_equal(pc.normals.view(np.uint32), oracleapi.estimate_normals(jit, 0.0004, 12).view(np.uint32))
    far = reg.PointCloud(np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [10, 10, 0]], np.float32))
    far.EstimateNormals()
    assert not far.normals.any()                   # fewer than three points in range: FitPlane returns the zero vector
    with pytest.raises(capi.OpbError):
        pc.EstimateNormals(0.1, 65)


def test_exactly_equidistant_targets_resolve_to_a_nearest_neighbour_documented_tie_rule():
    """Tie rule (documented deviation): among EXACTLY equidistant targets the device search keeps the lowest
    target index, nanoflann keeps the first one its depth-first walk meets (KDTree.h:177-196, KNNResultSet::addPoint), which depends
    on the tree and on the side of each cut the query lies on.  On a lattice built to tie everywhere -- targets on a 0.25 m grid,
    sources at the cell centres: eight exactly equidistant targets per query -- both must return A nearest neighbour (the same
    distance bit for bit, every query an inlier), and the device's choice must be the lowest index among the ties."""
    from oracle import refapi
    g = np.arange(6, dtype=np.float32) * np.float32(0.25)
    tgt = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    c = (np.arange(5, dtype=np.float32) * np.float32(0.25) + np.float32(0.125)).astype(np.float32)
    src = np.stack(np.meshgrid(c, c, c, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(3)
    tgt = tgt[rng.permutation(len(tgt))]            # index order unrelated to position
    par = reg.ICPParameter(1, 0.5, 1.0)
    r = reg.PointToPoint(reg.PointCloud(src), reg.PointCloud(tgt), np.eye(4), par)
    pairs = r.correspondence_set_index
    assert len(pairs) == len(src) and np.array_equal(pairs[:, 0], np.arange(len(src)))
    d_all = ((src[:, None, :] - tgt[None, :, :]) ** 2).sum(-1)     # exact in float32: every term is a multiple of 2^-6
    d_min = d_all.min(1)
    assert (np.isclose(d_all, d_min[:, None], rtol=0, atol=0).sum(1) == 8).all(), "the lattice must tie eight ways"
    mine = d_all[np.arange(len(src)), pairs[:, 1]]
    assert np.array_equal(mine, d_min), "not a nearest neighbour"
    assert np.array_equal(pairs[:, 1], np.argmax(d_all == d_min[:, None], 1)), "the device keeps the lowest index among exact ties"
    if refapi.available("f32"):
        ref = refapi.icp(src, tgt, None, np.eye(4), 1, 0.5, "f32")["pairs"]
        assert len(ref) == len(src)
        assert np.array_equal(d_all[np.arange(len(src)), ref[:, 1]], d_min), "the reference's choice is a nearest neighbour too"
        differ = int((ref[:, 1] != pairs[:, 1]).sum())
        print(f"exact 8-way ties: the reference's walk order and the lowest-index rule pick different (equidistant) targets for {differ} of {len(src)} queries")


def test_pair_list_off_the_critical_path(g):
    """opb_icp_set_async_pairs: the registration returns with pose and counters, the pair list lands in the caller's page-locked
    buffer behind it and is complete after opb_icp_wait_pairs -- the same list, the same result, as the synchronous call; a
    pageable buffer is filled inside the call as before."""
    import ctypes as C

    from onepiece_b200 import capi
    src, tgt, nrm = np.ascontiguousarray(g["src"], np.float32), np.ascontiguousarray(g["tgt"], np.float32), np.ascontiguousarray(g["nrm"], np.float32)
    want = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), reg.ICPParameter(6, 0.04, 1.0))
    ws = C.c_void_p()
    capi.check(capi.lib.opb_icp_create(0, None, C.byref(ws)))
    capi.check(capi.lib.opb_icp_set_async_pairs(ws, 1))
    pinned = C.c_void_p()
    capi.check(capi.lib.opb_host_alloc(C.byref(pinned), len(src) * 8))
    try:
        par, res = capi.IcpParams(6, 0.04, 1.0), capi.IcpResult()
        T0 = np.ascontiguousarray(np.eye(4, dtype=np.float32)).reshape(16)
        view = np.ctypeslib.as_array(C.cast(pinned, C.POINTER(C.c_int32)), shape=(len(src), 2))
        for rep in range(3):  # back-to-back calls: the next registration waits for the previous list to have left the device buffer
            view[:] = -7
            capi.check(capi.lib.opb_icp_point_to_plane(ws, src.ctypes.data, len(src), tgt.ctypes.data, nrm.ctypes.data, len(tgt), T0.ctypes.data,
                                                       C.byref(par), C.byref(res), pinned, len(src)))
            assert np.array_equal(np.array(res.T[:], np.float32).reshape(4, 4).T, want.T) and res.rmse == want.rmse
            assert res.n_local_pairs == len(want.correspondence_set_index)
            if rep < 2:
                capi.check(capi.lib.opb_icp_wait_pairs(ws))
                assert np.array_equal(view[: res.n_local_pairs], want.correspondence_set_index)
        capi.check(capi.lib.opb_icp_wait_pairs(ws))
        assert np.array_equal(view[: res.n_local_pairs], want.correspondence_set_index)
        pageable = np.full((len(src), 2), -7, np.int32)
        capi.check(capi.lib.opb_icp_point_to_plane(ws, src.ctypes.data, len(src), tgt.ctypes.data, nrm.ctypes.data, len(tgt), T0.ctypes.data,
                                                   C.byref(par), C.byref(res), pageable.ctypes.data, len(src)))
        assert np.array_equal(pageable[: res.n_local_pairs], want.correspondence_set_index)  # complete on return
    finally:
        capi.lib.opb_host_free(pinned)
        capi.lib.opb_icp_destroy(ws)
