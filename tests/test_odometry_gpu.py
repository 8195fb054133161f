"""Parity of the CUDA dense RGB-D odometry (through the C-ABI) with the oracle and the golden reference outputs.

Gates: pre-processed images bit-identical to the oracle's; correspondence lists bit-identical to the float32
reference build's (teacher-forced poses); one iteration's pose within 1e-6 of the float64 reference build; the whole
coarse-to-fine run identical to the oracle's iteration by iteration (same correspondence counts, poses within 1e-6 m /
1e-6 rad, identical final pair list) -- the reference itself is chaotic end to end, its float32 and float64 builds differ
by more than the north_star tolerance (printed)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_bit_equal
from onepiece_b200 import capi, scenes
from onepiece_b200.odometry import Odometry
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
    return np.load(os.path.join(GOLDEN, "odometry_small.npz"))


def camera(g):
    c = g["cam"]
    return scenes.Camera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5]), float(c[6]))


@pytest.mark.parametrize("depth_f32", [False, True])
def test_preprocessed_images_are_bit_identical_to_the_oracle(g, depth_f32):
    cam = camera(g)
    d = g["tgt_depth"]
    if depth_f32:
        d = (d.astype(np.float32) / np.float32(1000.0)).astype(np.float32)
        d[::7, ::5] = 0.0
        d[3, 3] = np.inf
    odo = Odometry(cam)
    f = odo.Frame(g["tgt_bgr"], d).preprocess()
    o = oracleapi.OracleFrame(g["tgt_bgr"], d).preprocess(cam.depth_scale)
    assert f.IsPreprocessedDense()
    for a in range(6):
        for l in range(3):
            assert_bit_equal(f.image(a, l), o.image(a, l), f"image {a} level {l}", nan_payload=False)


def test_teacher_forced_iterations(g):
    cam = camera(g)
    odo = Odometry(cam)
    S, T = odo.Frame(g["src_bgr"], g["src_depth"]), odo.Frame(g["tgt_bgr"], g["tgt_depth"])
    oS = oracleapi.OracleFrame(g["src_bgr"], g["src_depth"]).preprocess(cam.depth_scale)
    oT = oracleapi.OracleFrame(g["tgt_bgr"], g["tgt_depth"]).preprocess(cam.depth_scale)
    for k, (level, term) in enumerate(g["tf_cases"]):
        r = odo.single_iteration(S, T, int(level), g[f"tf{k}_T0"], int(term))
        assert np.array_equal(r["pairs"], g[f"tf{k}_pairs32"].astype(np.uint32)), f"correspondences, level {level} term {term}"
        o = oracleapi.single_iteration(oS, oT, cam, int(level), g[f"tf{k}_T0"], int(term))
        scale = np.abs(o["JTJ"]).max()
        assert np.abs(r["JTJ"] - o["JTJ"]).max() <= 1e-12 * scale
        assert np.abs(r["JTr"] - o["JTr"]).max() <= 1e-12 * max(1.0, np.abs(o["JTr"]).max()) * 1e2
        assert abs(r["r2"] - o["r2"]) <= 1e-12 * max(1.0, o["r2"])
        assert np.abs(r["T"] - o["T"]).max() <= 1.2e-7, "pose after the iteration vs the oracle"
        if bool(g[f"tf{k}_pairs64_equal"]):
            dt, dr = pose_delta(r["T"], g[f"tf{k}_T64"])
            ft, fr = pose_delta(g[f"tf{k}_T32"], g[f"tf{k}_T64"])
            assert dt < 1e-6 and dr < 1e-6, (level, term, dt, dr, "float32 reference:", ft, fr)


def _compare_runs(r, o, what):
    n = min(len(o["corr_per_iteration"]), 64)
    assert r.iterations == len(o["corr_per_iteration"]), what
    assert np.array_equal(r.corr_per_iteration[:n], o["corr_per_iteration"][:n]), what
    for i in range(n):
        dt, dr = pose_delta(r.T_per_iteration[i], o["T_per_iteration"][i])
        assert dt < 1e-6 and dr < 1e-6, (what, "iteration", i, dt, dr)
    assert np.array_equal(r.pixel_correspondence_set, o["pairs"]), what
    dt, dr = pose_delta(r.T, o["T"])
    assert dt < 1e-6 and dr < 1e-6, (what, dt, dr)
    assert r.tracking_success == o["success"]
    assert abs(r.rmse - o["rmse"]) <= 1e-7 * max(1.0, o["rmse"]), (what, r.rmse, o["rmse"])


@pytest.mark.parametrize("term", [0, 1, 2])
def test_whole_run_matches_the_oracle_and_the_golden_reference(g, term):
    cam = camera(g)
    odo = Odometry(cam)
    S, T = odo.Frame(g["src_bgr"], g["src_depth"]), odo.Frame(g["tgt_bgr"], g["tgt_depth"])
    r = odo.DenseTracking(S, T, np.eye(4), term)
    oS, oT = oracleapi.OracleFrame(g["src_bgr"], g["src_depth"]), oracleapi.OracleFrame(g["tgt_bgr"], g["tgt_depth"])
    o = oracleapi.dense_tracking_frames(oS, oT, cam, np.eye(4), term)
    # NormalizeIntensity rescaled the level-0 gray images in place, with the reference's sequential float32 mean
    assert_bit_equal(S.image(0, 0), oS.image(0, 0), "normalised source gray")
    assert_bit_equal(T.image(0, 0), oT.image(0, 0), "normalised target gray")
    _compare_runs(r, o, f"term {term}")
    # against the compiled reference's own runs: inside the band its float32 and float64 builds span
    T32, T64 = g[f"ms{term}_T_f32"], g[f"ms{term}_T_f64"]
    band_t, band_r = pose_delta(T32, T64)
    dt = min(pose_delta(r.T, T64)[0], pose_delta(r.T, T32)[0])
    dr = min(pose_delta(r.T, T64)[1], pose_delta(r.T, T32)[1])
    print(f"term {term}: GPU vs reference {dt:.2e} m {dr:.2e} rad; reference float32 vs float64 {band_t:.2e} m {band_r:.2e} rad")
    assert dt <= max(2 * band_t, 1e-5) and dr <= max(2 * band_r, 1e-4)
    # correspondence_set pairs xyz_s[v_s][u_s] with xyz_t[v_s][u_s] (reference quirk)
    assert r.correspondence_set.shape == (len(r.pixel_correspondence_set), 2, 3)
    v, u = r.pixel_correspondence_set[0, 0], r.pixel_correspondence_set[0, 1]
    z = oS.image(1, 0)[v, u]
    assert r.correspondence_set[0, 0, 2] == z


def test_frame_reuse_renormalises_like_the_reference(g):
    """A frame that is target in one call and source in the next is rescaled twice (Odometry.cpp:571-595)."""
    cam = camera(g)
    imgs = [scenes.room(cam, k) for k in (0, 2, 4)]
    odo = Odometry(cam)
    F = [odo.Frame(c, d, k) for k, (d, c, _) in enumerate(imgs)]
    O = [oracleapi.OracleFrame(c, d) for d, c, _ in imgs]
    for k in (1, 2):
        r = odo.DenseTracking(F[k], F[k - 1], np.eye(4), 0)
        o = oracleapi.dense_tracking_frames(O[k], O[k - 1], cam, np.eye(4), 0)
        _compare_runs(r, o, f"pair {k}")
    for k in range(3):
        assert_bit_equal(F[k].image(0, 0), O[k].image(0, 0), f"gray of frame {k} after the chain")
        assert_bit_equal(F[k].image(0, 1), O[k].image(0, 1), f"level-1 gray of frame {k} is not rescaled")


def test_mat_overload_normalises_before_the_pyramids(g):
    cam = camera(g)
    odo = Odometry(cam)
    r = odo.DenseTracking(g["src_bgr"], g["tgt_bgr"], g["src_depth"], g["tgt_depth"], np.eye(4), 0)
    o = oracleapi.dense_tracking(g["src_bgr"], g["tgt_bgr"], g["src_depth"], g["tgt_depth"], cam, np.eye(4), 0)
    _compare_runs(r, o, "cv::Mat overload")
    S, T = odo.Frame(g["src_bgr"], g["src_depth"]), odo.Frame(g["tgt_bgr"], g["tgt_depth"])
    r2 = odo.DenseTracking(S, T, np.eye(4), 0)
    assert not np.array_equal(r.T, r2.T), "the two overloads differ in the reference (normalisation order)"


def test_full_resolution_run_and_initial_pose():
    cam = scenes.Camera()
    d0, c0, _ = scenes.room(cam, 0)
    d1, c1, _ = scenes.room(cam, 2)
    T0 = scenes.se3_exp([1e-3, 0, -1e-3, 0, 1e-3, 0]).astype(np.float32)
    odo = Odometry(cam)
    o = oracleapi.dense_tracking_frames(oracleapi.OracleFrame(c1, d1), oracleapi.OracleFrame(c0, d0), cam, T0, 0)
    # every way of launching the solver loop sums the same exact products in double (only the order differs): same counts, same
    # pair lists, poses within 1e-6 at every iteration
    for form in (1, 2, 0):
        odo.SetLoopForm(form)
        r = odo.DenseTracking(odo.Frame(c1, d1), odo.Frame(c0, d0), T0, 0)
        _compare_runs(r, o, f"640x480, loop form {form}")
    odo.SetLoopForm(-1)
    assert r.tracking_success and len(r.pixel_correspondence_set) > 0.3 * 640 * 480


def test_set_multi_scale_and_early_exit():
    cam = scenes.Camera()
    c0 = scenes.Camera(cam.fx / 4, cam.fy / 4, cam.cx / 4, cam.cy / 4, 160, 120, 1000.0)
    d, c, _ = scenes.room(c0, 0)
    odo = Odometry(c0)
    odo.SetMultiScale(1)  # {4}
    # identical frames: every pixel corresponds to itself, the ratio exceeds 0.9 and the level ends after ONE iteration
    r = odo.DenseTracking(odo.Frame(c, d), odo.Frame(c, d), np.eye(4), 0)
    assert r.iterations == 1 and r.corr_per_iteration[0] > 0.9 * 160 * 120
    odo.SetMultiScale(2)  # {4, 4}: level 1 can never reach the full-resolution ratio, level 0 exits after one iteration
    r = odo.DenseTracking(odo.Frame(c, d), odo.Frame(c, d), np.eye(4), 0)
    assert r.iterations == 5


def test_degenerate_inputs_and_error_paths():
    c0 = scenes.Camera(128.0, 128.0, 80.0, 60.0, 160, 120, 1000.0)
    odo = Odometry(c0)
    d = np.zeros((120, 160), np.uint16)
    c = np.zeros((120, 160, 3), np.uint8)
    r = odo.DenseTracking(odo.Frame(c, d), odo.Frame(c, d), np.eye(4), 0)  # no valid depth at all
    assert not r.tracking_success and len(r.pixel_correspondence_set) == 0
    with pytest.raises(capi.OpbError) as e:
        odo.Frame(c, d.astype(np.int32))
    assert e.value.code == capi.OPB_ERR_UNSUPPORTED and "Unknown depth image type" in str(e.value)
    f = odo.Frame(c, d)
    with pytest.raises(capi.OpbError):
        odo.DenseTracking(f, f, np.eye(4), 7)
    other = Odometry(c0)
    with pytest.raises(capi.OpbError):
        other.DenseTracking(f, f, np.eye(4), 0)
