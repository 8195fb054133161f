"""The C oracle's dense RGB-D odometry restatement (oracle/opb_oracle.c) against the golden fixture generated from the
compiled reference (tests/golden/gen_golden.py --odometry) and, where oracle/_ref is present, the reference directly.

What is pinned, and how:
  * pre-processing filters: OpenCV is an un-vendored dependency of the reference; the restated filters are compared with
    the images real OpenCV (cv2 4.13) produced at fixture time -- identical NaN pattern, <= 1e-6 of the image's range;
  * correspondences: bit-identical lists against the float32 reference build, teacher-forced poses;
  * one solver iteration: next pose within 1e-6 of the float64 reference build (the float32 build sums its 6x6 system
    sequentially in float32 and is itself ~1e-5 away per iteration);
  * whole coarse-to-fine run: the reference is chaotic end to end (its order-dependent correspondence filter makes the
    float32 and float64 builds diverge); the oracle must stay within that band.
CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, sha
from onepiece_b200 import scenes
from oracle import oracleapi, refapi
from test_oracle_icp import pose_delta


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "odometry_small.npz"))


def camera(g):
    c = g["cam"]
    return scenes.Camera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5]), float(c[6]))


@pytest.fixture(scope="module")
def frames(g):
    cam = camera(g)
    S = oracleapi.OracleFrame(g["src_bgr"], g["src_depth"]).preprocess(cam.depth_scale)
    T = oracleapi.OracleFrame(g["tgt_bgr"], g["tgt_depth"]).preprocess(cam.depth_scale)
    return cam, S, T


def test_filters_match_real_opencv(g, frames):
    _, _, T = frames
    for a in range(6):
        for l in range(3):
            mine, cv = T.image(a, l), g[f"cv2_tgt_{a}_{l}"]
            assert mine.shape == cv.shape
            assert np.array_equal(np.isnan(mine), np.isnan(cv)), f"NaN pattern of image {a} level {l}"
            assert np.nanmax(np.abs(mine - cv)) <= 1e-6 * max(1.0, np.nanmax(np.abs(cv))), (a, l)
    # gray conversion is integer arithmetic: exact
    import importlib.util
    if importlib.util.find_spec("cv2"):
        import cv2
        rng = np.random.default_rng(0)
        c = rng.integers(0, 256, (36, 52, 3), dtype=np.uint8)
        assert np.array_equal(oracleapi.gray_u8(c), cv2.cvtColor(c, cv2.COLOR_RGB2GRAY))
        f = rng.random((36, 52)).astype(np.float32)
        assert np.array_equal(oracleapi.blur3(f), cv2.GaussianBlur(f, (3, 3), 0))
        assert np.abs(oracleapi.pyr_down(f) - cv2.pyrDown(f)).max() < 5e-7
        assert np.abs(oracleapi.sobel3(f, 1) - cv2.Sobel(f, cv2.CV_32F, 1, 0)).max() < 1e-6
        assert np.abs(oracleapi.sobel3(f, 0) - cv2.Sobel(f, cv2.CV_32F, 0, 1)).max() < 1e-6


def test_preprocessed_images_are_pinned(g, frames):
    _, S, T = frames
    for a in range(6):
        for l in range(3):
            assert sha(S.image(a, l)) == str(g[f"sha_src_{a}_{l}"])
            assert sha(T.image(a, l)) == str(g[f"sha_tgt_{a}_{l}"])


def test_teacher_forced_iterations_match_reference(g, frames):
    cam, S, T = frames
    for k, (level, term) in enumerate(g["tf_cases"]):
        o = oracleapi.single_iteration(S, T, cam, int(level), g[f"tf{k}_T0"], int(term))
        assert np.array_equal(o["pairs"], g[f"tf{k}_pairs32"].astype(np.uint32)), f"correspondences, level {level} term {term}"
        if not bool(g[f"tf{k}_pairs64_equal"]):
            continue  # the float64 build picked other pixels: its pose is not comparable
        dt, dr = pose_delta(o["T"], g[f"tf{k}_T64"])
        ft, fr = pose_delta(g[f"tf{k}_T32"], g[f"tf{k}_T64"])
        assert dt < 1e-6 and dr < 1e-6, (level, term, dt, dr, "float32 reference:", ft, fr)
        if term != 2:
            J = g[f"tf{k}_JTJ64"]
            assert np.abs(o["JTJ"] - J).max() <= 1e-6 * np.abs(J).max()
            # J^T r is a sum of signed terms (cancellation), and the float64 build evaluates the rows in mixed precision
            assert np.abs(o["JTr"] - g[f"tf{k}_JTr64"]).max() <= 1e-5 * max(1.0, np.abs(g[f"tf{k}_JTr64"]).max())


def test_whole_tracking_run_stays_inside_the_references_own_band(g):
    cam = camera(g)
    for term in (0, 1, 2):
        S = oracleapi.OracleFrame(g["src_bgr"], g["src_depth"])
        T = oracleapi.OracleFrame(g["tgt_bgr"], g["tgt_depth"])
        o = oracleapi.dense_tracking_frames(S, T, cam, np.eye(4), term)
        assert sha(S.image(0, 0)) == str(g["norm_src_sha"]) and sha(T.image(0, 0)) == str(g["norm_tgt_sha"]), \
            "NormalizeIntensity (in place on the level-0 gray) differs from the reference"
        T32, T64 = g[f"ms{term}_T_f32"], g[f"ms{term}_T_f64"]
        band_t, band_r = pose_delta(T32, T64)
        dt = min(pose_delta(o["T"], T64)[0], pose_delta(o["T"], T32)[0])
        dr = min(pose_delta(o["T"], T64)[1], pose_delta(o["T"], T32)[1])
        assert dt <= max(2 * band_t, 1e-5) and dr <= max(2 * band_r, 1e-4), (term, dt, dr, band_t, band_r)
        assert o["success"] == bool(g[f"ms{term}_ok_f64"])
        # until the first iteration where the float32 reference itself leaves the float64 one, counts are identical
        c32, c64 = g[f"ms{term}_corr_f32"], g[f"ms{term}_corr_f64"]
        n = min(len(c32), len(c64), len(o["corr_per_iteration"]))
        same = np.flatnonzero(c32[:n] != c64[:n])
        upto = int(same[0]) if len(same) else n
        # (a pixel on a rounding boundary may still flip: poses agree to ~1e-7 per iteration, not bitwise)
        assert np.all(np.abs(o["corr_per_iteration"][:upto] - c64[:upto]) <= np.maximum(2, 0.002 * c64[:upto]))
        assert np.array_equal(o["corr_per_iteration"][:8], c64[:8])


def test_correspondence_filter_is_order_dependent_like_the_reference():
    """AddElementToCorrespondenceMap reads the warped depth at the TARGET pixel but writes at the SOURCE pixel
    (DenseOdometryFunction.cpp:8-25): a one-pixel shift makes acceptance depend on raster order."""
    cam = scenes.Camera(100.0, 100.0, 8.0, 6.0, 16, 12, 1000.0)
    d = np.full((12, 16), 1.0, np.float32)
    d[:, ::2] = 1.01  # alternating depth so `existing > transformed` alternates
    T = np.eye(4)
    T[0, 3] = -0.01  # shifts every pixel one column to the left: target pixel = the previous source pixel
    p = oracleapi.correspondences(d, d, cam, T)
    assert 0 < len(p) < d.size - 12
    if refapi.available("f32"):
        assert np.array_equal(p, refapi.correspondences(d, d, cam, T, "f32"))


@pytest.mark.skipif(not refapi.available("f64"), reason="oracle/_ref not built")
def test_oracle_vs_compiled_reference_full_resolution():
    cam = scenes.Camera()
    d0, c0, _ = scenes.room(cam, 0)
    d1, c1, _ = scenes.room(cam, 1)
    S = oracleapi.OracleFrame(c1, d1).preprocess(cam.depth_scale)
    T = oracleapi.OracleFrame(c0, d0).preprocess(cam.depth_scale)
    si, ti = S.images(), T.images()
    T0 = scenes.se3_exp([1e-3, -2e-3, 1e-3, 2e-3, 1e-3, -1e-3]).astype(np.float32)
    for level in (1, 0):
        o = oracleapi.single_iteration(S, T, cam, level, T0, 0)
        r32 = refapi.single_iteration(si, ti, cam, level, T0, 0, "f32")
        assert np.array_equal(o["pairs"], r32["pairs"])
        r64 = refapi.single_iteration(si, ti, cam, level, T0, 0, "f64")
        if np.array_equal(r64["pairs"], r32["pairs"]):
            dt, dr = pose_delta(o["T"], r64["T"])
            assert dt < 1e-6 and dr < 1e-6
    # NormalizeIntensity: the float32 sequential mean of the reference, bit for bit
    n0 = refapi.correspondences(si[1][0], ti[1][0], cam, np.eye(4), "f32")
    sg, tg = refapi.normalize_intensity(si[0][0], ti[0][0], n0, "f32")
    oracleapi.dense_tracking_frames(S, T, cam, np.eye(4), 0)
    assert np.array_equal(S.image(0, 0), sg) and np.array_equal(T.image(0, 0), tg)
