"""The C oracle's ICP restatement against the golden fixture (float64 reference outputs) and, where oracle/_ref
is present, against the compiled reference directly.  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracleapi, refapi


def pose_delta(A, B):
    """(translation difference [m], rotation angle of A^T B [rad])"""
    R = A[:3, :3].T @ B[:3, :3]
    ang = np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return float(np.linalg.norm(A[:3, 3] - B[:3, 3])), float(ang)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "icp_small.npz"))


@pytest.mark.parametrize("mode", ["plane", "point"])
def test_oracle_icp_matches_golden_float64_reference(g, mode):
    nrm = g["nrm"] if mode == "plane" else None
    o = oracleapi.icp(g["src"], g["tgt"], nrm, np.eye(4), int(g["max_iter"]), float(g["threshold"]))
    dt, dr = pose_delta(o["T"], g[f"{mode}_T64"])
    # tolerance of BASELINE.json's north_star: 1e-5 m / 1e-4 rad; the float32 reference itself is further away
    ft, fr = pose_delta(g[f"{mode}_T32"], g[f"{mode}_T64"])
    assert dt < 1e-5 and dr < 1e-4, (dt, dr)
    assert dt <= ft + 1e-9, f"oracle {dt} m should be at least as close to float64 as the float32 reference ({ft} m)"
    assert np.array_equal(o["pairs"], g[f"{mode}_pairs64"])
    assert abs(o["rmse"] - float(g[f"{mode}_rmse64"])) < 1e-7


def test_nearest_is_exact_on_random_clouds():
    rng = np.random.default_rng(0)
    t = rng.normal(0, 1, (3000, 3)).astype(np.float32)
    q = rng.normal(0, 1, (500, 3)).astype(np.float32)
    nn = oracleapi.nearest(q, t)
    d = ((q[:, None, :].astype(np.float64) - t[None].astype(np.float64)) ** 2).sum(-1)
    best = d.min(1)
    assert np.all(d[np.arange(len(q)), nn] <= best * (1 + 1e-6))


def test_error_path_and_scaling():
    rng = np.random.default_rng(1)
    t = rng.normal(0, 1, (500, 3)).astype(np.float32)
    s = (t + 0.01).astype(np.float32)
    n = np.tile(np.array([0, 0, 1], np.float32), (500, 1))
    assert oracleapi.icp(s, t, n, np.eye(4), 3, 0.5, scaling=2.0) is None  # PointToPlane refuses scaling != 1
    a = oracleapi.icp(s, t, None, np.eye(4), 3, 0.5, scaling=1.0)
    b = oracleapi.icp(s, t, None, np.eye(4), 3, 1.0, scaling=2.0)  # same problem in scaled units
    assert np.allclose(a["T"], b["T"], atol=1e-5)


@pytest.mark.skipif(not refapi.available("f64"), reason="oracle/_ref not built")
@pytest.mark.parametrize("mode", ["plane", "point"])
def test_oracle_icp_vs_compiled_reference(g, mode):
    nrm = g["nrm"] if mode == "plane" else None
    T0 = np.eye(4)
    T0[:3, 3] = [0.004, -0.003, 0.002]
    o = oracleapi.icp(g["src"], g["tgt"], nrm, T0, 6, 0.04)
    r = refapi.icp(g["src"], g["tgt"], nrm, T0, 6, 0.04, "f64")
    dt, dr = pose_delta(o["T"], r["T"])
    assert dt < 1e-7 and dr < 1e-7
    assert np.array_equal(o["pairs"], r["pairs"])
    x = np.array([0.01, -0.02, 0.03, 0.1, -0.2, 0.05])
    assert np.abs(oracleapi.se3_exp(x) - refapi.se3_exp(x, "f64")).max() < 1e-14


@pytest.mark.skipif(not refapi.available("f64"), reason="oracle/_ref not built")
@pytest.mark.parametrize("mode", ["plane", "point"])
@pytest.mark.parametrize("max_iter", [0, 1, 2, 3])
def test_oracle_icp_unconverged_vs_compiled_reference(g, mode, max_iter):
    """Few iterations from a large offset: the closing CountInliers (ICP.cpp:90,206) reuses the neighbours found under the
    previous pose -- whatever their distance -- and only re-tests them with the final pose.  DenseSlam calls PointToPoint with
    max_iteration = 1 (example/DenseFusion/DenseSlam.cpp:72-90), so this is a caller-visible case."""
    nrm = g["nrm"] if mode == "plane" else None
    T0 = np.eye(4)
    T0[:3, 3] = [0.03, -0.01, 0.015]
    o = oracleapi.icp(g["src"], g["tgt"], nrm, T0, max_iter, 0.04)
    r = refapi.icp(g["src"], g["tgt"], nrm, T0, max_iter, 0.04, "f64")
    assert np.array_equal(o["pairs"], r["pairs"])
    if max_iter == 0:
        assert len(r["pairs"]) == 0 and np.isnan(o["rmse"]) and np.isnan(r["rmse"])
        return
    dt, dr = pose_delta(o["T"], r["T"])
    assert dt < 1e-6 and dr < 1e-6, (dt, dr)
    assert abs(o["rmse"] - r["rmse"]) < 1e-7
