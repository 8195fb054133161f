"""The float-exact EstimateRigidTransformation restatement and the RANSAC hypothesis evaluation built on it, pinned bit for bit to
the compiled reference (geometry::EstimateRigidTransformation, Geometry.cpp:107-151; TransformationModel.hpp:28-96).  GRANSAC's
own sampling is seeded from std::random_device and cannot be pinned; everything downstream of the sample can, and is."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracleapi, refapi
from test_kdtree_emulated import ransac_case


def _ref_kabsch(a, b):
    T = np.zeros(16, np.float64)
    L = refapi.lib("f32")
    L.ref_kabsch.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    L.ref_kabsch(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), len(a), T.ctypes.data_as(C.c_void_p))
    return T.reshape(4, 4).T.astype(np.float32)


def test_float_kabsch_is_the_references_bit_for_bit(ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(0)
    for trial in range(600):
        n = 8 if trial % 2 == 0 else int(rng.integers(3, 40))
        a = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
        if trial % 3 == 0:
            R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            if trial % 6 == 0:
                R[:, 0] *= -1                                   # a reflection: the det < 0 branch
            b = (a @ R.T + rng.uniform(-1, 1, 3) + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
        else:
            b = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
        assert np.array_equal(oracleapi.kabsch_f32(a, b).view(np.uint32), _ref_kabsch(a, b).view(np.uint32)), trial


def test_ransac_hypotheses_match_the_compiled_reference(ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    a, b, T_true, rng = ransac_case(1500)
    for trial in range(150):
        s8 = rng.choice(len(a), 8, replace=False).astype(np.int32)
        thr = [0.1, 0.05, 0.5][trial % 3]
        T, flags = oracleapi.ransac_hypothesis(a, b, s8, thr)
        frac, rflags = refapi.ransac_hypothesis(a, b, s8, thr)
        assert np.array_equal(flags, rflags) and frac == flags.sum() / len(a)
    # the selection rule: first strictly best
    samples = np.stack([rng.choice(len(a), 8, replace=False) for _ in range(200)]).astype(np.int32)
    samples[50] = samples[7]
    w, T, ids = oracleapi.ransac_select(a, b, samples, 0.1)
    counts = [int(oracleapi.ransac_hypothesis(a, b, s, 0.1)[1].sum()) for s in samples]
    assert w == int(np.argmax(counts)) and len(ids) == max(counts)
    assert T.shape == (4, 4) and T[3, 3] == 1
