"""Host-side float32 restatements inside the product library (pose inverse, frustum planes) against the
oracle, bit for bit.  No GPU needed: both are plain host functions exported by the C-ABI."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_bit_equal
from onepiece_b200 import capi, scenes
from oracle import oracleapi


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def random_pose(rng, scale=1.0):
    return scenes.se3_exp(rng.normal(0, 1, 6) * np.array([1, 1, 1, .5, .5, .5]) * scale).astype(np.float32)


def test_pose_inverse_matches_oracle_bitwise():
    rng = np.random.default_rng(7)
    for _ in range(1000):
        T = random_pose(rng)
        cm = np.ascontiguousarray(T.T).reshape(16)
        inv = np.zeros(16, np.float32)
        capi.lib.opb_pose_inverse(_p(cm), _p(inv))
        assert_bit_equal(inv.reshape(4, 4).T, oracleapi.pose_inverse(T), "pose inverse")


def test_pose_inverse_general_matrix():
    # the reference inverts whatever 4x4 it is given, rigid or not
    rng = np.random.default_rng(8)
    for _ in range(200):
        M = rng.normal(0, 1, (4, 4)).astype(np.float32)
        cm = np.ascontiguousarray(M.T).reshape(16)
        inv = np.zeros(16, np.float32)
        capi.lib.opb_pose_inverse(_p(cm), _p(inv))
        assert_bit_equal(inv.reshape(4, 4).T, oracleapi.pose_inverse(M), "general inverse")
        assert np.allclose(inv.reshape(4, 4).T @ M, np.eye(4), atol=1e-2)


@pytest.mark.parametrize("cam", [scenes.Camera(), scenes.Camera(517.3, 516.5, 318.6, 255.3, 640, 480, 5000.0),
                                 scenes.Camera().scaled(2)])
def test_frustum_planes_match_oracle_bitwise(cam):
    rng = np.random.default_rng(9)
    ov = oracleapi.OracleVolume(cam)
    for _ in range(200):
        T = random_pose(rng)
        cm = np.ascontiguousarray(T.T).reshape(16)
        pl = np.zeros(24, np.float32)
        capi.lib.opb_frustum_planes(cam.fx, cam.fy, cam.cy, cam.width, cam.height, 0.5, 5.0, _p(cm), _p(pl))
        po, _ = ov.frustum(T, np.zeros((1, 3), np.float32))
        assert_bit_equal(pl.reshape(6, 4), po, "frustum planes")
