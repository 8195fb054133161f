"""Pins the plain-C oracle (oracle/opb_oracle.c) against the reference's own translation units compiled
unmodified (oracle/_ref).  Runs wherever oracle/_ref exists (the build container; the .so also travels to
the GPU box); skipped otherwise."""
import numpy as np
import pytest

from conftest import assert_bit_equal, canon_triangles
from onepiece_b200 import scenes
from oracle import oracleapi, refapi

pytestmark = pytest.mark.skipif(not refapi.available("f32"), reason="oracle/_ref not built (needs /root/reference)")


def small_camera():
    c = scenes.Camera()
    return scenes.Camera(c.fx / 4, c.fy / 4, c.cx / 4, c.cy / 4, 160, 120, 1000.0)


def random_pose(rng, scale=1.0):
    return scenes.se3_exp(rng.normal(0, 1, 6) * np.array([1, 1, 1, .5, .5, .5]) * scale).astype(np.float32)


def test_pose_inverse():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        T = random_pose(rng)
        assert_bit_equal(refapi.pose_inverse(T).astype(np.float32), oracleapi.pose_inverse(T), "pose inverse")


def test_frustum_planes_and_containment():
    rng = np.random.default_rng(1)
    cam = scenes.Camera()
    ov = oracleapi.OracleVolume(cam)
    for _ in range(100):
        T = random_pose(rng)
        # points concentrated around the frustum boundary planes as well as far away
        pts = rng.normal(0, 2, (400, 3)).astype(np.float32)
        pr, mr = refapi.frustum(cam, T, 5.0, 0.5, pts)
        po, mo = ov.frustum(T, pts)
        assert_bit_equal(pr.astype(np.float32), po, "planes")
        assert np.array_equal(mr, mo)


def test_get_sdf_random_points():
    rng = np.random.default_rng(2)
    cam = scenes.Camera()
    d, _ = scenes.wavy_wall(cam, 1)
    rv, ov = refapi.RefVolume(cam, 0.005), oracleapi.OracleVolume(cam, 0.005)
    for _ in range(5):
        T = random_pose(rng, 0.1)
        pts = (rng.uniform(-1.5, 1.5, (20000, 3)) + np.array([0, 0, 2.0])).astype(np.float32)
        pts[:50] = 0  # camera centre: 0/0 projections
        pts[50:100, 2] = -1.0  # behind the camera
        assert_bit_equal(rv.get_sdf(d, T, pts), ov.get_sdf(d, T, pts), "GetSDF")


@pytest.mark.parametrize("u16", [False, True])
def test_bounding_prepare_integrate_small(u16):
    cam = small_camera()
    rng = np.random.default_rng(3)
    rv, ov = refapi.RefVolume(cam, 0.02), oracleapi.OracleVolume(cam, 0.02)
    for k in range(4):
        d, c = scenes.wavy_wall(cam, k)
        if u16:
            d = np.clip(np.rint(d * 1000), 0, 65535).astype(np.uint16)
        T = random_pose(rng, 0.1 if k else 0.0)
        mxr, mnr = rv.bounding(d, T)
        mxo, mno = ov.bounding(d, T)
        assert_bit_equal(mxr.astype(np.float32), mxo, "bbox max")
        assert_bit_equal(mnr.astype(np.float32), mno, "bbox min")
        rv.integrate(d, c, T)
        ov.integrate(d, c, T)
    ri, rvx = rv.download()
    oi, ovx = ov.download()
    assert np.array_equal(ri, oi)
    assert_bit_equal(rvx, ovx, "voxels")
    # cube list of one more frame, in the reference's loop order
    d, c = scenes.wavy_wall(cam, 9)
    if u16:
        d = np.clip(np.rint(d * 1000), 0, 65535).astype(np.uint16)
    assert np.array_equal(rv.prepare_cubes(d, np.eye(4)), ov.prepare_cubes(d, np.eye(4)))


def test_marching_cubes_all_256_cases():
    rng = np.random.default_rng(4)
    corners = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.float32)
    for case in range(256):
        sign = np.array([1.0 if (case >> i) & 1 else -1.0 for i in range(8)], np.float32)
        sdf = (sign * rng.uniform(0.01, 0.09, 8)).astype(np.float32)
        col = rng.uniform(0, 1, (8, 3)).astype(np.float32)
        c = (corners * 0.005 + rng.uniform(-1, 1, 3)).astype(np.float32)
        xr, cr = refapi.marching_cube_cell(c, sdf, col)
        xo, co = oracleapi.marching_cube_cell(c, sdf, col)
        assert_bit_equal(xr, xo, f"case {case} xyz")
        assert_bit_equal(cr, co, f"case {case} rgb")


def test_extract_mesh_small():
    cam = small_camera()
    rv, ov = refapi.RefVolume(cam, 0.02), oracleapi.OracleVolume(cam, 0.02)
    T = scenes.se3_exp([0.05, -0.02, 0.03, 0.02, -0.03, 0.01]).astype(np.float32)
    for k, pose in enumerate([np.eye(4, dtype=np.float32), T]):
        d, c = scenes.wavy_wall(cam, k)
        rv.integrate(d, c, pose)
        ov.integrate(d, c, pose)
    _, rp, rc, rt = rv.extract_mesh()
    op, oc = ov.extract_mesh()
    assert len(rp) == len(op) and len(rt) * 3 == len(rp)
    assert_bit_equal(canon_triangles(rp, rc), canon_triangles(op, oc), "mesh")


def test_integrate_full_size_one_frame():
    cam = scenes.Camera()
    rv, ov = refapi.RefVolume(cam, 0.005), oracleapi.OracleVolume(cam, 0.005)
    d, c = scenes.wavy_wall(cam, 0)
    rv.integrate(d, c, np.eye(4))
    n = ov.integrate(d, c, np.eye(4))
    assert n == rv.num_cubes() == ov.num_cubes()
    ri, rvx = rv.download()
    oi, ovx = ov.download()
    assert np.array_equal(ri, oi)
    assert_bit_equal(rvx, ovx, "voxels")
