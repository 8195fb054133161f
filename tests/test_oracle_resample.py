"""Oracle restatement of CubeHandler::Transform / TransformNearest / Merge (reference src/Integration/CubeHandler.h:145-338,
VoxelCube.cpp:6-50) pinned bit for bit to the compiled reference (oracle/_ref), including TransformNearest's forgotten
c_para (its result runs at VoxelResolution 0.01 whatever the source used)."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from fusion_common import small_scene
from onepiece_b200 import scenes
from oracle import oracleapi, refapi

T1 = scenes.se3_exp([0.03, -0.02, 0.05, 0.1, -0.2, 0.15]).astype(np.float32)


def _pair(res):
    cam, _, frames, *_ = small_scene()
    ov, rv = oracleapi.OracleVolume(cam, res), refapi.RefVolume(cam, res)
    for d, c, p in frames:
        ov.integrate(d, c, p)
        rv.integrate(d, c, p)
    return cam, ov, rv


@pytest.mark.parametrize("res", [0.02, 0.01])
@pytest.mark.parametrize("nearest", [True, False])
def test_transform_matches_the_compiled_reference(ref_available, res, nearest):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    _, ov, rv = _pair(res)
    a, b = ov.transform(T1, nearest), rv.transform(T1, nearest)
    assert a.resolution() == b.resolution() == np.float32(0.01 if nearest else res)
    (ai, av), (bi, bv) = a.download(), b.download()
    assert np.array_equal(ai, bi)
    assert_bit_equal(av, bv, "transformed voxels")
    assert (av[:, :, 1] > 0).sum() > 10000


def test_merge_matches_the_compiled_reference(ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    cam, ov, rv = _pair(0.02)
    ov2, rv2 = oracleapi.OracleVolume(cam, 0.02), refapi.RefVolume(cam, 0.02)
    d, c = scenes.wavy_wall(cam, 5)
    ov2.integrate(d, c, T1)
    rv2.integrate(d, c, T1)
    n_before = ov2.num_cubes()
    assert ov2.merge(ov) == 0
    rv2.merge(rv)
    (ai, av), (bi, bv) = ov2.download(), rv2.download()
    assert len(ai) > n_before and np.array_equal(ai, bi)
    assert_bit_equal(av, bv, "merged voxels")
    # Merge(another, trans) = Transform + Merge (CubeHandler.h:168-177)
    ov2.merge(ov.transform(T1, False))
    rv2.merge(rv, T1)
    (ai, av), (bi, bv) = ov2.download(), rv2.download()
    assert np.array_equal(ai, bi)
    assert_bit_equal(av, bv, "merged transformed voxels")
    # different resolutions: the reference warns and leaves the volume alone
    other = oracleapi.OracleVolume(cam, 0.01)
    assert ov2.merge(other) == -1


def test_transform_properties():
    """No reference needed: identity transform reproduces the volume; nearest with the source resolution is a pure gather."""
    cam, _, frames, ids, vox, _ = small_scene()
    ov = oracleapi.OracleVolume(cam, 0.02)
    ov.upload(ids, vox)
    same = ov.transform(np.eye(4), True, alloc_res=0.02)
    si, sv = same.download()
    assert np.array_equal(si, ids)
    assert_bit_equal(sv, vox, "identity nearest transform")
    tri = ov.transform(np.eye(4), False)
    ti, tv = tri.download()
    valid = vox[:, :, 1] > 0
    # identity trilinear: sampling exactly at voxel centres gives the voxel back wherever it is observed
    lookup = {tuple(i): k for k, i in enumerate(ti)}
    k = np.array([lookup[tuple(i)] for i in ids])
    assert np.allclose(tv[k][valid][:, 0], vox[valid][:, 0], atol=1e-5)
