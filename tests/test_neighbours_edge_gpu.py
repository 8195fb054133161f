"""Empty and degenerate inputs through the kernels either side of the hot path (depth pre-filter, resampling / merging, halo
exchange, mesh post-processing): nothing to do must mean nothing done, not a crash or garbage."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from onepiece_b200 import scenes
from oracle import oracleapi

pytestmark = pytest.mark.gpu


def _cam():
    c0 = scenes.Camera()
    return scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0)


def test_empty_volume_through_every_entry_point():
    from onepiece_b200.volume import CubeHandler
    cam = _cam()
    empty = CubeHandler(cam, 0.02, max_cubes=256)
    T = scenes.se3_exp([0.1, 0.0, -0.1, 0.2, 0.1, 0.0]).astype(np.float32)
    for r in (empty.Transform(T), empty.TransformNearest(T)):
        assert r.NumCubes() == 0 and r.CountMesh() == (0, 0)
    other = CubeHandler(cam, 0.02, max_cubes=256)
    assert other.Merge(empty) and other.NumCubes() == 0
    m = empty.ExtractTriangleMeshClustered(0.02)
    assert len(m.points) == 0 and len(m.triangles) == 0
    sharded = CubeHandler(cam, 0.02, max_cubes=256, shard=(0, 2, 0, 1))
    assert sharded.HaloCount() == 0
    sharded.HaloImport(np.zeros((0, 3), np.int32), np.zeros((0, 5, 64), np.float32), 0)
    assert sharded.NumGhostCubes() == 0
    # an all-invalid frame allocates nothing, and merging INTO an empty volume copies the other one
    d, c = scenes.wavy_wall(cam, 0)
    empty.IntegrateImage(np.zeros_like(d), c, np.eye(4, dtype=np.float32))
    assert empty.NumCubes() == 0
    full = CubeHandler(cam, 0.02, max_cubes=2048)
    full.IntegrateImage(d, c, np.eye(4, dtype=np.float32))
    grown = CubeHandler(cam, 0.02, max_cubes=2048)
    assert grown.Merge(full)
    (gi, gv), (fi, fv) = grown.GetCubeMap(), full.GetCubeMap()
    assert np.array_equal(gi, fi)
    assert_bit_equal(gv, fv, "merge into an empty volume is a copy")


def test_prefilter_on_images_without_a_signal():
    from onepiece_b200 import imageproc
    pf = imageproc.DepthPrefilter(160, 120)
    zeros16 = np.zeros((120, 160), np.uint16)
    conv, out = pf.run(zeros16, 1000.0)
    assert not conv.any() and not out.any()                      # constant image: copied
    one = zeros16.copy()
    one[60, 80] = 2000                                          # a single valid pixel among holes
    conv, out = pf.run(one, 1000.0)
    assert_bit_equal(out, oracleapi.bilateral_filter(oracleapi.convert_depth_32f(one, 1000.0)), "single pixel")
    assert out[60, 80] == np.float32(2.0) and np.count_nonzero(out) == 1
    far = np.full((120, 160), 65535, np.uint16)                 # largest raw value
    conv, out = pf.run(far, 1000.0)
    assert_bit_equal(conv, oracleapi.convert_depth_32f(far, 1000.0), "max u16")


def test_transform_by_identity_and_back():
    """Nearest-neighbour resampling by the identity at the source resolution is the volume itself; T then T^-1 returns every
    observed voxel to a cube that exists again."""
    from onepiece_b200.volume import CubeHandler
    cam = _cam()
    v = CubeHandler(cam, 0.02, max_cubes=4096)
    d, c = scenes.wavy_wall(cam, 1)
    v.IntegrateImage(d, c, np.eye(4, dtype=np.float32))
    same = v.TransformNearest(np.eye(4), result_resolution=0.0)
    (si, sv), (vi, vv) = same.GetCubeMap(), v.GetCubeMap()
    assert np.array_equal(si, vi)
    assert_bit_equal(sv, vv, "identity transform")
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [0.16, -0.32, 0.48]                               # whole cubes: 8 voxels of 2 cm per cube
    there = v.TransformNearest(T, result_resolution=0.0)
    back = there.TransformNearest(np.linalg.inv(T), result_resolution=0.0)
    bi, bv = back.GetCubeMap()
    lookup = {tuple(i): k for k, i in enumerate(bi)}
    observed = vv[:, :, 1] > 0
    for k, cid in enumerate(vi):
        if observed[k].any():
            assert tuple(cid) in lookup
            assert_bit_equal(bv[lookup[tuple(cid)]][observed[k]], vv[k][observed[k]], "round trip")
