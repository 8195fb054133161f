"""CubeHandler::Transform / TransformNearest / Merge on the device against the oracle, bit for bit."""
import numpy as np
import pytest

from conftest import assert_bit_equal, canon_triangles
from fusion_common import small_scene
from onepiece_b200 import scenes
from oracle import oracleapi

pytestmark = pytest.mark.gpu

T1 = scenes.se3_exp([0.03, -0.02, 0.05, 0.1, -0.2, 0.15]).astype(np.float32)


def _pair(res, max_cubes=1 << 14):
    from onepiece_b200.volume import CubeHandler
    cam, _, frames, *_ = small_scene()
    ov, gv = oracleapi.OracleVolume(cam, res), CubeHandler(cam, res, max_cubes=max_cubes)
    for d, c, p in frames:
        ov.integrate(d, c, p)
        gv.IntegrateImage(d, c, p)
    return cam, ov, gv


def _same(g, o, what):
    (gi, gvx), (oi, ovx) = g.GetCubeMap(), o.download()
    assert np.array_equal(gi, oi), f"{what}: cube sets differ ({len(gi)} vs {len(oi)})"
    assert_bit_equal(gvx, ovx, what)


@pytest.mark.parametrize("res", [0.02, 0.01])
@pytest.mark.parametrize("nearest", [True, False])
def test_transform(res, nearest):
    _, ov, gv = _pair(res)
    g = gv.TransformNearest(T1) if nearest else gv.Transform(T1)
    o = ov.transform(T1, nearest)
    assert np.float32(g.desc.voxel_resolution) == np.float32(o.resolution()) == np.float32(0.01 if nearest else res)
    _same(g, o, "transformed volume")
    _same(gv, ov, "source volume untouched")
    # the mesh of the result (what example/ImageSequenceIntegration.cpp:48-53 extracts) is the reference's too
    gp, gc, _ = g.ExtractTriangleMesh()
    op, oc = o.extract_mesh()
    assert len(gp) == len(op)
    assert np.array_equal(canon_triangles(gp, gc), canon_triangles(op, oc))


def test_transform_capacity_is_grown_or_reported():
    from onepiece_b200 import capi
    _, ov, gv = _pair(0.02)
    n = gv.NumCubes()
    g = gv.Transform(T1)                                  # automatic sizing
    assert g.NumCubes() == ov.transform(T1, False).num_cubes() > n
    with pytest.raises(capi.OpbError) as e:
        gv.Transform(T1, max_cubes=n // 2)                # a fixed capacity that is too small is an error, not a truncation
    assert e.value.code == capi.OPB_ERR_CAPACITY


def test_merge():
    from onepiece_b200 import capi
    from onepiece_b200.volume import CubeHandler
    cam, ov, gv = _pair(0.02)
    ov2, gv2 = oracleapi.OracleVolume(cam, 0.02), CubeHandler(cam, 0.02, max_cubes=1 << 14)
    d, c = scenes.wavy_wall(cam, 5)
    ov2.integrate(d, c, T1)
    gv2.IntegrateImage(d, c, T1)
    assert gv2.Merge(gv) and ov2.merge(ov) == 0
    _same(gv2, ov2, "merged volume")
    assert gv2.Merge(gv, T1)                               # Merge(another, trans) (CubeHandler.h:168-177)
    ov2.merge(ov.transform(T1, False))
    _same(gv2, ov2, "merged transformed volume")
    # integration continues on the merged volume, still bit-exact (interpolated weights -> IEEE-division path)
    d, c = scenes.wavy_wall(cam, 6)
    ov2.integrate(d, c, np.eye(4, dtype=np.float32))
    gv2.IntegrateImage(d, c, np.eye(4, dtype=np.float32))
    _same(gv2, ov2, "integration after merge")
    other = CubeHandler(cam, 0.01, max_cubes=64)
    assert gv2.Merge(other) is False                       # resolution mismatch: warning, nothing merged
    _same(gv2, ov2, "unchanged after refused merge")
    tiny = CubeHandler(cam, 0.02, max_cubes=8)                 # the reference's map is unbounded: the pool makes room
    assert tiny.Merge(gv) and tiny.NumCubes() == gv.NumCubes()
    ti, tv = tiny.GetCubeMap()
    gi, gvx = gv.GetCubeMap()
    ot, og = np.lexsort(ti.T[::-1]), np.lexsort(gi.T[::-1])
    assert np.array_equal(ti[ot], gi[og]) and np.array_equal(tv[ot].view(np.uint32), gvx[og].view(np.uint32))
