"""TriangleMesh::ClusteringSimplify / ComputeNormals on the device against the oracle, bit for bit and in the same order."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from fusion_common import small_scene
from oracle import oracleapi

pytestmark = pytest.mark.gpu


def _volume_and_mesh():
    from onepiece_b200.volume import CubeHandler
    cam, res, frames, *_ = small_scene()
    gv = CubeHandler(cam, res, max_cubes=4096)
    for d, c, p in frames:
        gv.IntegrateImage(d, c, p)
    pts, col, tri = gv.ExtractTriangleMesh()
    return gv, pts, col, tri


def _same(mesh, o, what):
    assert len(mesh.points) == len(o[0]) and len(mesh.triangles) == len(o[2]), what
    assert_bit_equal(mesh.points, o[0], what + ": points")
    if o[1] is not None:
        assert_bit_equal(mesh.colors, o[1], what + ": colours")
    assert np.array_equal(mesh.triangles, o[2]), what + ": triangles"


@pytest.mark.parametrize("grid", [0.02, 0.01, 0.05])
def test_clustering_simplify(grid):
    from onepiece_b200.mesh import TriangleMesh
    gv, pts, col, tri = _volume_and_mesh()
    m = TriangleMesh(pts, col, tri).ClusteringSimplify(grid)
    o = oracleapi.clustering_simplify(pts, col, tri, grid)
    _same(m, o, "MC mesh")
    assert 0 < len(m.points) < len(pts) // 3
    # shared vertices: the clustered mesh through a second, coarser pass; no colours this time
    m2 = TriangleMesh(m.points, None, m.triangles).ClusteringSimplify(2.5 * grid)
    _same(m2, oracleapi.clustering_simplify(m.points, None, m.triangles, 2.5 * grid), "second pass")
    # extract + simplify fused on the device
    _same(gv.ExtractTriangleMeshClustered(grid), o, "fused extract + simplify")


def test_compute_normals_and_meshes_with_normals():
    from onepiece_b200.mesh import TriangleMesh
    _, pts, col, tri = _volume_and_mesh()
    m = TriangleMesh(pts, col, tri).ClusteringSimplify(0.02)
    m.ComputeNormals()
    assert_bit_equal(m.normals, oracleapi.compute_normals(m.points, m.triangles), "vertex normals")
    raw = TriangleMesh(pts, col, tri)
    raw.ComputeNormals()                               # every vertex used once: the face normal itself
    assert_bit_equal(raw.normals, oracleapi.compute_normals(pts, tri), "MC mesh normals")
    m2 = m.ClusteringSimplify(0.05)                    # has normals -> recomputed on the result
    o2 = oracleapi.clustering_simplify(m.points, m.colors, m.triangles, 0.05)
    _same(m2, o2, "mesh with normals")
    assert_bit_equal(m2.normals, oracleapi.compute_normals(o2[0], o2[2]), "recomputed normals")


def test_edge_cases():
    from onepiece_b200 import capi
    from onepiece_b200.mesh import TriangleMesh
    one = TriangleMesh(np.array([[0.1, 0.1, 0.1], [1.1, 0.1, 0.1], [0.1, 1.1, 0.1]], np.float32), None, [[0, 1, 2]])
    same = one.ClusteringSimplify(0.0)                 # the reference prints an error and returns the mesh unchanged
    assert np.array_equal(same.points, one.points) and np.array_equal(same.triangles, one.triangles)
    kept = one.ClusteringSimplify(1.0)
    assert_bit_equal(kept.points, one.points, "lone triangle")
    gone = one.ClusteringSimplify(10.0)                # all corners in one cell: the triangle is dropped, nothing is left
    assert len(gone.points) == 0 and len(gone.triangles) == 0
    empty = TriangleMesh().ClusteringSimplify(0.1)
    assert len(empty.points) == 0
    bad = TriangleMesh(one.points, None, [[0, 1, 7]])
    with pytest.raises(capi.OpbError) as e:
        bad.ClusteringSimplify(1.0)
    assert e.value.code == capi.OPB_ERR_INVALID


def test_cubes_file_round_trip_through_the_device(tmp_path):
    from onepiece_b200.volume import CubeHandler
    gv, *_ = _volume_and_mesh()
    ids, vox = gv.GetCubeMap()
    assert gv.WriteToFile(str(tmp_path / "v.cubes"))
    other = CubeHandler(gv.camera, 0.02, max_cubes=4096)
    assert other.ReadFromFile(str(tmp_path / "v.cubes"))
    oi, ovx = other.GetCubeMap()
    stored = (np.abs(vox[:, :, 0]) < 1) & (vox[:, :, 1] != 0)
    expect = vox.copy()
    expect[~stored] = (999.0, 0.0, -1.0, -1.0, -1.0)
    assert np.array_equal(oi, ids)
    assert_bit_equal(ovx, expect, "volume after .cubes round trip")
    assert other.CountMesh() == gv.CountMesh()         # Marching Cubes only reads what the format stores
    assert not other.ReadFromFile(str(tmp_path / "missing.cubes"))


@pytest.mark.parametrize("grid", [0.01, 0.05, 0.5])
def test_point_cloud_downsample(grid):
    """PointCloud::DownSample on the device: same points in the same order, colours and normals averaged alike; grid 0.5 puts
    thousands of points into one cell (the long-segment sort)."""
    from onepiece_b200 import registration as reg
    from test_oracle_meshpost import _room_cloud
    pts, col, nrm = _room_cloud()
    cloud, gc = reg.PointCloud(pts, nrm).DownSample(grid, colors=col)
    op, oc, on = oracleapi.downsample(pts, col, nrm, grid)
    assert len(cloud.points) == len(op)
    assert_bit_equal(cloud.points, op, "points")
    assert_bit_equal(gc, oc, "colours")
    assert_bit_equal(cloud.normals, on, "normals")
    plain = reg.PointCloud(pts).DownSample(grid)
    assert_bit_equal(plain.points, op, "points only") and plain.normals is None
    assert len(reg.PointCloud(np.zeros((0, 3), np.float32)).DownSample(grid).points) == 0


def test_clustering_with_cells_that_hold_thousands_of_vertices():
    from onepiece_b200.mesh import TriangleMesh
    _, pts, col, tri = _volume_and_mesh()
    for grid in (0.3, 1.5):
        m = TriangleMesh(pts, col, tri).ClusteringSimplify(grid)
        _same(m, oracleapi.clustering_simplify(pts, col, tri, grid), f"grid {grid}")
