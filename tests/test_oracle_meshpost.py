"""Mesh post-processing and on-disk formats: the oracle's ClusteringSimplify / ComputeNormals pinned
bit for bit to the compiled reference (oracle/_ref), and the Python writers / readers of the reference's file formats checked
byte for byte against files the reference itself writes and reads."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from fusion_common import small_scene
from oracle import oracleapi, refapi


def _mc_mesh():
    *_, (pts, col) = small_scene()
    return pts, col, np.arange(len(pts), dtype=np.uint32).reshape(-1, 3)


@pytest.mark.parametrize("grid", [0.02, 0.01, 0.05])
def test_clustering_simplify_matches_the_compiled_reference(ref_available, grid):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    pts, col, tri = _mc_mesh()
    a = oracleapi.clustering_simplify(pts, col, tri, grid)
    b = refapi.clustering_simplify(pts, col, tri, grid)
    assert len(a[0]) == len(b[0]) < len(pts) and len(a[2]) == len(b[2]) < len(tri)
    assert_bit_equal(a[0], b[0], "clustered points")
    assert_bit_equal(a[1], b[1], "clustered colours")
    assert np.array_equal(a[2], b[2])
    # a mesh with shared vertices (the clustered mesh itself) through a second, coarser pass
    a2 = oracleapi.clustering_simplify(a[0], a[1], a[2], 2.5 * grid)
    b2 = refapi.clustering_simplify(a[0], a[1], a[2], 2.5 * grid)
    assert_bit_equal(a2[0], b2[0], "second pass points")
    assert np.array_equal(a2[2], b2[2])


def test_compute_normals_and_simplify_with_normals(ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    pts, col, tri = _mc_mesh()
    p, c, t = oracleapi.clustering_simplify(pts, col, tri, 0.02)
    n1, n2 = oracleapi.compute_normals(p, t), refapi.compute_normals(p, t)
    assert_bit_equal(n1, n2, "vertex normals")
    ln = np.linalg.norm(n1, axis=1)
    assert ((np.abs(ln - 1) < 1e-5) | (ln == 0)).all() and (ln > 0).mean() > 0.99   # unit length, or zero where all faces are degenerate
    # a mesh WITH normals: CompactMesh recomputes them on the simplified mesh (MeshSimplification.cpp:338-342)
    rp, rc, rt, rn, _ = refapi.clustering_simplify(p, c, t, 0.05, with_normals=True)
    op, oc, ot = oracleapi.clustering_simplify(p, c, t, 0.05)
    assert_bit_equal(op, rp, "points")
    assert np.array_equal(ot, rt)
    assert_bit_equal(oracleapi.compute_normals(op, ot), rn, "recomputed normals")


def test_clustering_edge_cases():
    assert oracleapi.clustering_simplify(np.zeros((3, 3)), None, [[0, 1, 2]], 0.0) is None      # the reference's error path
    p, c, t = oracleapi.clustering_simplify(np.zeros((3, 3)), None, [[0, 1, 2]], 1.0)           # all corners in one cell
    assert len(p) == 0 and len(t) == 0 and c is None
    tri_pts = np.array([[0.1, 0.1, 0.1], [1.1, 0.1, 0.1], [0.1, 1.1, 0.1]], np.float32)
    p, _, t = oracleapi.clustering_simplify(tri_pts, None, [[0, 1, 2]], 1.0)
    assert_bit_equal(p, tri_pts, "a lone triangle survives unchanged")
    assert np.array_equal(t, [[0, 1, 2]])


def test_ply_writer_is_byte_identical_to_the_reference(ref_available, tmp_path):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    from onepiece_b200.mesh import write_ply
    pts, col, tri = _mc_mesh()
    p, c, t = oracleapi.clustering_simplify(pts, col, tri, 0.02)
    n = oracleapi.compute_normals(p, t)
    for name, nrm, colors in (("full", n, c), ("plain", None, None), ("colors", None, c)):
        ours, theirs = tmp_path / f"ours_{name}.ply", tmp_path / f"ref_{name}.ply"
        assert write_ply(str(ours), p, nrm, colors, t)
        assert refapi.write_ply(theirs, p, nrm, colors, t)
        assert ours.read_bytes() == theirs.read_bytes(), name


def test_cubes_stream_round_trips_through_the_reference(ref_available, tmp_path):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    from onepiece_b200 import formats
    cam, res, _, ids, vox, _ = small_scene()
    # ours -> the reference's ReadFromFile
    formats.write_cubes(str(tmp_path / "ours.cubes"), ids, vox)
    rv = refapi.RefVolume(cam, res)
    assert rv.read(str(tmp_path / "ours.cubes"))
    ri, rvx = rv.download()
    stored = (np.abs(vox[:, :, 0]) < 1) & (vox[:, :, 1] != 0)
    expect = vox.copy()
    expect[~stored] = (999.0, 0.0, -1.0, -1.0, -1.0)       # what the format does not store comes back as a default voxel
    assert np.array_equal(ri, ids)
    assert_bit_equal(rvx, expect, "our .cubes file read by the reference")
    # the reference's WriteToFile -> ours
    rv2 = refapi.RefVolume(cam, res)
    rv2.upload(ids, vox)
    assert rv2.write(str(tmp_path / "ref.cubes"))
    gi, gv = formats.read_cubes(str(tmp_path / "ref.cubes"))
    order = np.lexsort((gi[:, 2], gi[:, 1], gi[:, 0]))
    assert np.array_equal(gi[order], ids)
    assert_bit_equal(gv[order], expect, "the reference's .cubes file read by us")
    assert formats.read_cubes(str(tmp_path / "missing.cubes")) is None


def test_trajectory_round_trip(tmp_path):
    from onepiece_b200 import formats, scenes
    poses = [scenes.se3_exp(np.array([0.01, -0.02, 0.03, 0.1, 0.2, -0.1]) * k) for k in range(4)]
    formats.write_trajectory(str(tmp_path / "trajectory.txt"), poses)
    back = formats.read_trajectory(str(tmp_path / "trajectory.txt"))
    assert len(back) == 4 and all(np.array_equal(a, b) for a, b in zip(poses, back))


def _room_cloud():
    from onepiece_b200 import scenes
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 2, c0.fy / 2, c0.cx / 2, c0.cy / 2, 320, 240, 1000.0)
    d, c, _, n = scenes.room(cam, 0, with_normals=True)
    m = (d > 0).reshape(-1)
    return scenes.backproject(d, cam), (c.reshape(-1, 3)[m] / 255.0).astype(np.float32), np.ascontiguousarray(n.reshape(-1, 3)[m])


@pytest.mark.parametrize("grid", [0.01, 0.05, 0.5])
def test_downsample_matches_the_compiled_reference(ref_available, grid):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    pts, col, nrm = _room_cloud()
    a, b = oracleapi.downsample(pts, col, nrm, grid), refapi.downsample(pts, col, nrm, grid)
    assert len(a[0]) == len(b[0]) < len(pts)
    for x, y, what in zip(a, b, ("points", "colours", "normals")):
        assert_bit_equal(x, y, "down-sampled " + what)
    a, b = oracleapi.downsample(pts, None, None, grid), refapi.downsample(pts, None, None, grid)
    assert_bit_equal(a[0], b[0], "points only") and a[1] is None and a[2] is None


def test_estimate_normals_matches_the_compiled_reference(ref_available):
    """PointCloud::EstimateNormals: the nanoflann restatement returns the same neighbours in the same order as the compiled
    reference -- also on the raw sensor-like cloud, whose quantised depth makes many neighbour distances exactly equal -- and
    the Eigen JacobiSVD restatement is bit-exact, so every normal is bit-identical."""
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    from onepiece_b200 import scenes
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0)
    d, _, _ = scenes.room(cam, 0)
    pts = scenes.backproject(d, cam)
    jit = (pts + np.random.default_rng(0).normal(0, 1e-4, pts.shape)).astype(np.float32)
    a, (b, _) = oracleapi.estimate_normals(jit), refapi.estimate_normals(jit)
    assert_bit_equal(a, b, "normals of the tie-free cloud")
    assert np.abs(np.linalg.norm(a, axis=1) - 1).max() < 1e-5
    a, (b, _) = oracleapi.estimate_normals(pts), refapi.estimate_normals(pts)
    assert_bit_equal(a, b, "normals of the raw cloud (equal distances ordered by the tree traversal)")
    # and the full 640x480 frame (307,200 points), the size the device path is checked at against this oracle
    full = scenes.backproject(scenes.room(scenes.Camera(), 0)[0], scenes.Camera())
    a, (b, _) = oracleapi.estimate_normals(full), refapi.estimate_normals(full)
    assert_bit_equal(a, b, "normals of the full raw frame")
    # fewer than three points in range: FitPlane's warning path returns the zero vector
    far = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [10, 10, 0]], np.float32)
    assert not oracleapi.estimate_normals(far, 0.1, 30).any()
