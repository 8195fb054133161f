"""geometry::KDTree<3>, PointCloud::EstimateNormals and registration::ComputeFPFHFeature on the device
against the oracle, which is itself pinned bit for bit to the compiled reference (tests/test_oracle_kdtree.py): same tree, same
neighbours in the same order, same descriptors."""
import time

import numpy as np
import pytest

from oracle import oracleapi
from test_kdtree_emulated import canonical_tree

pytestmark = pytest.mark.gpu


def _reg():
    from onepiece_b200 import registration as reg
    return reg


def _surface(n, seed=5, noise=0.002):
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    z = (2.0 + 0.3 * np.sin(3 * u[:, 0]) * np.cos(2 * u[:, 1])).astype(np.float32)
    return (np.stack([u[:, 0], u[:, 1], z], 1) + rng.normal(0, noise, (n, 3))).astype(np.float32)


def _room_cloud(div):
    from onepiece_b200 import scenes
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / div, c0.fy / div, c0.cx / div, c0.cy / div, 640 // div, 480 // div, 1000.0)
    d, _, _ = scenes.room(cam, 0)
    return scenes.backproject(d, cam)


def _clouds():
    rng = np.random.default_rng(3)
    yield "random-100k", rng.uniform(-1, 1, (100000, 3)).astype(np.float32)
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), np.arange(3), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.02
    yield "lattice", g[rng.permutation(len(g))]
    d = rng.uniform(-1, 1, (2000, 3)).astype(np.float32)
    yield "duplicates", np.concatenate([d, d[:500], d[:100]])
    p = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    p[:, 2] = 0.5
    yield "plane", p
    yield "sensor-160x120", _room_cloud(4)
    yield "tiny", rng.uniform(-1, 1, (7, 3)).astype(np.float32)
    yield "eleven", rng.uniform(-1, 1, (11, 3)).astype(np.float32)


@pytest.mark.parametrize("name,pts", list(_clouds()), ids=[c[0] for c in _clouds()])
def test_device_tree_and_searches_are_the_references(name, pts):
    reg = _reg()
    tree = reg.KDTree()
    tree.BuildTree(pts)
    vind, ni, nf, box = tree.Dump()
    ov, oni, onf, obox = oracleapi.kdtree_dump(pts)
    assert np.array_equal(vind, ov), "point permutation (planeSplit)"
    assert len(ni) == len(oni) and canonical_tree(ni, nf) == canonical_tree(oni, onf) and np.array_equal(box, obox)
    rng = np.random.default_rng(11)
    qs = np.concatenate([pts[:3000], rng.uniform(-1.5, 1.5, (300, 3)).astype(np.float32)])
    for search, args, mode in [(tree.KnnSearch, (1,), 0), (tree.KnnSearch, (30,), 0), (tree.KnnSearch, (64,), 0),
                               (tree.KnnRadiusSearch, (30, 0.01), 2), (tree.RadiusSearch, (0.1, 100), 1),
                               (tree.RadiusSearch, (0.05, 20), 1), (tree.RadiusSearch, (0.25, 100), 1)]:
        idx, dist, cnt = search(qs, *args)
        k = args[0] if mode != 1 else args[1]
        radius = 0.0 if mode == 0 else (args[1] if mode == 2 else args[0])
        a = oracleapi.kdtree_search(pts, qs, mode, k, radius)
        assert np.array_equal(cnt, a[2]) and np.array_equal(idx, a[0]), (name, mode, args, np.nonzero((idx != a[0]).any(1))[0][:5])
        assert np.array_equal(dist.view(np.uint32), a[1].view(np.uint32))


def test_kdtree_argument_errors():
    from onepiece_b200 import capi
    reg = _reg()
    tree = reg.KDTree()
    tree.BuildTree(np.zeros((0, 3), np.float32))
    idx, dist, cnt = tree.KnnSearch(np.zeros((4, 3), np.float32), 3)
    assert (cnt == 0).all() and (idx == -1).all()
    tree.BuildTree(_surface(100))
    with pytest.raises(capi.OpbError):
        tree.KnnSearch(np.zeros((1, 3), np.float32), 65)
    with pytest.raises(capi.OpbError):
        tree.RadiusSearch(np.zeros((1, 3), np.float32), 0.1, 500)
    with pytest.raises(capi.OpbError):
        tree.KnnSearch(np.zeros((1, 3), np.float32), 0)
    with pytest.raises(ValueError):
        reg.ComputeFPFHFeature(reg.PointCloud(_surface(10)))


@pytest.mark.parametrize("case", ["surface", "capped", "sparse", "lattice", "denseslam"])
def test_fpfh_matches_the_oracle_bit_for_bit(case):
    """ComputeFPFHFeature: below and above the radius search's 2.5 k early stop, isolated points (NaN rows), tied distances, and
    the way DenseSlam calls it (DownSample(0.05) of a 640x480 frame, normals (0.1, 30), FPFH (100, 0.25): DenseSlam.h:49-56)."""
    reg = _reg()
    if case == "lattice":
        g = np.stack(np.meshgrid(np.arange(50), np.arange(50), indexing="ij"), -1).reshape(-1, 2).astype(np.float32) * 0.04
        pts, knn, radius = np.concatenate([g, np.full((len(g), 1), 2.0, np.float32)], 1), 50, 0.05
    elif case == "denseslam":
        pts, knn, radius = reg.PointCloud(_room_cloud(1)).DownSample(0.05).points, 100, 0.25
    else:
        pts, knn, radius = {"surface": (_surface(20000), 100, 0.1), "capped": (_surface(20000), 100, 0.25),
                            "sparse": (_surface(300), 100, 0.01)}[case]
    pc = reg.PointCloud(pts)
    pc.EstimateNormals(0.1, 30)
    on = oracleapi.estimate_normals(pts, 0.1, 30)
    assert np.array_equal(pc.normals.view(np.uint32), on.view(np.uint32)), "normals"
    pc.normals = np.nan_to_num(pc.normals)
    t0 = time.perf_counter()
    f = reg.ComputeFPFHFeature(pc, knn, radius)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    of = oracleapi.fpfh(pts, pc.normals, knn, radius)
    dt_o = time.perf_counter() - t0
    same = (f.view(np.uint32) == of.view(np.uint32)) | (np.isnan(f) & np.isnan(of))
    print(f"fpfh[{case}]: {len(pts)} points, device {dt * 1e3:.1f} ms (tree build + search + features, host buffers), oracle {dt_o * 1e3:.0f} ms")
    assert same.all(), f"{np.count_nonzero(~same.all(1))} of {len(pts)} descriptors differ, first at {np.argmax(~same.all(1))}"
    assert np.isnan(f).any() == (case == "sparse")


def test_estimate_normals_full_frame_matches_the_oracle():
    """640x480 raw sensor-like cloud (quantised depth: many exactly equal neighbour distances): 307,200 normals, all bit-identical"""
    reg = _reg()
    pts = _room_cloud(1)
    pc = reg.PointCloud(pts)
    pc.EstimateNormals()
    t0 = time.perf_counter()
    pc.EstimateNormals()
    dt = time.perf_counter() - t0
    on = oracleapi.estimate_normals(pts)
    same = (pc.normals.view(np.uint32) == on.view(np.uint32)).all(1)
    print(f"EstimateNormals (k-d tree order): {len(pts)} points in {dt * 1e3:.1f} ms")
    assert same.all(), f"{np.count_nonzero(~same)} of {len(pts)} normals differ"


def test_feature_matching_and_rejection_match_the_oracle():
    """FeatureMatching3D on the device and RejectMatchesRanSaPC in the library's host code, chained the way RansacRegistration
    chains them (GlobalRegistration.cpp:232-240), index for index against the oracle (itself pinned to the compiled reference)."""
    from test_oracle_kdtree import _two_frames_features
    reg = _reg()
    (ps, fs), (pt, ft) = _two_frames_features()
    t0 = time.perf_counter()
    m = reg.FeatureMatching3D(fs, ft)
    dt = time.perf_counter() - t0
    want = oracleapi.feature_matching(fs, ft)
    print(f"FeatureMatching3D: {len(fs)} x {len(ft)} descriptors in {dt * 1e3:.2f} ms (host buffers)")
    assert np.array_equal(m, want)
    engine = reg.DefaultRandomEngine()
    kept = m
    for _ in range(3):
        kept = reg.RejectMatchesRanSaPC(ps, pt, engine, kept)
    assert np.array_equal(kept, oracleapi.reject_matches(ps, pt, want, 3)) and 0 < len(kept) < len(m)
    fs2, ft2 = fs.copy(), ft.copy()
    fs2[5] = np.nan                       # a NaN source matches nothing
    ft2[100:140] = ft2[900:940]           # duplicate descriptors: exact ties, settled by the tree's visiting order
    ft2[61] = ft2[62] = ft2[63]
    fs2[300], fs2[301] = ft2[905], ft2[63]
    m2 = reg.FeatureMatching3D(fs2, ft2)
    assert len(m2) == len(fs) - 1 and np.array_equal(m2, oracleapi.feature_matching(fs2, ft2))
    # a NaN target row: the reference's tree is poisoned and returns wrong neighbours; the device scans exhaustively instead
    ft2[7] = np.nan
    m3 = reg.FeatureMatching3D(fs2, ft2)
    finite = np.ones(len(ft2), bool)
    finite[7] = False
    probe = m3[:400]
    d = ((fs2[probe[:, 0], None, :].astype(np.float64) - ft2[None, finite, :]) ** 2).sum(-1)
    best = ((fs2[probe[:, 0]].astype(np.float64) - ft2[probe[:, 1]]) ** 2).sum(-1)
    assert len(m3) == len(fs) - 1 and (best <= d.min(1) * (1 + 1e-6)).all()
    assert len(reg.FeatureMatching3D(fs[:0], ft)) == 0 and len(reg.FeatureMatching3D(fs, ft[:0])) == 0


def test_ransac_with_forced_samples_is_the_oracles_and_registration_recovers_the_motion():
    """EstimateRigidTransformationRANSAC on the device: with the hypotheses forced, the winner, its motion and its inliers are the
    oracle's bit for bit (the oracle is pinned hypothesis by hypothesis to the compiled reference); with its own sampler the whole
    RansacRegistration chain recovers the true camera motion between two views."""
    from test_kdtree_emulated import ransac_case
    from test_oracle_kdtree import _two_frames_features
    from onepiece_b200 import capi, scenes
    reg = _reg()
    a, b, T_true, rng = ransac_case(3000)
    samples = np.stack([rng.choice(len(a), 8, replace=False) for _ in range(4000)]).astype(np.int32)
    samples[170] = samples[30]
    for thr in (0.1, 0.03):
        T, ids, w, s8 = reg.EstimateRigidTransformationRANSAC(a, b, threshold=thr, samples=samples)
        ow, oT, oids = oracleapi.ransac_select(a, b, samples, thr)
        assert w == ow and np.array_equal(s8, samples[w]) and np.array_equal(ids, oids)
        assert np.array_equal(T.view(np.uint32), oT.view(np.uint32))
    t0 = time.perf_counter()
    T, ids, w, s8 = reg.EstimateRigidTransformationRANSAC(a, b, 40000, 0.1, seed=7)
    dt = time.perf_counter() - t0
    print(f"EstimateRigidTransformationRANSAC: 40,000 hypotheses x {len(a)} pairs in {dt * 1e3:.1f} ms (host buffers)")
    oT, oflags = oracleapi.ransac_hypothesis(a, b, s8, 0.1)
    assert len(set(s8.tolist())) == 8 and np.array_equal(T.view(np.uint32), oT.view(np.uint32)) and np.array_equal(ids, np.nonzero(oflags)[0])
    assert len(ids) > 0.45 * len(a) and np.abs(T - T_true).max() < 0.05
    T2, ids2, w2, _ = reg.EstimateRigidTransformationRANSAC(a, b, 40000, 0.1, seed=7)
    assert w2 == w and np.array_equal(T2, T)                      # same seed, same answer
    # the reference's corner cases
    T0, ids0, _, _ = reg.EstimateRigidTransformationRANSAC(a[:5], b[:5], 10, 0.1)
    assert not T0.any() and len(ids0) == 0                        # fewer than eight pairs: the zero matrix
    with pytest.raises(capi.OpbError):
        reg.EstimateRigidTransformationRANSAC(a[:8], b[:8], 10, 0.1)
    # the whole chain on two views of the room (DenseSlam's parameters)
    (ps, fs), (pt, ft) = _two_frames_features()
    para = reg.RANSACParameter(max_iteration=40000, threshold=0.1)
    t0 = time.perf_counter()
    res = reg.RansacRegistration(reg.PointCloud(ps), reg.PointCloud(pt), fs, ft, para, seed=1)
    dt = time.perf_counter() - t0
    T_true = np.linalg.inv(scenes.room_pose(12)) @ scenes.room_pose(0)     # source = frame 0, target = frame 12
    print(f"RansacRegistration: {len(ps)} / {len(pt)} points, {len(res.correspondence_set_index)} inliers, rmse {res.rmse:.4f}, {dt * 1e3:.1f} ms")
    assert np.linalg.norm(res.T[:3, 3] - T_true[:3, 3]) < 0.1 and np.abs(res.T[:3, :3] - T_true[:3, :3]).max() < 0.1
