"""The nanoflann kd-tree restatement (build, the three searches the reference uses, libstdc++'s std::sort tie order) and the
FPFH restatement on top of it, pinned to the compiled reference (oracle/_ref): same neighbours, same order, same bits.
SURVEY.md §8f rank 5; reference: src/Geometry/KDTree.h:60-262, 3rdparty/nanoflann/include/nanoflann.hpp,
src/Registration/3DFeature.cpp:7-131."""
import numpy as np
import pytest

from oracle import oracleapi, refapi


def _clouds():
    rng = np.random.default_rng(3)
    yield "random", rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), np.arange(3), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.02
    yield "lattice", g[rng.permutation(len(g))]          # every distance tied many times over
    d = rng.uniform(-1, 1, (2000, 3)).astype(np.float32)
    yield "duplicates", np.concatenate([d, d[:500], d[:100]])
    p = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    p[:, 2] = 0.5
    yield "plane", p                                       # one query of this cloud drives std::sort into its heap-sort fallback
    yield "tiny", rng.uniform(-1, 1, (7, 3)).astype(np.float32)


def surface(n, seed=5, noise=0.002):
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    z = (2.0 + 0.3 * np.sin(3 * u[:, 0]) * np.cos(2 * u[:, 1])).astype(np.float32)
    return (np.stack([u[:, 0], u[:, 1], z], 1) + rng.normal(0, noise, (n, 3))).astype(np.float32)


SEARCHES = [(oracleapi.KD_KNN, 1, 0.0), (oracleapi.KD_KNN, 30, 0.0), (oracleapi.KD_KNN_RADIUS, 30, 0.01),
            (oracleapi.KD_RADIUS, 100, 0.1), (oracleapi.KD_RADIUS, 20, 0.05), (oracleapi.KD_RADIUS, 100, 0.25)]


@pytest.mark.parametrize("name,pts", list(_clouds()), ids=[c[0] for c in _clouds()])
def test_kdtree_searches_match_the_compiled_reference(ref_available, name, pts):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    qs = np.concatenate([pts, rng.uniform(-1.5, 1.5, (200, 3)).astype(np.float32)])   # members and outside points
    for mode, k, radius in SEARCHES:
        a = oracleapi.kdtree_search(pts, qs, mode, k, radius)
        b = refapi.kdtree_search(pts, qs, mode, k, radius)
        assert np.array_equal(a[2], b[2]), (name, mode, k, radius)
        assert np.array_equal(a[0], b[0]), (name, mode, k, radius, np.nonzero((a[0] != b[0]).any(1))[0][:5])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    # the radius search's early stop really is exercised: more than 2.5 k points lie inside the radius somewhere
    if name == "random":
        d2 = ((pts[None, :200] - pts[:, None]) ** 2).sum(-1)
        assert ((d2 < 0.25).sum(0) > 250).any()


def test_kdtree_dump_is_a_consistent_tree():
    pts = surface(3000)
    vind, ni, nf, box = oracleapi.kdtree_dump(pts)
    assert sorted(vind.tolist()) == list(range(len(pts)))
    assert np.array_equal(box[:3], pts.min(0)) and np.array_equal(box[3:], pts.max(0))
    leaves = ni[ni[:, 2] < 0]
    assert (leaves[:, 1] - leaves[:, 0] <= 10).all() and (leaves[:, 1] - leaves[:, 0]).sum() == len(pts)
    for left, right, c1, c2, feat in ni[ni[:, 2] >= 0][:50]:
        assert ni[c1, 0] == left and ni[c1, 1] == ni[c2, 0] and ni[c2, 1] == right
        i = int(np.nonzero((ni[:, 0] == left) & (ni[:, 1] == right) & (ni[:, 2] == c1))[0][0])
        assert pts[vind[left:ni[c1, 1]], feat].max() == nf[i, 0] and pts[vind[ni[c2, 0]:right], feat].min() == nf[i, 1]


@pytest.mark.parametrize("case", ["surface", "capped", "sparse", "lattice"])
def test_fpfh_matches_the_compiled_reference(ref_available, case):
    """ComputeFPFHFeature bit for bit: below the radius search's 2.5 k cap, above it (DenseSlam's own parameters put ~300 points
    in the ball, so the traversal order decides which 250 are seen), isolated points (0 * inf = NaN rows, kept), tied distances."""
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    if case == "lattice":
        g = np.stack(np.meshgrid(np.arange(50), np.arange(50), indexing="ij"), -1).reshape(-1, 2).astype(np.float32) * 0.04
        pts, knn, radius = np.concatenate([g, np.full((len(g), 1), 2.0, np.float32)], 1), 50, 0.05
    else:
        pts, knn, radius = {"surface": (surface(4000), 100, 0.1), "capped": (surface(4000), 100, 0.25),
                            "sparse": (surface(300), 100, 0.01)}[case]
    normals = np.nan_to_num(oracleapi.estimate_normals(pts, 0.1, 30))
    a, (b, _) = oracleapi.fpfh(pts, normals, knn, radius), refapi.fpfh(pts, normals, knn, radius)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{np.count_nonzero(~same.all(1))} of {len(pts)} descriptors differ"
    if case == "sparse":
        assert np.isnan(a).any()
    else:
        assert np.isfinite(a).all() and (a >= 0).all() and a.any(1).all()


def _two_frames_features():
    """two 320x240 views of the S2 room the way DenseSlam's submap back end prepares them: DownSample(0.05), normals (0.1, 30),
    FPFH (100, 0.25)"""
    from onepiece_b200 import scenes
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 2, c0.fy / 2, c0.cx / 2, c0.cy / 2, 320, 240, 1000.0)
    out = []
    for k in (0, 12):
        d, _, _ = scenes.room(cam, k)
        down = oracleapi.downsample(scenes.backproject(d, cam), None, None, 0.05)[0]
        nrm = np.nan_to_num(oracleapi.estimate_normals(down, 0.1, 30))
        out.append((down, oracleapi.fpfh(down, nrm, 100, 0.25)))
    return out


def test_feature_matching_and_rejection_match_the_compiled_reference(ref_available):
    """FeatureMatching3D (KDTree<33>, k = 1) and RejectMatchesRanSaPC (std::default_random_engine + uniform_int_distribution as
    libstdc++ implements them) index for index -- also with duplicate features (ties) and with NaN target rows, which poison
    nanoflann's boxes (std::min / std::max keep their first argument against NaN) and make the reference prune real neighbours."""
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    (ps, fs), (pt, ft) = _two_frames_features()
    a, b = oracleapi.feature_matching(fs, ft), refapi.feature_matching(fs, ft)
    assert len(a) == len(fs) and np.array_equal(a, b)
    brute = ((fs[:500, None, :] - ft[None, :, :]) ** 2).sum(-1).argmin(1)
    assert (brute == a[:500, 1]).mean() > 0.999          # it is the nearest neighbour (float64 brute force, up to rounding)
    s2, t2 = fs.copy(), ft.copy()
    s2[5] = np.nan
    t2[0] = np.nan
    t2[77] = np.nan
    t2[100:110] = t2[200:210]
    s2[300] = t2[205]
    a2, b2 = oracleapi.feature_matching(s2, t2), refapi.feature_matching(s2, t2)
    assert len(a2) == len(fs) - 1 and np.array_equal(a2, b2)
    assert (a2[:, 1] != a[np.isin(a[:, 0], a2[:, 0]), 1]).mean() > 0.01   # the NaN rows really do change what the reference returns
    for rounds, cand, diff in [(1, 4, 0.1), (3, 4, 0.1), (3, 2, 0.05), (2, 8, 0.01)]:
        x, y = oracleapi.reject_matches(ps, pt, a, rounds, cand, diff), refapi.reject_matches(ps, pt, a, rounds, cand, diff)
        assert 0 < len(x) < len(a) and np.array_equal(x, y), (rounds, cand, diff)
    assert len(oracleapi.reject_matches(ps, pt, a[:0])) == 0
