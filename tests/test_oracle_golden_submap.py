"""The C oracle against tests/golden/submap_small.npz -- outputs of the compiled reference recorded by
tests/golden/gen_golden_submap.py -- for the submap back end (k-d tree searches, normals, FPFH, matching, rejection, Kabsch, RANSAC
hypotheses, SimpleBA).  Unlike tests/test_oracle_kdtree.py / test_oracle_ransac.py / test_oracle_ba.py this needs no oracle/_ref,
so the oracle stays pinned wherever the reference cannot be built."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracleapi
from test_oracle_ba import _run


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "submap_small.npz"))


def _same(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()


def test_kdtree_searches_against_the_golden_fixture(g):
    for name, mode, k, radius in (("knn30", 0, 30, 0.0), ("knnradius", 2, 30, 0.01), ("radius", 1, 100, 0.1), ("radius_capped", 1, 40, 0.25)):
        idx, dist, cnt = oracleapi.kdtree_search(g["src"], g["queries"], mode, k, radius)
        assert np.array_equal(cnt, g[f"kd_{name}_count"]) and np.array_equal(idx, g[f"kd_{name}_index"]), name
        assert _same(dist, g[f"kd_{name}_dist"]), name
    assert (g["kd_radius_capped_count"] == 40).any()


def test_normals_fpfh_matching_rejection_against_the_golden_fixture(g):
    feats = []
    for name in ("src", "tgt"):
        nrm = np.nan_to_num(oracleapi.estimate_normals(g[name], 0.1, 30))
        assert _same(nrm, g[f"normals_{name}"]), name
        f = oracleapi.fpfh(g[name], nrm, 100, 0.25)
        assert _same(f, g[f"fpfh_{name}"]), name
        feats.append(f)
    m = oracleapi.feature_matching(feats[0], feats[1])
    assert np.array_equal(m, g["matches"])
    assert np.array_equal(oracleapi.reject_matches(g["src"], g["tgt"], m, 3, 4, 0.1), g["matches_kept3"])
    assert np.array_equal(oracleapi.reject_matches(g["src"], g["tgt"], m, 1, 2, 0.05), g["matches_kept1_c2"])


def test_kabsch_ransac_and_ba_against_the_golden_fixture(g):
    kept = g["matches_kept3"]
    a, b = g["src"][kept[:, 0]], g["tgt"][kept[:, 1]]
    for s8, T, flags in zip(g["ransac_samples"], g["ransac_kabsch"], g["ransac_flags"]):
        assert _same(oracleapi.kabsch_f32(a[s8], b[s8]), T)
        To, fo = oracleapi.ransac_hypothesis(a, b, s8, 0.05)
        assert _same(To, T) and np.array_equal(fo, flags)
    refined = _run(oracleapi.lib(), "orc_simple_ba", g["ba_start"], g["ba_sid"], g["ba_tid"], g["ba_off"], g["ba_a"], g["ba_b"], 5)
    assert np.abs(refined - g["ba_refined5"]).max() < 1e-4
