"""Generates tests/golden/submap_small.npz: outputs of the COMPILED REFERENCE (oracle/_ref, built from /root/reference by
oracle/Makefile) for the submap back end on small stored inputs, so that the C oracle stays pinned where
the reference cannot be built: geometry::KDTree<3> searches, PointCloud::EstimateNormals, registration::ComputeFPFHFeature,
FeatureMatching3D, RejectMatchesRanSaPC, geometry::EstimateRigidTransformation, the RANSAC hypothesis evaluation and
optimization::SimpleBA.  Inputs are stored too, so nothing depends on the host's libm.  Run in the build container:

    python tests/golden/gen_golden_submap.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refapi  # noqa: E402
from test_oracle_ba import _graph, _run  # noqa: E402

rng = np.random.default_rng(42)
p = C.c_void_p


def ptr(a):
    return a.ctypes.data_as(p)


def surface(n):
    u = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    z = (2.0 + 0.3 * np.sin(3 * u[:, 0]) * np.cos(2 * u[:, 1])).astype(np.float32)
    return (np.stack([u[:, 0], u[:, 1], z], 1) + rng.normal(0, 0.002, (n, 3))).astype(np.float32)


out = {}
# two overlapping clouds, with a lattice patch (tied distances) and duplicated points in the first
src = surface(900)
lattice = np.stack(np.meshgrid(np.arange(8), np.arange(8), indexing="ij"), -1).reshape(-1, 2).astype(np.float32) * 0.03
src = np.concatenate([src, np.concatenate([lattice - 0.5, np.full((64, 1), 2.0, np.float32)], 1), src[:20]]).astype(np.float32)
T = np.eye(4)
T[:3, :3] = np.linalg.qr(np.eye(3) + 0.05 * rng.normal(size=(3, 3)))[0]
T[:3, 3] = [0.05, -0.02, 0.03]
tgt = (src[rng.permutation(len(src))[:850]] @ T[:3, :3].T + T[:3, 3] + rng.normal(0, 0.001, (850, 3))).astype(np.float32)
out["src"], out["tgt"] = src, tgt
queries = np.concatenate([src[:150], rng.uniform(-1.5, 1.5, (30, 3)).astype(np.float32)])
out["queries"] = queries
for name, mode, k, radius in (("knn30", 0, 30, 0.0), ("knnradius", 2, 30, 0.01), ("radius", 1, 100, 0.1), ("radius_capped", 1, 40, 0.25)):
    idx, dist, cnt = refapi.kdtree_search(src, queries, mode, k, radius)
    out[f"kd_{name}_index"], out[f"kd_{name}_dist"], out[f"kd_{name}_count"] = idx, dist, cnt
feats = []
for name, cloud in (("src", src), ("tgt", tgt)):
    nrm, _ = refapi.estimate_normals(cloud, 0.1, 30)
    nrm = np.nan_to_num(nrm)
    f, _ = refapi.fpfh(cloud, nrm, 100, 0.25)
    out[f"normals_{name}"], out[f"fpfh_{name}"] = nrm, f
    feats.append(f)
m = refapi.feature_matching(feats[0], feats[1])
out["matches"] = m
out["matches_kept3"] = refapi.reject_matches(src, tgt, m, 3, 4, 0.1)
out["matches_kept1_c2"] = refapi.reject_matches(src, tgt, m, 1, 2, 0.05)
# Kabsch and RANSAC hypotheses on the kept matches
kept = out["matches_kept3"]
a, b = np.ascontiguousarray(src[kept[:, 0]]), np.ascontiguousarray(tgt[kept[:, 1]])
samples = np.stack([rng.choice(len(a), 8, replace=False) for _ in range(40)]).astype(np.int32)
L = refapi.lib("f32")
L.ref_kabsch.argtypes = [p, p, C.c_long, p]
kab = []
for s8 in samples:
    Tk = np.zeros(16, np.float64)
    sa, sb = np.ascontiguousarray(a[s8]), np.ascontiguousarray(b[s8])
    L.ref_kabsch(ptr(sa), ptr(sb), 8, ptr(Tk))
    kab.append(Tk.reshape(4, 4).T.astype(np.float32))
out["ransac_samples"], out["ransac_kabsch"] = samples, np.stack(kab)
out["ransac_flags"] = np.stack([refapi.ransac_hypothesis(a, b, s8, 0.05)[1] for s8 in samples])
# SimpleBA
true, start, sid, tid, off, pa, pb = _graph(5, 120, seed=9)
out["ba_start"], out["ba_sid"], out["ba_tid"], out["ba_off"], out["ba_a"], out["ba_b"] = start, sid, tid, off, pa, pb
out["ba_refined5"] = _run(L, "ref_simple_ba", start, sid, tid, off, pa, pb, 5)
path = os.path.join(ROOT, "tests", "golden", "submap_small.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path) // 1024, "KiB;", len(src), len(tgt), "points,", len(m), "matches ->", len(kept), "kept")
