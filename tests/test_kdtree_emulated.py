"""The DEVICE code of onepiece_b200/csrc/opb_kdtree.cu (tree build with the parallel Hoare partition, the explicit-stack walk, the
std::sort restatement, normals, FPFH) executed in the CPU container on the host-thread CUDA emulator in tests/emulate, and
compared with the oracle bit for bit.  This is a check of the kernels' logic before they reach a GPU -- not a product path: the
emulator lives under tests/ and the library itself still has no CPU route.  The -m gpu tests repeat the comparison on the device."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracleapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
p = C.c_void_p


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("kdtree_emu"))
    src = open(os.path.join(ROOT, "onepiece_b200", "csrc", "opb_kdtree.cu")).read()
    a = src.index("namespace opb\n{")
    b = src.index("} // namespace opb\n\nusing namespace opb;") + len("} // namespace opb\n")
    dev = src[a:b].replace("extern __shared__ float knn_smem[];", "")
    # the harness instantiates the building kernel with 64 threads per CTA, which keeps the emulation fast (the -m gpu tests run
    # all three CTA sizes the library picks from)
    open(os.path.join(out, "kdtree_device.inc"), "w").write('#include "opb_fitplane.cuh"\n' + dev)
    shutil.copy(os.path.join(ROOT, "onepiece_b200", "csrc", "opb_fitplane.cuh"), out)
    for f in ("cuda_emu.h", "kdtree_emu.cpp", os.path.join("stubs", "opb_common.cuh")):
        shutil.copy(os.path.join(ROOT, "tests", "emulate", f), out)
    lib = os.path.join(out, "libkdtree_emu.so")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-msse4.2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", out,
                        "-o", lib, os.path.join(out, "kdtree_emu.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    L = C.CDLL(lib)
    L.emu_build.restype = p
    L.emu_build.argtypes = [p, C.c_long]
    L.emu_destroy.argtypes = [p]
    L.emu_dump.restype = C.c_long
    L.emu_dump.argtypes = [p] * 5
    L.emu_search.argtypes = [p, p, C.c_long, C.c_int, C.c_int, C.c_float, p, p, p]
    L.emu_normals.argtypes = [p, C.c_float, C.c_int, p]
    L.emu_fpfh.argtypes = [p, p, C.c_int, C.c_float, p]
    L.emu_match.argtypes = [p, C.c_long, p, C.c_long, p]
    L.emu_match_tree.argtypes = [p, C.c_long, p, C.c_long, p]
    L.emu_ransac.argtypes = [p, p, C.c_long, C.c_int, C.c_double, C.c_uint64, p, p, p, p]
    return L


def _ptr(a):
    return a.ctypes.data_as(p)


def canonical_tree(ni, nf):
    """nodes in pre-order as (left, right, divfeat, divlow, divhigh): independent of the order the nodes were allocated in"""
    out, stack = [], [0]
    while stack:
        i = stack.pop()
        out.append((int(ni[i, 0]), int(ni[i, 1]), int(ni[i, 4]), float(nf[i, 0]), float(nf[i, 1])))
        if ni[i, 2] >= 0:
            stack.extend((int(ni[i, 3]), int(ni[i, 2])))
    return out


def clouds():
    rng = np.random.default_rng(3)
    u = rng.uniform(-1, 1, (400, 2)).astype(np.float32)
    z = (2.0 + 0.3 * np.sin(3 * u[:, 0]) * np.cos(2 * u[:, 1])).astype(np.float32)
    yield "surface", (np.stack([u[:, 0], u[:, 1], z], 1) + rng.normal(0, 0.002, (400, 3))).astype(np.float32)
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(2), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.02
    yield "lattice", g[rng.permutation(len(g))]
    d = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    yield "duplicates", np.concatenate([d, d[:150], d[:50]])
    yield "eleven", rng.uniform(-1, 1, (11, 3)).astype(np.float32)


@pytest.mark.parametrize("name,pts", list(clouds()), ids=[c[0] for c in clouds()])
def test_emulated_device_code_matches_the_oracle(emu, name, pts):
    n = len(pts)
    h = emu.emu_build(_ptr(pts), n)
    try:
        vind, ni, nf, box = np.zeros(n, np.int32), np.zeros((2 * n + 2, 5), np.int32), np.zeros((2 * n + 2, 2), np.float32), np.zeros(6, np.float32)
        m = emu.emu_dump(h, _ptr(vind), _ptr(ni), _ptr(nf), _ptr(box)) // 1000
        ov, oni, onf, obox = oracleapi.kdtree_dump(pts)
        assert np.array_equal(vind, ov), "point permutation (planeSplit)"
        assert m == len(oni) and canonical_tree(ni, nf) == canonical_tree(oni, onf) and np.array_equal(box, obox)
        rng = np.random.default_rng(5)
        qs = np.concatenate([pts[:200], rng.uniform(-1.5, 1.5, (40, 3)).astype(np.float32)])
        for mode, k, radius in [(0, 1, 0.0), (0, 30, 0.0), (2, 30, 0.01), (1, 100, 0.1), (1, 20, 0.05), (1, 60, 0.25)]:
            idx, dist, cnt = np.zeros((len(qs), k), np.int32), np.zeros((len(qs), k), np.float32), np.zeros(len(qs), np.int32)
            emu.emu_search(h, _ptr(qs), len(qs), mode, k, radius, _ptr(idx), _ptr(dist), _ptr(cnt))
            a = oracleapi.kdtree_search(pts, qs, mode, k, radius)
            assert np.array_equal(cnt, a[2]) and np.array_equal(idx, a[0]), (name, mode, k, radius)
            assert np.array_equal(dist.view(np.uint32), a[1].view(np.uint32))
        nrm = np.zeros_like(pts)
        emu.emu_normals(h, 0.1, 30, _ptr(nrm))
        on = oracleapi.estimate_normals(pts, 0.1, 30)
        assert np.array_equal(nrm.view(np.uint32), on.view(np.uint32))
        on = np.nan_to_num(on)
        # the warp-per-point FPFH kernel shuffles per neighbour, which the emulator pays with two barriers each: small cases only
        for knn, radius in {"surface": [(40, 0.25)], "lattice": [(60, 0.1)], "eleven": [(100, 0.1), (40, 0.25)]}.get(name, []):
            f = np.zeros((n, 33), np.float32)
            emu.emu_fpfh(h, _ptr(on), knn, radius, _ptr(f))
            of = oracleapi.fpfh(pts, on, knn, radius)
            assert ((f.view(np.uint32) == of.view(np.uint32)) | (np.isnan(f) & np.isnan(of))).all(), (name, knn, radius)
    finally:
        emu.emu_destroy(h)


def test_emulated_feature_matching_matches_the_oracle(emu):
    """fpfh_match_kernel (exhaustive 33-D nearest neighbour) against the oracle's KDTree<33> restatement on real descriptors"""
    from test_oracle_kdtree import _two_frames_features
    (_, fs), (_, ft) = _two_frames_features()
    fs, ft = np.ascontiguousarray(fs[:300]), np.ascontiguousarray(ft[:600])
    fs[7] = np.nan                                            # a NaN source feature matches nothing
    nearest = np.zeros(len(fs), np.int32)
    emu.emu_match(_ptr(fs), len(fs), _ptr(ft), len(ft), _ptr(nearest))
    want = oracleapi.feature_matching(fs, ft)
    got = np.stack([np.nonzero(nearest >= 0)[0], nearest[nearest >= 0]], 1).astype(np.int32)
    assert nearest[7] == -1 and np.array_equal(got, want)
    # the reference's own way -- a KDTree<33> over the targets -- with duplicate descriptors, whose tie only the tree order settles
    ft[40:60] = ft[300:320]
    ft[61] = ft[62] = ft[63]
    fs[100] = ft[305]
    fs[101] = ft[63]
    emu.emu_match_tree(_ptr(fs), len(fs), _ptr(ft), len(ft), _ptr(nearest))
    want = oracleapi.feature_matching(fs, ft)
    got = np.stack([np.nonzero(nearest >= 0)[0], nearest[nearest >= 0]], 1).astype(np.int32)
    assert nearest[7] == -1 and np.array_equal(got, want)
    emu.emu_match(_ptr(fs), len(fs), _ptr(ft), len(ft), _ptr(nearest))
    assert (nearest[nearest >= 0] != want[:, 1]).any()        # the exhaustive scan orders those ties by index instead
    ft[5] = np.nan
    emu.emu_match_tree(_ptr(fs), len(fs), _ptr(ft), len(ft), _ptr(nearest))
    assert (nearest == -2).all()                              # non-finite targets are detected (the library then scans exhaustively)


def ransac_case(n=600, outliers=0.5, seed=4):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    R *= np.sign(np.linalg.det(R))
    b = (a @ R.T + 0.3 + rng.normal(0, 0.02, (n, 3))).astype(np.float32)
    bad = rng.random(n) < outliers
    b[bad] = rng.uniform(-2, 2, (int(bad.sum()), 3)).astype(np.float32)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, 0.3
    return a, b, T, rng


def test_emulated_ransac_matches_the_oracle(emu):
    """ransac_score_kernel / ransac_winner_kernel with forced samples: winner, motion and inliers are the oracle's (which is
    pinned hypothesis by hypothesis to the compiled reference); with drawn samples the eight indices are distinct and in range"""
    a, b, T_true, rng = ransac_case()
    samples = np.stack([rng.choice(len(a), 8, replace=False) for _ in range(300)]).astype(np.int32)
    samples[17] = samples[3]                                  # a duplicate hypothesis: the earlier one must win a tie
    for thr in (0.1, 0.03):
        motion, flags, s8 = np.zeros(12, np.float32), np.zeros(len(a), np.uint8), np.zeros(8, np.int32)
        w = emu.emu_ransac(_ptr(a), _ptr(b), len(a), len(samples), thr, 0, _ptr(samples), _ptr(motion), _ptr(flags), _ptr(s8))
        ow, oT, oids = oracleapi.ransac_select(a, b, samples, thr)
        assert w == ow and np.array_equal(s8, samples[w]) and np.array_equal(np.nonzero(flags)[0], oids)
        assert np.array_equal(motion[:9].reshape(3, 3).view(np.uint32), oT[:3, :3].view(np.uint32))
        assert np.array_equal(motion[9:].view(np.uint32), oT[:3, 3].view(np.uint32))
    motion, flags, s8 = np.zeros(12, np.float32), np.zeros(len(a), np.uint8), np.zeros(8, np.int32)
    w = emu.emu_ransac(_ptr(a), _ptr(b), len(a), 400, 0.1, 12345, None, _ptr(motion), _ptr(flags), _ptr(s8))
    assert 0 <= w < 400 and len(set(s8.tolist())) == 8 and s8.min() >= 0 and s8.max() < len(a)
    oT, oflags = oracleapi.ransac_hypothesis(a, b, s8, 0.1)
    assert np.array_equal(flags, oflags) and np.array_equal(motion[:9].reshape(3, 3).view(np.uint32), oT[:3, :3].view(np.uint32))
    assert flags.sum() > 0.4 * len(a) and np.abs(motion[:9].reshape(3, 3) - T_true[:3, :3]).max() < 0.05   # it found the motion
