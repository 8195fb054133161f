"""optimization::SimpleBA (= Optimizer::FastBA, the pose-graph refinement DenseSlam runs over its submaps: SURVEY.md §8f rank 5,
reference src/Optimization/SimpleBA.cpp:18-157) -- oracle only so far: the C restatement against the compiled reference.  The
per-pair 6x6 blocks are float sums in both (1e-5 relative); the solve is SimplicialLDLT<float> there and a dense double LDL^T
here, so refined poses are compared to 1e-4.  No device code yet: this pins what a device version will be gated against."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracleapi, refapi

p = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(p)


def _se3(x):
    from onepiece_b200 import scenes
    return scenes.se3_exp(np.asarray(x, np.float64)).astype(np.float32)


def _graph(n_poses=6, pts_per_pair=400, seed=1, noise=0.002):
    """frames observing a common cloud from poses T_i (camera-to-world); pair (s, t) holds the same world points in both frames"""
    rng = np.random.default_rng(seed)
    true = [np.eye(4, dtype=np.float32)] + [_se3(rng.normal(0, 0.15, 6)) for _ in range(n_poses - 1)]
    pairs = [(i, i + 1) for i in range(n_poses - 1)] + [(0, n_poses - 1), (1, n_poses - 2)]
    a, b, off = [], [], [0]
    for s, t in pairs:
        w = rng.uniform(-1, 1, (pts_per_pair, 3)) + [0, 0, 3]
        inv_s, inv_t = np.linalg.inv(true[s].astype(np.float64)), np.linalg.inv(true[t].astype(np.float64))
        a.append((w @ inv_s[:3, :3].T + inv_s[:3, 3] + rng.normal(0, noise, w.shape)).astype(np.float32))
        b.append((w @ inv_t[:3, :3].T + inv_t[:3, 3] + rng.normal(0, noise, w.shape)).astype(np.float32))
        off.append(off[-1] + pts_per_pair)
    start = [true[0]] + [(_se3(rng.normal(0, 0.03, 6)) @ T).astype(np.float32) for T in true[1:]]
    return (np.stack(true), np.stack(start), np.array([s for s, _ in pairs], np.int32), np.array([t for _, t in pairs], np.int32),
            np.array(off, np.int64), np.concatenate(a), np.concatenate(b))


def _run(lib, name, poses, sid, tid, off, a, b, iters):
    P = np.ascontiguousarray(np.stack([T.T for T in poses]).astype(np.float32)).reshape(-1)   # column-major per pose
    f = getattr(lib, name)
    f.argtypes = [C.c_int, p, C.c_int, p, p, p, p, p, C.c_int]
    f(len(poses), _ptr(P), len(sid), _ptr(sid), _ptr(tid), _ptr(off), _ptr(a), _ptr(b), iters)
    return np.stack([m.T for m in P.reshape(-1, 4, 4)])


def test_ba_blocks_and_refinement_match_the_compiled_reference(ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    true, start, sid, tid, off, a, b = _graph()
    R, O = refapi.lib("f32"), oracleapi.lib()
    for k in (0, 3, len(sid) - 1):
        sl = slice(off[k], off[k + 1])
        out_r, out_o = np.zeros(156, np.float32), np.zeros(156, np.float32)
        Ps, Pt = np.ascontiguousarray(start[sid[k]].T), np.ascontiguousarray(start[tid[k]].T)
        R.ref_ba_blocks.argtypes = [p, p, p, p, C.c_long, p]
        O.orc_ba_blocks.argtypes = [p, p, p, p, C.c_long, p]
        aa, bb = np.ascontiguousarray(a[sl]), np.ascontiguousarray(b[sl])
        R.ref_ba_blocks(_ptr(Ps), _ptr(Pt), _ptr(aa), _ptr(bb), len(aa), _ptr(out_r))
        O.orc_ba_blocks(_ptr(Ps), _ptr(Pt), _ptr(aa), _ptr(bb), len(aa), _ptr(out_o))
        assert np.abs(out_r - out_o).max() <= 1e-5 * np.abs(out_r).max(), k
    for iters in (1, 5):
        pr = _run(R, "ref_simple_ba", start, sid, tid, off, a, b, iters)
        po = _run(O, "orc_simple_ba", start, sid, tid, off, a, b, iters)
        assert np.abs(pr - po).max() < 1e-4, iters
        assert np.array_equal(pr[0], start[0])                       # the first pose is never touched
    # and it does what it is for: five iterations pull the perturbed poses onto the true ones
    before = np.abs(start - true).max()
    assert np.abs(po - true).max() < 0.1 * before
    # fewer than three poses: untouched (SimpleBA.cpp:84-88)
    two = _run(O, "orc_simple_ba", start[:2], sid[:1], tid[:1], off[:2], a, b, 5)
    assert np.array_equal(two, start[:2])
