"""optimization::SimpleBA / Optimizer::FastBA through the library on the device, against the oracle (pinned to the compiled
reference by tests/test_oracle_ba.py).  Tolerance 1e-4 on the poses: the device accumulates the per-pair sums in double, the
reference in float.  (The kernel and the host half are also covered on the CPU by tests/test_ba_cpu.py.)"""
import time

import numpy as np
import pytest

from oracle import oracleapi
from test_oracle_ba import _graph, _run

pytestmark = pytest.mark.gpu


def test_simple_ba_matches_the_oracle():
    from onepiece_b200 import capi, optimization as opt
    for n_poses, pts in ((6, 400), (24, 5000)):
        true, start, sid, tid, off, a, b = _graph(n_poses, pts, seed=n_poses)
        cs = [opt.Correspondence(int(s), int(t), a[off[k]:off[k + 1]], b[off[k]:off[k + 1]]) for k, (s, t) in enumerate(zip(sid, tid))]
        opt.Optimizer().FastBA(cs, start, 5)
        t0 = time.perf_counter()
        got = opt.Optimizer().FastBA(cs, start, 5)
        dt = time.perf_counter() - t0
        want = _run(oracleapi.lib(), "orc_simple_ba", start, sid, tid, off, a, b, 5)
        print(f"FastBA: {n_poses} poses, {len(cs)} frame pairs x {pts} point pairs, 5 iterations in {dt * 1e3:.2f} ms (host buffers)")
        assert np.abs(got - want).max() < 1e-4
        assert np.array_equal(got[0], start[0]) and np.abs(got - true).max() < 0.1 * np.abs(start - true).max()
    two = opt.SimpleBA(cs[:1], start[:2], 5)
    assert np.array_equal(two, start[:2])                         # fewer than three poses: returned as they are
    with pytest.raises(capi.OpbError):
        opt.SimpleBA(cs[:2], start, 5)                            # unconnected components
