"""optimization::SimpleBA in the library (csrc/opb_ba.cu), checked without a GPU: the device kernel that reduces every frame
pair to 28 sums runs on the host-thread CUDA emulator against numpy, and the host half (blocks from sums, assembly, LDL^T, pose
update; host code in the product too) runs through the C-ABI test hook against the oracle, which is pinned to the compiled
reference (tests/test_oracle_ba.py).  The full opb_simple_ba call needs a device and is not run here."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracleapi
from test_oracle_ba import _graph, _run

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
p = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(p)


def _sums(poses, sid, tid, off, a, b):
    """the 28 sums per frame pair in float64 numpy (float32 transformed points, like the kernel)"""
    out = np.zeros((len(sid), 28))
    for k, (s, t) in enumerate(zip(sid, tid)):
        sl = slice(off[k], off[k + 1])
        Ps, Pt = poses[s].astype(np.float32), poses[t].astype(np.float32)
        q1 = ((a[sl] @ Ps[:3, :3].T).astype(np.float32) + Ps[:3, 3]).astype(np.float32).astype(np.float64)
        q2 = ((b[sl] @ Pt[:3, :3].T).astype(np.float32) + Pt[:3, 3]).astype(np.float32).astype(np.float64)
        iu = np.triu_indices(3)
        out[k] = np.concatenate([[len(q1)], q1.sum(0), q2.sum(0), (q1.T @ q1)[iu], (q2.T @ q2)[iu], (q1.T @ q2).reshape(-1)])
    return out


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ba_emu"))
    src = open(os.path.join(ROOT, "onepiece_b200", "csrc", "opb_ba.cu")).read()
    a = src.index("namespace opb\n{")
    b = src.index("} // namespace opb\n\nusing namespace opb;") + len("} // namespace opb\n")
    open(os.path.join(out, "ba_device.inc"), "w").write(src[a:b])
    for f in ("cuda_emu.h", "ba_emu.cpp", os.path.join("stubs", "opb_common.cuh")):
        shutil.copy(os.path.join(ROOT, "tests", "emulate", f), out)
    lib = os.path.join(out, "libba_emu.so")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-msse4.2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", out,
                        "-o", lib, os.path.join(out, "ba_emu.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    L = C.CDLL(lib)
    L.emu_ba_sums.argtypes = [p, p, p, p, C.c_int, p, p, p]
    return L


def test_emulated_pair_sums_kernel_matches_numpy(emu):
    true, start, sid, tid, off, a, b = _graph(pts_per_pair=333)
    rm = np.ascontiguousarray(start.astype(np.float32)).reshape(-1)           # row-major poses
    sums = np.zeros((len(sid), 28))
    emu.emu_ba_sums(_ptr(rm), _ptr(sid), _ptr(tid), _ptr(off), len(sid), _ptr(a), _ptr(b), _ptr(sums))
    want = _sums(start, sid, tid, off, a, b)
    assert np.array_equal(sums[:, 0], want[:, 0])
    assert np.abs(sums - want).max() <= 2e-6 * np.abs(want).max()


def test_host_half_matches_the_oracle():
    """blocks from sums + assembly + LDL^T + pose update through the C-ABI hook, iterated five times with numpy standing in for
    the kernel, against the oracle's SimpleBA"""
    from onepiece_b200 import capi
    true, start, sid, tid, off, a, b = _graph()
    poses = start.copy()
    for _ in range(5):
        P = np.ascontiguousarray(np.stack([T.T for T in poses]).astype(np.float32)).reshape(-1)
        capi.check(capi.lib.opb_simple_ba_from_sums(len(poses), _ptr(P), len(sid), _ptr(sid), _ptr(tid), _ptr(_sums(poses, sid, tid, off, a, b))))
        poses = np.stack([m.T for m in P.reshape(-1, 4, 4)])
    want = _run(oracleapi.lib(), "orc_simple_ba", start, sid, tid, off, a, b, 5)
    assert np.abs(poses - want).max() < 1e-4
    assert np.array_equal(poses[0], start[0]) and np.abs(poses - true).max() < 0.1 * np.abs(start - true).max()
    # argument checks shared with opb_simple_ba
    bad = tid.copy()
    bad[0] = 0
    with pytest.raises(capi.OpbError):
        capi.check(capi.lib.opb_simple_ba_from_sums(len(poses), _ptr(P), len(sid), _ptr(sid), _ptr(bad), _ptr(_sums(poses, sid, tid, off, a, b))))
    with pytest.raises(capi.OpbError):                            # fewer pairs than poses - 1: unconnected
        capi.check(capi.lib.opb_simple_ba_from_sums(len(poses), _ptr(P), 2, _ptr(sid), _ptr(tid), _ptr(_sums(poses, sid, tid, off, a, b))))
