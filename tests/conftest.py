import os
import sys

import numpy as np
import pytest

# Several ICP workspaces of ONE process exchange packets in tests/test_fusion_gpu.py (a kernel of one spins until another's has
# run): every stream needs its own hardware queue, or a kernel can be queued behind the very kernel that waits for it.  Real
# deployments run one process per GPU.  (Read by the CUDA driver when the context is created, i.e. after this line.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _gpu_available() -> bool:
    try:
        from onepiece_b200 import capi
        return capi.lib.opb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a device must fail loudly, not skip: the product has no CPU path
    pass


@pytest.fixture(scope="session")
def ref_available():
    from oracle import refapi
    return refapi.available("f32")


def bits(a):
    a = np.ascontiguousarray(a, np.float32)
    return a.view(np.uint32)


def assert_bit_equal(a, b, what="", nan_payload=True):
    """nan_payload=False: NaNs compare equal whatever their sign/payload bits (x86 SSE and the GPU canonicalise differently)"""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not nan_payload:
        a = np.where(np.isnan(a), np.float32(np.nan), a).astype(np.float32)
        b = np.where(np.isnan(b), np.float32(np.nan), b).astype(np.float32)
    same = a.view(np.uint32) == b.view(np.uint32)
    if not same.all():
        idx = np.argwhere(~same)
        first = tuple(idx[0])
        raise AssertionError(f"{what}: {len(idx)} of {same.size} floats differ bitwise; first at {first}: "
                             f"{a[first]!r} vs {b[first]!r}")


def canon_triangles(points, colors):
    """Order-independent form of a 3-vertices-per-triangle mesh: rows of 18 floats sorted lexicographically."""
    p = np.ascontiguousarray(points, np.float32).reshape(-1, 9)
    c = np.ascontiguousarray(colors, np.float32).reshape(-1, 9)
    a = np.concatenate([p, c], 1)
    return a[np.lexsort(a.T[::-1])]


def sha(a) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
