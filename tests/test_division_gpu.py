"""The voxel-update kernel divides several numerators by one divisor with a shared refined reciprocal
(opb_volume.cu, quotient_by).  This must equal IEEE division bit for bit on tame operands."""
import ctypes as C

import pytest

from onepiece_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
def test_shared_reciprocal_quotient_equals_div_rn(mode):
    bad = C.c_uint64(123)
    capi.check(capi.lib.opb_selftest_quotient(0, 4_000_000_000, 12345 + mode, mode, C.byref(bad)))
    assert bad.value == 0
