"""The C-ABI library loads and exports every symbol include/onepiece_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "onepiece_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(opb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    names = declared_functions()
    assert "opb_volume_integrate" in names and "opb_volume_extract_mesh" in names and len(names) >= 20


def test_library_exports_every_declared_symbol():
    from onepiece_b200 import capi
    missing = [n for n in declared_functions() if not hasattr(capi.lib, n)]
    assert not missing, f"declared in the header but not exported by the .so: {missing}"


def test_python_binding_covers_the_header():
    from onepiece_b200 import capi
    unbound = [n for n in declared_functions() if n not in capi.SIGNATURES]
    assert not unbound, f"no ctypes signature for: {unbound}"


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the create call must fail loudly with OPB_ERR_CUDA, not compute on the CPU."""
    from onepiece_b200 import capi
    if capi.lib.opb_device_count() > 0:
        pytest.skip("a GPU is present")
    d = capi.VolumeDesc()
    capi.lib.opb_volume_desc_default(C.byref(d))
    h = C.c_void_p()
    rc = capi.lib.opb_volume_create(C.byref(d), C.byref(h))
    assert rc == capi.OPB_ERR_CUDA and not h.value
    assert b"no CPU path" in capi.lib.opb_last_error()


def test_product_never_imports_the_oracle():
    """Only tests/, bench.py and __graft_entry__.py may touch oracle/: neither the package nor the developer scripts do."""
    offenders = []
    for top in ("onepiece_b200", "scripts", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            if os.path.basename(dirpath) in ("build", "__pycache__"):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".sh")):
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"(from|import)\s+oracle\b|oracle/|opb_oracle|libopref", src):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_desc_defaults_match_reference_defaults():
    from onepiece_b200 import capi
    d = capi.VolumeDesc()
    capi.lib.opb_volume_desc_default(C.byref(d))
    # CubeHandler.h:363-364, VoxelCube.h:27, Integrator.h:24, Camera.h:94-105
    assert (d.voxel_resolution, d.truncation, d.near_plane, d.far_plane) == pytest.approx((0.01, 0.1, 0.5, 5.0))
    assert (d.width, d.height, d.depth_scale) == (640, 480, 1000.0)
    assert d.fx == pytest.approx(514.817) and d.cy == pytest.approx(238.447)
