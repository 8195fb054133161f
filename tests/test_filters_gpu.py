"""Depth pre-filter on the device against the oracle (bit for bit) and against the recorded OpenCV outputs (tolerance)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_bit_equal
from oracle import oracleapi

pytestmark = pytest.mark.gpu


def test_small_golden_image():
    from onepiece_b200 import imageproc
    z = np.load(os.path.join(GOLDEN, "bilateral_small.npz"))
    conv = imageproc.ConvertDepthTo32F(z["depth_u16"], 1000.0)
    assert_bit_equal(conv, z["converted"], "ConvertDepthTo32F")
    out = imageproc.BilateralFilter(conv)
    assert_bit_equal(out, oracleapi.bilateral_filter(z["converted"]), "BilateralFilter vs oracle")
    assert np.abs(out - z["cv2_plain"]).max() <= 1e-6 and np.abs(out - z["cv2_optimized"]).max() <= 3e-6
    out5 = imageproc.DepthPrefilter(160, 120).run(z["converted"], 1.0, 5, 0.05, 2.0)[1]
    assert_bit_equal(out5, oracleapi.bilateral_filter(z["converted"], 5, 0.05, 2.0), "d=5")


@pytest.mark.parametrize("scene", ["wall_f32", "room_u16"])
def test_full_frame_bit_exact(scene):
    from onepiece_b200 import imageproc, scenes
    cam = scenes.Camera()
    if scene == "wall_f32":
        d, _ = scenes.wavy_wall(cam, 3)      # float metres, 2 % zeros, 1 mm noise
        scale = 1.0
    else:
        d, _, _ = scenes.room(cam, 2)        # u16 millimetres with depth edges
        scale = cam.depth_scale
    pf = imageproc.DepthPrefilter(cam.width, cam.height)
    conv, out = pf.run(d, scale)
    oc = oracleapi.convert_depth_32f(d, scale)
    assert_bit_equal(conv, oc, "converted")
    assert_bit_equal(out, oracleapi.bilateral_filter(oc), "filtered")
    conv2, out2 = pf.run(d, scale)           # cached range table
    assert_bit_equal(out2, out, "second run")


def test_edge_cases():
    from onepiece_b200 import capi, imageproc
    pf = imageproc.DepthPrefilter(33, 9)     # ragged tile sizes
    rng = np.random.default_rng(3)
    img = (1.0 + rng.random((9, 33))).astype(np.float32)
    assert_bit_equal(pf.run(img, 1.0)[1], oracleapi.bilateral_filter(img), "ragged")
    const = np.full((9, 33), 2.5, np.float32)
    assert_bit_equal(pf.run(const, 1.0)[1], const, "constant image is copied")
    assert_bit_equal(pf.run(img, 1.0, 15, 0.5, 3.0)[1], oracleapi.bilateral_filter(img, 15, 0.5, 3.0), "largest diameter")
    with pytest.raises(capi.OpbError) as e:
        pf.run(img, 1.0, 17)
    assert e.value.code == capi.OPB_ERR_UNSUPPORTED
    with pytest.raises(capi.OpbError) as e:  # the reference exits on unknown depth types
        pf.run(img.astype(np.float64), 1.0)
    assert e.value.code == capi.OPB_ERR_UNSUPPORTED


def test_prefiltered_integration_equals_filter_then_integrate():
    """ImageSequenceIntegration.cpp:36-40: convert, filter, integrate -- fused on the device vs the oracle chain."""
    from onepiece_b200 import imageproc, scenes
    from onepiece_b200.volume import CubeHandler
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0)
    pf = imageproc.DepthPrefilter(cam.width, cam.height)
    gpu = CubeHandler(cam, 0.02, max_cubes=4096)
    ov = oracleapi.OracleVolume(cam, 0.02)
    for k in range(2):
        d, c, T = scenes.room(cam, 4 * k)
        pf.integrate(gpu, d, c, T)
        ov.integrate(oracleapi.bilateral_filter(oracleapi.convert_depth_32f(d, cam.depth_scale)), c, T)
    gi, gv = gpu.GetCubeMap()
    oi, ovx = ov.download()
    assert np.array_equal(gi, oi)
    assert_bit_equal(gv, ovx, "voxels after pre-filtered integration")
