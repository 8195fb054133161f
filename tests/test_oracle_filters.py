"""The oracle's depth pre-filter (tool::ConvertDepthTo32F + tool::BilateralFilter = cv::bilateralFilter) against outputs of
the real OpenCV function recorded in tests/golden/bilateral_small.npz (tests/golden/gen_golden_filters.py).  OpenCV's own two
code paths differ from each other by 2.9e-6 m on this image, so the pin is a tolerance: 1e-6 m to the plain path, 3e-6 m to
the dispatched SIMD path."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import oracleapi


def _golden():
    return np.load(os.path.join(GOLDEN, "bilateral_small.npz"))


def test_convert_depth_matches_opencv_input_bit_for_bit():
    z = _golden()
    c = oracleapi.convert_depth_32f(z["depth_u16"], 1000.0)
    assert np.array_equal(c.view(np.uint32), z["converted"].view(np.uint32))
    f = oracleapi.convert_depth_32f(z["converted"], 1000.0)           # CV_32FC1 input: plain copy
    assert np.array_equal(f.view(np.uint32), z["converted"].view(np.uint32))


def test_bilateral_against_recorded_opencv_outputs():
    z = _golden()
    o = oracleapi.bilateral_filter(z["converted"], 7, 0.03, 4.5)
    assert np.abs(o - z["cv2_plain"]).max() <= 1e-6
    assert np.abs(o - z["cv2_optimized"]).max() <= 3e-6
    assert np.abs(z["cv2_plain"] - z["cv2_optimized"]).max() > 1e-6     # the two OpenCV paths themselves disagree
    o5 = oracleapi.bilateral_filter(z["converted"], 5, 0.05, 2.0)
    assert np.abs(o5 - z["cv2_plain_d5"]).max() <= 1e-6


def test_bilateral_properties():
    z = _golden()
    src = z["converted"]
    o = oracleapi.bilateral_filter(src)
    assert (o[src == 0] == 0).all()                                     # holes stay holes: neighbours > 4 sigma away weigh 0
    assert np.abs(o - src)[src > 0].max() < 0.05                        # edge preserving: nothing moves by more than the range scale
    const = np.full((12, 16), 1.25, np.float32)
    assert np.array_equal(oracleapi.bilateral_filter(const), const)     # constant image is copied
