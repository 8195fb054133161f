"""Generates tests/golden/bilateral_small.npz: cv2.bilateralFilter outputs (the OpenCV call behind the reference's
tool::BilateralFilter, src/Tool/ImageProcessing.cpp:64-67) on a small synthetic depth image, from the two code paths of
the cv2 build available in the build container (dispatched SIMD and cv2.setUseOptimized(False)).  They differ from each
other in the last bits, which is why the oracle is pinned to them by tolerance.  Run in the build container:

    python tests/golden/gen_golden_filters.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from onepiece_b200 import scenes  # noqa: E402

c0 = scenes.Camera()
cam = scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0)
d16, _, _ = scenes.room(cam, 3)
rng = np.random.default_rng(7)
d16 = d16.copy()
d16[rng.random(d16.shape) < 0.02] = 0                      # sensor holes
d16 = (d16.astype(np.int32) + rng.integers(-3, 4, d16.shape) * (d16 > 0)).clip(0, 65535).astype(np.uint16)  # mm noise
src = d16.astype(np.float32) / np.float32(1000.0)          # tool::ConvertDepthTo32F
cv2.setNumThreads(1)
cv2.setUseOptimized(True)
opt = cv2.bilateralFilter(src, 7, 0.03, 4.5)
cv2.setUseOptimized(False)
plain = cv2.bilateralFilter(src, 7, 0.03, 4.5)
d5 = cv2.bilateralFilter(src, 5, 0.05, 2.0)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bilateral_small.npz"), depth_u16=d16, converted=src, cv2_optimized=opt,
                    cv2_plain=plain, cv2_plain_d5=d5, cv2_version=cv2.__version__)
print("cv2", cv2.__version__, "paths differ by", float(np.abs(opt - plain).max()))
