"""Python mirror of the two one_piece::tool functions on the fusion path (reference src/Tool/ImageProcessing.h:19-20,
ImageProcessing.cpp:64-91) over the C-ABI: same names and argument meaning, computed on the GPU by libonepiece_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .volume import _ptr, depth_type_of, pose_colmajor


class DepthPrefilter:
    """Device workspace for one image size (opb_prefilter)."""

    def __init__(self, width: int, height: int, device: int = 0, stream=None):
        self.width, self.height = width, height
        self._h = C.c_void_p()
        capi.check(capi.lib.opb_prefilter_create(device, C.c_void_p(stream) if stream else None, width, height, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            capi.lib.opb_prefilter_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, depth, depth_scale: float, d: int = 7, sigma_color: float = 0.03, sigma_space: float = 4.5):
        """-> (converted, filtered) float32 images: tool::ConvertDepthTo32F then tool::BilateralFilter."""
        depth = np.ascontiguousarray(depth)
        assert depth.shape == (self.height, self.width), depth.shape
        conv = np.zeros(depth.shape, np.float32)
        out = np.zeros(depth.shape, np.float32)
        capi.check(capi.lib.opb_prefilter_run(self._h, _ptr(depth), depth_type_of(depth), depth_scale, d, sigma_color, sigma_space,
                                              _ptr(conv), _ptr(out)))
        return conv, out

    def integrate(self, volume, depth, rgb, pose, d: int = 7, sigma_color: float = 0.03, sigma_space: float = 4.5):
        """ConvertDepthTo32F + BilateralFilter + volume.IntegrateImage(filtered_depth, rgb, pose), the sequence of
        example/ImageSequenceIntegration.cpp:36-40, with the filtered image staying on the device."""
        depth = np.ascontiguousarray(depth)
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        p = pose_colmajor(pose)
        capi.check(capi.lib.opb_volume_integrate_prefiltered(volume._h, self._h, _ptr(depth), depth_type_of(depth), _ptr(rgb), _ptr(p),
                                                             d, sigma_color, sigma_space))


_shared = {}


def _workspace(shape, device):
    key = (shape, device)
    if key not in _shared:
        _shared[key] = DepthPrefilter(shape[1], shape[0], device)
    return _shared[key]


def ConvertDepthTo32F(depth, depth_scale: float, device: int = 0):
    """tool::ConvertDepthTo32F(depth, refined_depth, depth_scale) (ImageProcessing.cpp:68-91)"""
    depth = np.ascontiguousarray(depth)
    return _workspace(depth.shape, device).run(depth, depth_scale)[0]


def BilateralFilter(source, range: int = 7, device: int = 0):  # noqa: A002  (the reference's parameter name)
    """tool::BilateralFilter(source, target, range = 7) = cv::bilateralFilter(source, target, range, 0.03, 4.5)"""
    source = np.ascontiguousarray(source, np.float32)
    return _workspace(source.shape, device).run(source, 1.0, range)[1]
