"""Python mirror of one_piece::odometry::Odometry's dense tracking (reference src/Odometry/Odometry.h:29-37,78-105) and of
geometry::RGBDFrame (src/Geometry/RGBDFrame.h:11-67) over the C-ABI: same names, argument meaning and result fields as
the reference, computed on the GPU by libonepiece_b200.so.  No CPU path."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _depth_type(depth):
    if depth.dtype == np.float32:
        return capi.OPB_DEPTH_F32
    if depth.dtype == np.uint16:
        return capi.OPB_DEPTH_U16
    return -1  # the library reports the reference's "Unknown depth image type"


@dataclass
class DenseTrackingResult:
    """odometry::DenseTrackingResult (Odometry.h:29-37)"""
    T: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    pixel_correspondence_set: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.uint32))
    correspondence_set: np.ndarray = field(default_factory=lambda: np.zeros((0, 2, 3), np.float32))
    rmse: float = 1e6
    tracking_success: bool = False
    # per-iteration trace (not in the reference's result; used by the parity tests)
    iterations: int = 0
    corr_per_iteration: np.ndarray = None
    T_per_iteration: np.ndarray = None


class RGBDFrame:
    """geometry::RGBDFrame(rgb, depth): raw images now, the dense-tracking cache on first use (mutated by DenseTracking)."""

    def __init__(self, odometry: "Odometry", rgb, depth, frame_id=-1):
        self.odometry = odometry
        self.frame_id = frame_id
        self.rgb = np.ascontiguousarray(rgb, np.uint8)
        self.depth = np.ascontiguousarray(depth)
        h = C.c_void_p()
        capi.check(capi.lib.opb_frame_create(odometry.handle, _ptr(self.rgb), _ptr(self.depth), _depth_type(self.depth), C.byref(h)))
        self.handle = h

    def IsPreprocessedDense(self):
        return bool(capi.lib.opb_frame_is_preprocessed(self.handle))

    def preprocess(self):
        capi.check(capi.lib.opb_frame_preprocess(self.odometry.handle, self.handle))
        return self

    def image(self, what, level):
        """what: 0 gray, 1 depth32f, 2 gray dx, 3 gray dy, 4 depth dx, 5 depth dy"""
        o = self.odometry
        out = np.zeros((o.height >> level, o.width >> level), np.float32)
        capi.check(capi.lib.opb_frame_image(o.handle, self.handle, what, level, _ptr(out)))
        return out

    def Release(self):
        if getattr(self, "handle", None):
            capi.lib.opb_frame_destroy(self.handle)
            self.handle = None

    __del__ = Release


class Odometry:
    """odometry::Odometry(camera) restricted to the dense path: SetMultiScale, DenseTracking (both overloads)."""

    def __init__(self, camera, device=0, levels=3, iterations=(4, 8, 16)):
        self.camera = camera
        self.device = device
        self.width, self.height = camera.width, camera.height
        self.handle = None
        self._create(levels, list(iterations))

    def _create(self, levels, iterations):
        if self.handle:
            capi.lib.opb_odometry_destroy(self.handle)
            self.handle = None
        d = capi.OdometryDesc()
        capi.lib.opb_odometry_desc_default(C.byref(d))
        c = self.camera
        d.fx, d.fy, d.cx, d.cy, d.width, d.height, d.depth_scale = c.fx, c.fy, c.cx, c.cy, c.width, c.height, c.depth_scale
        d.levels = levels
        for i in range(capi.OPB_ODO_MAX_LEVELS):
            d.iterations[i] = iterations[i] if i < len(iterations) else 0
        d.device = self.device
        h = C.c_void_p()
        capi.check(capi.lib.opb_odometry_create(C.byref(d), C.byref(h)))
        self.handle = h
        self.levels, self.iterations = levels, iterations[:levels]

    def SetMultiScale(self, layer_count):
        """Odometry::SetMultiScale (Odometry.h:101-105): iter_count_per_level.resize(layer_count, 4)"""
        it = (self.iterations + [4] * layer_count)[:layer_count]
        self._create(layer_count, it)

    def Frame(self, rgb, depth, frame_id=-1):
        return RGBDFrame(self, rgb, depth, frame_id)

    def set_profiling(self, on=True):
        capi.check(capi.lib.opb_odometry_set_profiling(self.handle, int(on)))

    def SetLoopForm(self, form: int):
        """1 (default): second persistent solver loop; 2: first persistent form; 0: one launch pair per iteration
        (include/onepiece_b200.h, opb_odometry_set_loop_form)"""
        capi.check(capi.lib.opb_odometry_set_loop_form(self.handle, int(form)))

    def last_tracking_ms(self):
        ms, tail = C.c_float(0), C.c_float(0)
        capi.check(capi.lib.opb_odometry_last_timing(self.handle, C.byref(ms), C.byref(tail)))
        self.last_solve_tail_us = tail.value
        return ms.value

    def _result(self, res, pairs, xyz):
        n = res.n_correspondences
        k = min(res.iterations, capi.OPB_ODO_MAX_TRACE)
        return DenseTrackingResult(
            T=np.array(res.T[:], np.float32).reshape(4, 4).T.copy(),
            pixel_correspondence_set=pairs[:n].copy() if pairs is not None else None,
            correspondence_set=xyz[:n].reshape(-1, 2, 3).copy() if xyz is not None else None,
            rmse=res.rmse, tracking_success=bool(res.tracking_success), iterations=res.iterations,
            corr_per_iteration=np.array(res.corr_per_iteration[:k], np.int64),
            T_per_iteration=np.stack([np.array(res.T_per_iteration[i][:], np.float32).reshape(4, 4).T for i in range(k)])
            if k else np.zeros((0, 4, 4), np.float32))

    def DenseTracking(self, *args, term_type=0, want_correspondences=True):
        """DenseTracking(source_frame, target_frame, initial_T[, term_type]) (Odometry.cpp:526-608) or
        DenseTracking(source_color, target_color, source_depth, target_depth, initial_T[, term_type]) (:463-523)"""
        npx = self.width * self.height
        pairs = np.zeros((npx, 4), np.uint32) if want_correspondences else None
        xyz = np.zeros((npx, 6), np.float32) if want_correspondences else None
        res = capi.TrackingResult()
        if isinstance(args[0], RGBDFrame):
            src, tgt = args[0], args[1]
            init_T = args[2] if len(args) > 2 else np.eye(4)
            if len(args) > 3:
                term_type = args[3]
            T0 = np.ascontiguousarray(np.asarray(init_T, np.float32).reshape(4, 4).T).reshape(16)
            capi.check(capi.lib.opb_odometry_dense_tracking_frames(self.handle, src.handle, tgt.handle, _ptr(T0), term_type,
                                                                   C.byref(res), _ptr(pairs), npx if want_correspondences else 0,
                                                                   _ptr(xyz)))
        else:
            sc, tc, sd, td = args[:4]
            init_T = args[4] if len(args) > 4 else np.eye(4)
            if len(args) > 5:
                term_type = args[5]
            sc, tc = np.ascontiguousarray(sc, np.uint8), np.ascontiguousarray(tc, np.uint8)
            sd, td = np.ascontiguousarray(sd), np.ascontiguousarray(td)
            T0 = np.ascontiguousarray(np.asarray(init_T, np.float32).reshape(4, 4).T).reshape(16)
            capi.check(capi.lib.opb_odometry_dense_tracking(self.handle, _ptr(sc), _ptr(tc), _ptr(sd), _ptr(td), _depth_type(sd), _ptr(T0),
                                                            term_type, C.byref(res), _ptr(pairs), npx if want_correspondences else 0,
                                                            _ptr(xyz)))
        return self._result(res, pairs, xyz)

    def single_iteration(self, source: RGBDFrame, target: RGBDFrame, level, T, term_type=0):
        """one DoSingleIteration* at a pyramid level from pose T (teacher forcing) -> dict(T, JTJ, JTr, r2, pairs)"""
        Tcm = np.ascontiguousarray(np.asarray(T, np.float32).reshape(4, 4).T).reshape(16).copy()
        sums = np.zeros(43)
        cap = (self.width >> level) * (self.height >> level)
        pairs = np.zeros((cap, 4), np.uint32)
        n = C.c_size_t(0)
        capi.check(capi.lib.opb_odometry_single_iteration(self.handle, source.handle, target.handle, level, _ptr(Tcm), term_type,
                                                          _ptr(sums), _ptr(pairs), cap, C.byref(n)))
        return dict(T=Tcm.reshape(4, 4).T.astype(np.float64), JTJ=sums[:36].reshape(6, 6).copy(), JTr=sums[36:42].copy(), r2=sums[42],
                    pairs=pairs[: n.value].copy())

    def close(self):
        if getattr(self, "handle", None):
            capi.lib.opb_odometry_destroy(self.handle)
            self.handle = None

    __del__ = close
