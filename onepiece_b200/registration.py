"""Python mirror of one_piece::registration (reference src/Registration/ICP.h:13-26) over the C-ABI: same function
names, argument meaning and result fields as the reference, computed on the GPU by libonepiece_b200.so."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class ICPParameter:
    """registration::ICPParameter (ICP.h:13-19)"""
    max_iteration: int = 30
    threshold: float = 0.2
    scaling: float = 1.0


@dataclass
class RegistrationResult:
    """registration::RegistrationResult (RegistrationResult.h:8-16)"""
    T: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    correspondence_set_index: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    correspondence_set: tuple = (None, None)
    rmse: float = 0.0
    T_iterated: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    ok: bool = True


class PointCloud:
    """The three arrays of geometry::PointCloud that ICP reads (PointCloud.h:52-54)."""

    def __init__(self, points, normals=None):
        self.points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        self.normals = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)

    def HasNormals(self):
        return self.normals is not None and len(self.normals) == len(self.points) and len(self.points) > 0

    def EstimateNormals(self, radius: float = 0.1, knn: int = 30, device: int = 0):
        """PointCloud::EstimateNormals(radius, knn) (reference src/Geometry/PointCloud.cpp:102-144) on the GPU; fills self.normals.
        Neighbours come from the reference's own k-d tree walk (KDTree below), so equal distances are ordered as there.
        OPB_NORMALS_KNN=grid (developer knob) takes the uniform-grid search instead: same result unless distances tie."""
        self.normals = np.zeros_like(self.points)
        if os.environ.get("OPB_NORMALS_KNN", "tree") == "grid":
            capi.check(capi.lib.opb_icp_estimate_normals(_Workspace.get(device), _ptr(self.points), len(self.points), radius, knn,
                                                         _ptr(self.normals)))
            return
        tree = KDTree(device)
        tree.BuildTree(self.points)
        capi.check(capi.lib.opb_kdtree_estimate_normals(tree._h, radius, knn, _ptr(self.normals)))

    def DownSample(self, grid_len: float, colors=None, device: int = 0):
        """PointCloud::DownSample(grid_len) (reference src/Geometry/PointCloud.cpp:145-189) on the GPU -> new PointCloud
        (and the averaged colours when `colors` is given: returns (cloud, colors))."""
        col = None if colors is None else np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
        nrm = self.normals if self.HasNormals() else None
        op, oc, on = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_size_t(0)
        capi.check(capi.lib.opb_pointcloud_downsample(device, _ptr(self.points), _ptr(col), _ptr(nrm), len(self.points), grid_len,
                                                      C.byref(op), C.byref(oc) if col is not None else None,
                                                      C.byref(on) if nrm is not None else None, C.byref(n)))

        def take(p):
            if not p.value:
                return np.zeros((0, 3), np.float32)
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n.value * 3,)).reshape(-1, 3).copy()
            capi.lib.opb_free(p)
            return a
        out = PointCloud(take(op), take(on) if nrm is not None else None)
        return (out, take(oc)) if col is not None else out


class KDTree:
    """geometry::KDTree<3> (reference src/Geometry/KDTree.h:60-262) on the GPU: the same tree as the reference's nanoflann builds
    and the same walk, so every search returns the reference's neighbours in the reference's order (ties, early stop and all).
    One device workspace per (device) is shared by all instances; BuildTree replaces its contents."""
    _by_device = {}

    def __init__(self, device: int = 0):
        if device not in KDTree._by_device:
            h = C.c_void_p()
            capi.check(capi.lib.opb_kdtree_create(device, C.byref(h)))
            KDTree._by_device[device] = h
        self._h = KDTree._by_device[device]
        self.n = 0

    def BuildTree(self, points):
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        self.n = len(pts)
        capi.check(capi.lib.opb_kdtree_build(self._h, _ptr(pts), len(pts)))

    def _search(self, queries, mode, k, radius):
        qs = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
        idx = np.zeros((len(qs), k), np.int32)
        dist = np.zeros((len(qs), k), np.float32)
        cnt = np.zeros(len(qs), np.int32)
        capi.check(capi.lib.opb_kdtree_search(self._h, _ptr(qs), len(qs), mode, k, radius, _ptr(idx), _ptr(dist), _ptr(cnt)))
        return idx, dist, cnt

    def KnnSearch(self, queries, k):
        """-> (indices [nq,k] padded with -1, squared distances, result counts)"""
        return self._search(queries, 0, k, 0.0)

    def RadiusSearch(self, queries, radius, max_result):
        """`radius` bounds the SQUARED distance, as in the reference (KDTree.h:131)"""
        return self._search(queries, 1, max_result, radius)

    def KnnRadiusSearch(self, queries, k, radius):
        return self._search(queries, 2, k, radius)

    def Dump(self):
        """test hook: (vind, node ints [m,5], node floats [m,2], root box) in the device's allocation order, root = node 0"""
        m = C.c_size_t(0)
        capi.check(capi.lib.opb_kdtree_dump(self._h, None, None, None, None, C.byref(m)))
        vind = np.zeros(self.n, np.int32)
        ni = np.zeros((m.value, 5), np.int32)
        nf = np.zeros((m.value, 2), np.float32)
        box = np.zeros(6, np.float32)
        capi.check(capi.lib.opb_kdtree_dump(self._h, _ptr(vind), _ptr(ni), _ptr(nf), _ptr(box), C.byref(m)))
        return vind, ni, nf, box


def ComputeFPFHFeature(pcd: "PointCloud", knn: int = 100, radius: float = 0.1, device: int = 0):
    """registration::ComputeFPFHFeature(pcd, features, knn, radius) (reference src/Registration/3DFeature.cpp:83-131) on the GPU
    -> [n, 33] float32.  The cloud needs normals (the reference reads pcd.normals unchecked)."""
    if not pcd.HasNormals():
        raise ValueError("ComputeFPFHFeature: the point cloud has no normals")
    tree = KDTree(device)
    tree.BuildTree(pcd.points)
    out = np.zeros((len(pcd.points), 33), np.float32)
    capi.check(capi.lib.opb_kdtree_fpfh(tree._h, _ptr(pcd.normals), knn, radius, _ptr(out)))
    return out


def FeatureMatching3D(source_feature, target_feature, device: int = 0):
    """registration::FeatureMatching3D(source_feature, target_feature, matching_index) (reference
    src/Registration/GlobalRegistration.cpp:29-73) on the GPU -> [m, 2] int32 (source index, target index)"""
    sf = np.ascontiguousarray(source_feature, np.float32).reshape(-1, 33)
    tf = np.ascontiguousarray(target_feature, np.float32).reshape(-1, 33)
    pairs = np.zeros((len(sf), 2), np.int32)
    m = C.c_size_t(0)
    capi.check(capi.lib.opb_kdtree_feature_matching(KDTree(device)._h, _ptr(sf), len(sf), _ptr(tf), len(tf), _ptr(pairs), C.byref(m)))
    return pairs[:m.value].copy()


class DefaultRandomEngine:
    """std::default_random_engine as libstdc++ defines it (minstd_rand0); default-constructed state 1"""

    def __init__(self, seed: int = 1):
        self.state = C.c_uint32(seed)


def RejectMatchesRanSaPC(source_points, target_points, engine: DefaultRandomEngine, init_matches, candidate_num: int = 4,
                         difference: float = 0.1):
    """registration::RejectMatchesRanSaPC (reference src/Registration/GlobalRegistration.cpp:75-108) -> the kept matches; the
    engine advances exactly as the reference's does, so three calls on one engine reproduce RansacRegistration's three."""
    s = np.ascontiguousarray(source_points, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(target_points, np.float32).reshape(-1, 3)
    p = np.ascontiguousarray(init_matches, np.int32).reshape(-1, 2).copy()
    m = C.c_size_t(len(p))
    capi.check(capi.lib.opb_reject_matches(_ptr(s), len(s), _ptr(t), len(t), _ptr(p), C.byref(m), C.byref(engine.state), candidate_num,
                                           difference))
    return p[:m.value].copy()


def EstimateRigidTransformationRANSAC(source_points, target_points, max_iteration: int = 2000, threshold: float = 0.1, seed: int = 0,
                                      samples=None, device: int = 0):
    """geometry::EstimateRigidTransformationRANSAC(correspondence_set, inliers, inlier_ids, max_iteration, threshold) (reference
    src/Geometry/Ransac.cpp:7-41) on the GPU over the pairs source_points[i] -> target_points[i]
    -> (T 4x4 float32, inlier_ids int32 ascending, best_iteration, best_sample[8]).  `samples` ([max_iteration, 8] pair indices)
    forces the hypotheses; otherwise hypothesis h draws from a generator keyed by (seed, h)."""
    a = np.ascontiguousarray(source_points, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(target_points, np.float32).reshape(-1, 3)
    if len(a) != len(b):
        raise ValueError("source_points and target_points must pair up")
    forced = None
    if samples is not None:
        forced = np.ascontiguousarray(samples, np.int32).reshape(-1, 8)
        max_iteration = len(forced)
    T = np.zeros(16, np.float32)
    ids = np.zeros(max(len(a), 1), np.int32)
    m, it = C.c_size_t(0), C.c_int32(-1)
    s8 = np.zeros(8, np.int32)
    capi.check(capi.lib.opb_ransac_rigid_transformation(KDTree(device)._h, _ptr(a), _ptr(b), len(a), max_iteration, threshold, seed,
                                                        _ptr(forced), _ptr(T), _ptr(ids), C.byref(m), C.byref(it), _ptr(s8)))
    return T.reshape(4, 4).T.copy(), ids[:m.value].copy(), it.value, s8


@dataclass
class RANSACParameter:
    """registration::RANSACParameter (reference src/Registration/GlobalRegistration.h:12-24)"""
    max_iteration: int = 30
    threshold: float = 0.2
    scaling: float = 1.0
    max_nn: int = 100
    max_nn_normal: int = 30
    search_radius_normal: float = 0.1
    voxel_len: float = 0.1
    search_radius: float = 0.25


def RansacRegistration(source_feature_pcd: "PointCloud", target_feature_pcd: "PointCloud", source_features, target_features,
                       r_para: RANSACParameter = RANSACParameter(), seed: int = 0, device: int = 0):
    """registration::RansacRegistration(source_feature_pcd, target_feature_pcd, source_features, target_features, r_para)
    (reference src/Registration/GlobalRegistration.cpp:219-267), every step on the library: descriptor matching, three rejection
    passes on one default-seeded engine, RANSAC, ComputeRMSE -> RegistrationResult."""
    matches = FeatureMatching3D(source_features, target_features, device)
    engine = DefaultRandomEngine()
    for _ in range(3):
        matches = RejectMatchesRanSaPC(source_feature_pcd.points, target_feature_pcd.points, engine, matches)
    a = source_feature_pcd.points[matches[:, 0]]
    b = target_feature_pcd.points[matches[:, 1]]
    T, ids, _, _ = EstimateRigidTransformationRANSAC(a, b, r_para.max_iteration, r_para.threshold, seed, None, device)
    res = RegistrationResult(T=T, correspondence_set_index=matches[ids], correspondence_set=(a[ids], b[ids]), ok=len(ids) > 0)
    if len(ids):
        # ComputeRMSE (GlobalRegistration.cpp:7-15): float sum of squared norms in inlier order, sqrt(sum / count)
        e = (a[ids] @ T[:3, :3].T + T[:3, 3] - b[ids]).astype(np.float32)
        total = np.float32(0)
        for v in (e * e).sum(1, dtype=np.float32):
            total = np.float32(total + v)
        res.rmse = float(np.sqrt(total / np.float32(len(ids))))
    return res


class _Workspace:
    _by_device = {}

    @classmethod
    def get(cls, device=0):
        if device not in cls._by_device:
            h = C.c_void_p()
            capi.check(capi.lib.opb_icp_create(device, None, C.byref(h)))
            cls._by_device[device] = h
        return cls._by_device[device]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _run(source: PointCloud, target: PointCloud, init_T, icp_para: ICPParameter, plane: bool, device=0, workspace=None,
         want_pairs: bool = True):
    ws = workspace if workspace is not None else _Workspace.get(device)
    par = capi.IcpParams(icp_para.max_iteration, icp_para.threshold, icp_para.scaling)
    res = capi.IcpResult()
    T0 = np.ascontiguousarray(np.asarray(init_T, np.float32).reshape(4, 4).T).reshape(16)
    ns, nt = len(source.points), len(target.points)
    pairs = np.zeros((max(ns, 1), 2), np.int32) if want_pairs else None
    cap = ns if want_pairs else 0
    if plane:
        rc = capi.lib.opb_icp_point_to_plane(ws, _ptr(source.points), ns, _ptr(target.points),
                                             _ptr(target.normals) if target.HasNormals() else None, nt, _ptr(T0),
                                             C.byref(par), C.byref(res), _ptr(pairs), cap)
    else:
        rc = capi.lib.opb_icp_point_to_point(ws, _ptr(source.points), ns, _ptr(target.points), nt, _ptr(T0), C.byref(par),
                                             C.byref(res), _ptr(pairs), cap)
    if rc == capi.OPB_ERR_INVALID and res.status == capi.OPB_ERR_INVALID:
        # the reference prints the error and returns a default-constructed result (ICP.cpp:159-163)
        print(capi.lib.opb_last_error().decode())
        return RegistrationResult(ok=False)
    capi.check(rc)
    idx = pairs[: res.n_local_pairs].copy() if want_pairs else np.zeros((0, 2), np.int32)
    out = RegistrationResult()
    out.T = np.array(res.T[:], np.float32).reshape(4, 4).T.copy()
    out.T_iterated = np.array(res.T_iterated[:], np.float32).reshape(4, 4).T.copy()
    out.correspondence_set_index = idx
    out.correspondence_set = (source.points[idx[:, 0]], target.points[idx[:, 1]])
    out.rmse = res.rmse
    return out


class DeviceCloud:
    """geometry::PointCloud kept in HBM (opb_cloud): built on the device from a depth image with PointCloud::LoadFromDepth's
    arithmetic (reference src/Geometry/PointCloud.cpp:72-100), reused as source of one registration, target of the next and as
    the frame CubeHandler.IntegrateCloud integrates.  Loads only enqueue; `size` waits."""

    def __init__(self, device: int = 0, stream=None):
        self._h = C.c_void_p()
        self.device = device
        capi.check(capi.lib.opb_cloud_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib.opb_cloud_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        self.close()

    def LoadFromDepth(self, depth, camera, bgr=None):
        """depth [H, W] float32 metres or uint16 raw; bgr [H, W, 3] uint8 optional (kept for IntegrateCloud).  Arrays must stay
        alive until `size` (or a call that uses the cloud) has returned: they are held here."""
        depth = np.ascontiguousarray(depth)
        dt = capi.OPB_DEPTH_F32 if depth.dtype == np.float32 else (capi.OPB_DEPTH_U16 if depth.dtype == np.uint16 else -1)
        bgr = None if bgr is None else np.ascontiguousarray(bgr, np.uint8)
        self._keep = [depth, bgr]
        capi.check(capi.lib.opb_cloud_load_from_depth(self._h, _ptr(depth), dt, _ptr(bgr), camera.fx, camera.fy, camera.cx, camera.cy,
                                                      camera.width, camera.height, camera.depth_scale))
        return self

    def SetPoints(self, points):
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        self._keep = [pts]
        capi.check(capi.lib.opb_cloud_set_points(self._h, _ptr(pts), len(pts)))
        return self

    def SetNormals(self, normals):
        n = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        self._keep.append(n)
        capi.check(capi.lib.opb_cloud_set_normals(self._h, _ptr(n), len(n)))
        return self

    @property
    def size(self) -> int:
        n = C.c_size_t(0)
        capi.check(capi.lib.opb_cloud_size(self._h, C.byref(n)))
        self._keep = []
        return n.value

    def Download(self, normals: bool = False):
        n = self.size
        pts = np.zeros((n, 3), np.float32)
        nrm = np.zeros((n, 3), np.float32) if normals else None
        capi.check(capi.lib.opb_cloud_download(self._h, _ptr(pts), _ptr(nrm)))
        return (pts, nrm) if normals else pts


def _run_clouds(source: DeviceCloud, target: DeviceCloud, init_T, icp_para: ICPParameter, plane: bool, workspace=None, want_pairs: bool = True):
    ws = workspace if workspace is not None else _Workspace.get(source.device)
    par = capi.IcpParams(icp_para.max_iteration, icp_para.threshold, icp_para.scaling)
    res = capi.IcpResult()
    T0 = np.ascontiguousarray(np.asarray(init_T, np.float32).reshape(4, 4).T).reshape(16)
    ns = source.size
    pairs = np.zeros((max(ns, 1), 2), np.int32) if want_pairs else None
    fn = capi.lib.opb_icp_point_to_plane_clouds if plane else capi.lib.opb_icp_point_to_point_clouds
    rc = fn(ws, source._h, target._h, _ptr(T0), C.byref(par), C.byref(res), _ptr(pairs), ns if want_pairs else 0)
    if rc == capi.OPB_ERR_INVALID and res.status == capi.OPB_ERR_INVALID:
        print(capi.lib.opb_last_error().decode())
        return RegistrationResult(ok=False)
    capi.check(rc)
    out = RegistrationResult()
    out.T = np.array(res.T[:], np.float32).reshape(4, 4).T.copy()
    out.T_iterated = np.array(res.T_iterated[:], np.float32).reshape(4, 4).T.copy()
    out.correspondence_set_index = pairs[: res.n_local_pairs].copy() if want_pairs else np.zeros((0, 2), np.int32)
    out.rmse = res.rmse
    return out


def PointToPlaneClouds(source: DeviceCloud, target: DeviceCloud, init_T=np.eye(4), icp_para=ICPParameter(), **kw):
    """registration::PointToPlane (ICP.cpp:146-224) on device-resident clouds (the target carries its normals)"""
    return _run_clouds(source, target, init_T, icp_para, True, **kw)


def PointToPointClouds(source: DeviceCloud, target: DeviceCloud, init_T=np.eye(4), icp_para=ICPParameter(), **kw):
    """registration::PointToPoint (ICP.cpp:31-107) on device-resident clouds"""
    return _run_clouds(source, target, init_T, icp_para, False, **kw)


def PointToPlane(source, target, init_T=np.eye(4), icp_para=ICPParameter(), **kw):
    """registration::PointToPlane (ICP.cpp:146-224)"""
    return _run(source, target, init_T, icp_para, True, **kw)


def PointToPoint(source, target, init_T=np.eye(4), icp_para=ICPParameter(), **kw):
    """registration::PointToPoint (ICP.cpp:31-107)"""
    return _run(source, target, init_T, icp_para, False, **kw)


def last_nn(n, device=0):
    nn = np.zeros(n, np.int32)
    capi.check(capi.lib.opb_icp_last_nn(_Workspace.get(device), _ptr(nn), n))
    return nn


def last_prev_pose(device=0):
    """the pose the neighbours of last_nn were searched under (start_T before the last iteration's update), 4x4"""
    T = np.zeros(16, np.float32)
    capi.check(capi.lib.opb_icp_last_prev_pose(_Workspace.get(device), _ptr(T)))
    return T.reshape(4, 4).T.copy()
