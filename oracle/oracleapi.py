"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/libopb_oracle.so, the plain-C restatement of the
reference algorithm (oracle/opb_oracle.c).  Imported by tests/, bench.py's cpu_baseline / --impl reference legs
and __graft_entry__.smoke() only; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libopb_oracle.so")
_LIB = None

c_f, c_d, c_i, c_l, c_p = C.c_float, C.c_double, C.c_int, C.c_long, C.c_void_p


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("opb_oracle.c", "opb_oracle.h")]
    src.append(os.path.join(_HERE, "..", "onepiece_b200", "csrc", "opb_mc_table.h"))
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "oracle", "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    build()
    L = C.CDLL(_LIB_PATH)
    L.orc_volume_create.restype = c_p
    L.orc_volume_create.argtypes = [c_f] * 4 + [c_i, c_i] + [c_f] * 5
    L.orc_volume_destroy.argtypes = [c_p]
    L.orc_volume_clear.argtypes = [c_p]
    L.orc_pose_inverse.argtypes = [c_p, c_p]
    L.orc_frustum_planes.argtypes = [c_p, c_p, c_p]
    L.orc_frustum_contains.restype = c_i
    L.orc_frustum_contains.argtypes = [c_p, c_f, c_f, c_f]
    L.orc_volume_bounding.argtypes = [c_p, c_p, c_i, c_p, c_p, c_p]
    L.orc_volume_get_sdf.argtypes = [c_p, c_p, c_i, c_p, c_p, c_l, c_p]
    L.orc_volume_prepare_cubes.restype = c_l
    L.orc_volume_prepare_cubes.argtypes = [c_p, c_p, c_i, c_p, c_p, c_l]
    L.orc_volume_integrate.restype = c_l
    L.orc_volume_integrate.argtypes = [c_p, c_p, c_i, c_p, c_p]
    L.orc_volume_num_cubes.restype = c_l
    L.orc_volume_num_cubes.argtypes = [c_p]
    L.orc_volume_download.argtypes = [c_p, c_p, c_p]
    L.orc_volume_upload.argtypes = [c_p, c_p, c_p, c_l]
    L.orc_volume_extract_mesh.restype = c_l
    L.orc_volume_extract_mesh.argtypes = [c_p, C.POINTER(c_p), C.POINTER(c_p)]
    L.orc_marching_cube_cell.restype = c_i
    L.orc_marching_cube_cell.argtypes = [c_p] * 5
    L.orc_free.argtypes = [c_p]
    L.orc_nearest.argtypes = [c_p, c_l, c_p, c_l, c_p]
    L.orc_icp.restype = c_l
    L.orc_icp.argtypes = [c_p, c_l, c_p, c_p, c_l, c_p, c_i, c_d, c_d, c_p, c_p, c_p, c_p]
    L.orc_se3_exp.argtypes = [c_p, c_p]
    L.orc_gray_u8.argtypes = [c_p, c_i, c_p]
    L.orc_blur3.argtypes = [c_p, c_i, c_i, c_p]
    L.orc_pyr_down.argtypes = [c_p, c_i, c_i, c_p]
    L.orc_sobel3.argtypes = [c_p, c_i, c_i, c_i, c_p]
    L.orc_frame_create.restype = c_p
    L.orc_frame_create.argtypes = [c_p, c_p, c_i, c_i, c_i]
    L.orc_frame_destroy.argtypes = [c_p]
    L.orc_frame_image.restype = c_l
    L.orc_frame_image.argtypes = [c_p, c_i, c_i, c_p]
    L.orc_dense_tracking_frames.argtypes = [c_p, c_p] + [c_f] * 5 + [c_p, c_i, c_p, c_p, c_l]
    L.orc_dense_tracking.argtypes = [c_p, c_p, c_p, c_p, c_i, c_i, c_i] + [c_f] * 5 + [c_p, c_i, c_p, c_p, c_l]
    L.orc_correspondences.restype = c_l
    L.orc_correspondences.argtypes = [c_p, c_p, c_i, c_i] + [c_f] * 4 + [c_p, c_p, c_l]
    L.orc_single_iteration.restype = c_l
    L.orc_single_iteration.argtypes = [c_p, c_p, c_i] + [c_f] * 4 + [c_p, c_i, c_p, c_p, c_l]
    L.orc_frame_preprocess.argtypes = [c_p, c_f]
    L.orc_clustering_simplify.restype = c_i
    L.orc_clustering_simplify.argtypes = [c_p, c_p, c_p, c_p, c_p, c_f]
    L.orc_compute_normals.argtypes = [c_p, c_l, c_p, c_l, c_p]
    L.orc_estimate_normals.argtypes = [c_p, c_l, c_f, c_i, c_p]
    L.orc_kdtree_search.argtypes = [c_p, c_l, c_p, c_l, c_i, c_i, c_f, c_l, c_p, c_p, c_p]
    L.orc_kdtree_dump.restype = c_l
    L.orc_kdtree_dump.argtypes = [c_p, c_l, c_p, c_p, c_p, c_p]
    L.orc_fpfh.restype = c_i
    L.orc_fpfh.argtypes = [c_p, c_p, c_l, c_i, c_f, c_p]
    L.orc_feature_matching.restype = c_l
    L.orc_feature_matching.argtypes = [c_p, c_l, c_p, c_l, c_p]
    L.orc_reject_matches.restype = c_l
    L.orc_reject_matches.argtypes = [c_p, c_p, c_p, c_l, c_i, c_i, c_f]
    L.orc_kabsch_f32.argtypes = [c_p, c_p, c_l, c_p]
    L.orc_ransac_hypothesis.restype = c_l
    L.orc_ransac_hypothesis.argtypes = [c_p, c_p, c_l, c_p, c_d, c_p, c_p]
    L.orc_ransac_select.restype = c_l
    L.orc_ransac_select.argtypes = [c_p, c_p, c_l, c_p, c_l, c_d, c_p, c_p, c_p]
    L.orc_downsample.restype = c_l
    L.orc_downsample.argtypes = [c_p, c_p, c_p, c_l, c_f, c_p, c_p, c_p]
    L.orc_volume_transform.restype = c_p
    L.orc_volume_transform.argtypes = [c_p, c_p, c_i, c_f]
    L.orc_volume_merge.restype = c_i
    L.orc_volume_merge.argtypes = [c_p, c_p]
    L.orc_volume_resolution.restype = c_f
    L.orc_volume_resolution.argtypes = [c_p]
    L.orc_convert_depth_32f.argtypes = [c_p, c_i, c_l, c_f, c_p]
    L.orc_bilateral_filter.argtypes = [c_p, c_i, c_i, c_i, c_d, c_d, c_p]
    _LIB = L
    return L


def _ptr(a):
    return a.ctypes.data_as(c_p) if a is not None else None


def _pose_cm(pose):
    return np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(4, 4).T).reshape(16)


def pose_inverse(pose):
    p = _pose_cm(pose)
    out = np.zeros(16, np.float32)
    lib().orc_pose_inverse(_ptr(p), _ptr(out))
    return out.reshape(4, 4).T.copy()


class OracleVolume:
    """CPU restatement of one_piece::integration::CubeHandler."""

    def __init__(self, cam, voxel_resolution=0.01, truncation=0.1, near=0.5, far=5.0):
        self.L = lib()
        self.cam = cam
        self.h = self.L.orc_volume_create(cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.depth_scale,
                                          voxel_resolution, truncation, near, far)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_volume_destroy(self.h)
            self.h = None

    def clear(self):
        self.L.orc_volume_clear(self.h)

    def frustum(self, pose, points):
        p = _pose_cm(pose)
        planes = np.zeros(24, np.float32)
        self.L.orc_frustum_planes(self.h, _ptr(p), _ptr(planes))
        pts = np.ascontiguousarray(points, np.float32)
        mask = np.array([self.L.orc_frustum_contains(_ptr(planes), float(a), float(b), float(c)) for a, b, c in pts], bool)
        return planes.reshape(6, 4), mask

    def bounding(self, depth, pose):
        depth = np.ascontiguousarray(depth)
        p = _pose_cm(pose)
        mx = np.zeros(3, np.float32)
        mn = np.zeros(3, np.float32)
        self.L.orc_volume_bounding(self.h, _ptr(depth), int(depth.dtype == np.uint16), _ptr(p), _ptr(mx), _ptr(mn))
        return mx, mn

    def get_sdf(self, depth, pose, points):
        depth = np.ascontiguousarray(depth)
        pts = np.ascontiguousarray(points, np.float32)
        out = np.zeros(len(pts), np.float32)
        p = _pose_cm(pose)
        self.L.orc_volume_get_sdf(self.h, _ptr(depth), int(depth.dtype == np.uint16), _ptr(p), _ptr(pts), len(pts), _ptr(out))
        return out

    def prepare_cubes(self, depth, pose):
        depth = np.ascontiguousarray(depth)
        p = _pose_cm(pose)
        cap = 1 << 22
        ids = np.zeros((cap, 3), np.int32)
        n = self.L.orc_volume_prepare_cubes(self.h, _ptr(depth), int(depth.dtype == np.uint16), _ptr(p), _ptr(ids), cap)
        return ids[:n].copy()

    def integrate(self, depth, bgr, pose) -> int:
        depth = np.ascontiguousarray(depth)
        bgr = np.ascontiguousarray(bgr, np.uint8)
        p = _pose_cm(pose)
        return self.L.orc_volume_integrate(self.h, _ptr(depth), int(depth.dtype == np.uint16), _ptr(bgr), _ptr(p))

    def num_cubes(self) -> int:
        return self.L.orc_volume_num_cubes(self.h)

    def download(self):
        n = self.num_cubes()
        ids = np.zeros((n, 3), np.int32)
        vox = np.zeros((n, 512, 5), np.float32)
        self.L.orc_volume_download(self.h, _ptr(ids), _ptr(vox))
        order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))
        return ids[order], vox[order]

    def upload(self, ids, vox):
        ids = np.ascontiguousarray(ids, np.int32)
        vox = np.ascontiguousarray(vox, np.float32)
        self.L.orc_volume_upload(self.h, _ptr(ids), _ptr(vox), len(ids))

    def transform(self, trans, nearest: bool, alloc_res=None):
        """CubeHandler::Transform / TransformNearest -> new OracleVolume.  alloc_res None = what the reference does:
        the source resolution for Transform, CubePara's default 0.01 for TransformNearest (it forgets to copy c_para)."""
        if alloc_res is None:
            alloc_res = 0.01 if nearest else self.L.orc_volume_resolution(self.h)
        p = _pose_cm(trans)
        o = OracleVolume.__new__(OracleVolume)
        o.L, o.cam = self.L, self.cam
        o.h = self.L.orc_volume_transform(self.h, _ptr(p), int(nearest), alloc_res)
        return o

    def merge(self, other) -> int:
        return self.L.orc_volume_merge(self.h, other.h)

    def resolution(self) -> float:
        return self.L.orc_volume_resolution(self.h)

    def extract_mesh(self):
        """-> (points [nv,3], colors [nv,3]); triangle i = vertices 3i..3i+2"""
        xyz, rgb = c_p(), c_p()
        n = self.L.orc_volume_extract_mesh(self.h, C.byref(xyz), C.byref(rgb))
        pts = np.ctypeslib.as_array(C.cast(xyz, C.POINTER(c_f)), shape=(max(n, 1) * 3,))[: n * 3].reshape(n, 3).copy()
        col = np.ctypeslib.as_array(C.cast(rgb, C.POINTER(c_f)), shape=(max(n, 1) * 3,))[: n * 3].reshape(n, 3).copy()
        self.L.orc_free(xyz)
        self.L.orc_free(rgb)
        return pts, col


def marching_cube_cell(corners, sdf, colors):
    corners = np.ascontiguousarray(corners, np.float32)
    sdf = np.ascontiguousarray(sdf, np.float32)
    colors = np.ascontiguousarray(colors, np.float32)
    xyz = np.zeros((15, 3), np.float32)
    rgb = np.zeros((15, 3), np.float32)
    n = lib().orc_marching_cube_cell(_ptr(corners), _ptr(sdf), _ptr(colors), _ptr(xyz), _ptr(rgb))
    return xyz[:n], rgb[:n]


def nearest(query, target):
    q = np.ascontiguousarray(query, np.float32)
    t = np.ascontiguousarray(target, np.float32)
    nn = np.zeros(len(q), np.int32)
    lib().orc_nearest(_ptr(q), len(q), _ptr(t), len(t), _ptr(nn))
    return nn


def icp(src, tgt, tgt_normals, init_T, max_iter=30, threshold=0.2, scaling=1.0):
    """CPU restatement of registration::PointToPlane / PointToPoint -> dict(T, T_iterated, pairs, rmse) or None for
    the reference's error path."""
    src = np.ascontiguousarray(src, np.float32)
    tgt = np.ascontiguousarray(tgt, np.float32)
    nrm = np.ascontiguousarray(tgt_normals, np.float32) if tgt_normals is not None else None
    T0 = _pose_cm(init_T)
    T = np.zeros(16)
    Ti = np.zeros(16)
    pairs = np.zeros((max(len(src), 1), 2), np.int32)
    rmse = c_d(0)
    n = lib().orc_icp(_ptr(src), len(src), _ptr(tgt), _ptr(nrm), len(tgt), _ptr(T0), max_iter, threshold, scaling, _ptr(T),
                      _ptr(Ti), _ptr(pairs), C.byref(rmse))
    if n < 0:
        return None
    return dict(T=T.reshape(4, 4).T.copy(), T_iterated=Ti.reshape(4, 4).T.copy(), pairs=pairs[:n].copy(), rmse=rmse.value)


def se3_exp(x):
    x = np.ascontiguousarray(x, np.float64)
    T = np.zeros(16)
    lib().orc_se3_exp(_ptr(x), _ptr(T))
    return T.reshape(4, 4).T.copy()


class TrackingResult(C.Structure):
    _fields_ = [("T", c_d * 16), ("rmse", c_d), ("tracking_success", c_i), ("n_correspondences", c_l), ("iterations", c_i),
                ("corr_per_iteration", c_l * 64), ("T_per_iteration", (c_d * 16) * 64)]


def _tracking_dict(r, pairs):
    k = min(r.iterations, 64)
    return dict(T=np.array(r.T[:]).reshape(4, 4).T.copy(), rmse=r.rmse, success=bool(r.tracking_success),
                pairs=pairs[: r.n_correspondences].copy(), corr_per_iteration=np.array(r.corr_per_iteration[:k]),
                T_per_iteration=np.stack([np.array(r.T_per_iteration[i][:]).reshape(4, 4).T for i in range(k)]) if k else None)


def gray_u8(bgr):
    b = np.ascontiguousarray(bgr, np.uint8)
    out = np.zeros(b.shape[:2], np.uint8)
    lib().orc_gray_u8(_ptr(b), out.size, _ptr(out))
    return out


def blur3(img):
    a = np.ascontiguousarray(img, np.float32)
    out = np.zeros_like(a)
    lib().orc_blur3(_ptr(a), a.shape[1], a.shape[0], _ptr(out))
    return out


def pyr_down(img):
    a = np.ascontiguousarray(img, np.float32)
    out = np.zeros((a.shape[0] // 2, a.shape[1] // 2), np.float32)
    lib().orc_pyr_down(_ptr(a), a.shape[1], a.shape[0], _ptr(out))
    return out


def sobel3(img, dx):
    a = np.ascontiguousarray(img, np.float32)
    out = np.zeros_like(a)
    lib().orc_sobel3(_ptr(a), a.shape[1], a.shape[0], int(dx), _ptr(out))
    return out


class OracleFrame:
    """CPU restatement of geometry::RGBDFrame's dense-tracking cache (RGBDFrame.h:33-44)."""

    def __init__(self, bgr, depth):
        self.L = lib()
        b = np.ascontiguousarray(bgr, np.uint8)
        d = np.ascontiguousarray(depth)
        self.h_, self.w_ = d.shape
        self.h = self.L.orc_frame_create(_ptr(b), _ptr(d), int(d.dtype == np.uint16), self.w_, self.h_)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_frame_destroy(self.h)
            self.h = None

    def image(self, what, level):
        out = np.zeros((self.h_ >> level, self.w_ >> level), np.float32)
        self.L.orc_frame_image(self.h, what, level, _ptr(out))
        return out

    def images(self):
        return [[self.image(a, l) for l in range(3)] for a in range(6)]

    def preprocess(self, depth_scale=1000.0):
        self.L.orc_frame_preprocess(self.h, depth_scale)
        return self


def dense_tracking_frames(source: OracleFrame, target: OracleFrame, cam, init_T=np.eye(4), term_type=0):
    r = TrackingResult()
    pairs = np.zeros((source.w_ * source.h_, 4), np.uint32)
    T0 = _pose_cm(init_T)
    lib().orc_dense_tracking_frames(source.h, target.h, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_scale, _ptr(T0), term_type,
                                    C.byref(r), _ptr(pairs), len(pairs))
    return _tracking_dict(r, pairs)


def dense_tracking(src_bgr, tgt_bgr, src_depth, tgt_depth, cam, init_T=np.eye(4), term_type=0):
    r = TrackingResult()
    sb, tb = np.ascontiguousarray(src_bgr, np.uint8), np.ascontiguousarray(tgt_bgr, np.uint8)
    sd, td = np.ascontiguousarray(src_depth), np.ascontiguousarray(tgt_depth)
    h, w = sd.shape
    pairs = np.zeros((w * h, 4), np.uint32)
    T0 = _pose_cm(init_T)
    lib().orc_dense_tracking(_ptr(sb), _ptr(tb), _ptr(sd), _ptr(td), int(sd.dtype == np.uint16), w, h, cam.fx, cam.fy, cam.cx, cam.cy,
                             cam.depth_scale, _ptr(T0), term_type, C.byref(r), _ptr(pairs), len(pairs))
    return _tracking_dict(r, pairs)


def correspondences(src_depth, tgt_depth, cam, T):
    sd = np.ascontiguousarray(src_depth, np.float32)
    td = np.ascontiguousarray(tgt_depth, np.float32)
    h, w = sd.shape
    Tcm = _pose_cm(T)
    pairs = np.zeros((w * h, 4), np.uint32)
    n = lib().orc_correspondences(_ptr(sd), _ptr(td), w, h, cam.fx, cam.fy, cam.cx, cam.cy, _ptr(Tcm), _ptr(pairs), w * h)
    return pairs[:n].copy()


def single_iteration(source: OracleFrame, target: OracleFrame, cam, level, T, term_type=0):
    """One teacher-forced solver iteration at a pyramid level -> dict(T, JTJ, JTr, r2, pairs)"""
    Tcm = _pose_cm(T)
    sums = np.zeros(43)
    pairs = np.zeros(((source.w_ >> level) * (source.h_ >> level), 4), np.uint32)
    n = lib().orc_single_iteration(source.h, target.h, level, cam.fx, cam.fy, cam.cx, cam.cy, _ptr(Tcm), term_type, _ptr(sums),
                                   _ptr(pairs), len(pairs))
    return dict(T=Tcm.reshape(4, 4).T.astype(np.float64), JTJ=sums[:36].reshape(6, 6).copy(), JTr=sums[36:42].copy(), r2=sums[42],
                pairs=pairs[:n].copy())


def convert_depth_32f(depth, depth_scale):
    """tool::ConvertDepthTo32F"""
    depth = np.ascontiguousarray(depth)
    out = np.zeros(depth.shape, np.float32)
    lib().orc_convert_depth_32f(_ptr(depth), int(depth.dtype == np.uint16), depth.size, depth_scale, _ptr(out))
    return out


def bilateral_filter(src, d=7, sigma_color=0.03, sigma_space=4.5):
    """tool::BilateralFilter(source, target, range = 7) = cv::bilateralFilter(source, target, range, 0.03, 4.5)"""
    src = np.ascontiguousarray(src, np.float32)
    out = np.zeros_like(src)
    lib().orc_bilateral_filter(_ptr(src), src.shape[1], src.shape[0], d, sigma_color, sigma_space, _ptr(out))
    return out


def clustering_simplify(points, colors, triangles, grid_len):
    """TriangleMesh::ClusteringSimplify(grid_len) -> (points, colors, triangles) of the simplified mesh, or None for the
    reference's error path (grid_len <= 0)."""
    pts = np.array(points, np.float32, copy=True).reshape(-1, 3)
    col = None if colors is None else np.array(colors, np.float32, copy=True).reshape(-1, 3)
    tri = np.array(triangles, np.uint32, copy=True).reshape(-1, 3)
    nv, nt = c_l(len(pts)), c_l(len(tri))
    rc = lib().orc_clustering_simplify(_ptr(pts), _ptr(col), C.byref(nv), _ptr(tri), C.byref(nt), grid_len)
    if rc != 0:
        return None
    return pts[: nv.value].copy(), (None if col is None else col[: nv.value].copy()), tri[: nt.value].copy()


def compute_normals(points, triangles):
    """TriangleMesh::ComputeNormals"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    tri = np.ascontiguousarray(triangles, np.uint32).reshape(-1, 3)
    out = np.zeros_like(pts)
    lib().orc_compute_normals(_ptr(pts), len(pts), _ptr(tri), len(tri), _ptr(out))
    return out


def downsample(points, colors, normals, grid_len):
    """PointCloud::DownSample(grid_len) -> (points, colors or None, normals or None)"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    col = None if colors is None else np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
    nrm = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    op = np.zeros_like(pts)
    oc = None if col is None else np.zeros_like(pts)
    on = None if nrm is None else np.zeros_like(pts)
    n = lib().orc_downsample(_ptr(pts), _ptr(col), _ptr(nrm), len(pts), grid_len, _ptr(op), _ptr(oc), _ptr(on))
    return op[:n].copy(), None if oc is None else oc[:n].copy(), None if on is None else on[:n].copy()


def estimate_normals(points, radius=0.1, knn=30):
    """PointCloud::EstimateNormals(radius, knn)"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.zeros_like(pts)
    lib().orc_estimate_normals(_ptr(pts), len(pts), radius, knn, _ptr(out))
    return out


KD_KNN, KD_RADIUS, KD_KNN_RADIUS = 0, 1, 2


def kdtree_search(points, queries, mode, k, radius=0.0):
    """geometry::KDTree<3>::{KnnSearch, RadiusSearch, KnnRadiusSearch} -> (index [nq,k] -1 padded, dist, count)"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    qs = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
    idx = np.zeros((len(qs), k), np.int32)
    dist = np.zeros((len(qs), k), np.float32)
    cnt = np.zeros(len(qs), np.int32)
    lib().orc_kdtree_search(_ptr(pts), len(pts), _ptr(qs), len(qs), mode, k, radius, k, _ptr(idx), _ptr(dist), _ptr(cnt))
    return idx, dist, cnt


def kdtree_dump(points):
    """the nanoflann tree over `points`: vind, nodes in pre-order (ints [m,5]: left,right,child1,child2,divfeat; floats
    [m,2]: divlow,divhigh), root box (lo xyz, hi xyz)"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    n = len(pts)
    vind = np.zeros(n, np.int32)
    ni = np.zeros((2 * n + 1, 5), np.int32)
    nf = np.zeros((2 * n + 1, 2), np.float32)
    box = np.zeros(6, np.float32)
    m = lib().orc_kdtree_dump(_ptr(pts), n, _ptr(vind), _ptr(ni), _ptr(nf), _ptr(box))
    return vind, ni[:m].copy(), nf[:m].copy(), box


def fpfh(points, normals, knn=100, radius=0.1):
    """registration::ComputeFPFHFeature -> [n,33]"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    out = np.zeros((len(pts), 33), np.float32)
    rc = lib().orc_fpfh(_ptr(pts), _ptr(nrm), len(pts), knn, radius, _ptr(out))
    if rc != 0:
        raise RuntimeError("orc_fpfh: std::sort heap fallback reached (not restated)")
    return out


def feature_matching(src_feat, tgt_feat):
    """registration::FeatureMatching3D -> [m,2] (source, target) indices"""
    sf = np.ascontiguousarray(src_feat, np.float32).reshape(-1, 33)
    tf = np.ascontiguousarray(tgt_feat, np.float32).reshape(-1, 33)
    pairs = np.zeros((len(sf), 2), np.int32)
    m = lib().orc_feature_matching(_ptr(sf), len(sf), _ptr(tf), len(tf), _ptr(pairs))
    return pairs[:m].copy()


def reject_matches(src_pts, tgt_pts, pairs, rounds=3, candidate_num=4, difference=0.1):
    """registration::RejectMatchesRanSaPC x rounds on one default-seeded engine (GlobalRegistration.cpp:168-172) -> kept pairs"""
    s = np.ascontiguousarray(src_pts, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tgt_pts, np.float32).reshape(-1, 3)
    p = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2).copy()
    m = lib().orc_reject_matches(_ptr(s), _ptr(t), _ptr(p), len(p), rounds, candidate_num, difference)
    return p[:m].copy()


def kabsch_f32(a, b):
    """geometry::EstimateRigidTransformation, float build, bit for bit -> T 4x4 float32"""
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
    T = np.zeros(16, np.float32)
    lib().orc_kabsch_f32(_ptr(a), _ptr(b), len(a), _ptr(T))
    return T.reshape(4, 4)


def ransac_hypothesis(a, b, sample8, threshold):
    """one GRANSAC iteration with the given eight pairs -> (T 4x4 float32, inlier flags uint8)"""
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
    s8 = np.ascontiguousarray(sample8, np.int32)
    T = np.zeros(16, np.float32)
    flags = np.zeros(len(a), np.uint8)
    lib().orc_ransac_hypothesis(_ptr(a), _ptr(b), len(a), _ptr(s8), threshold, _ptr(T), _ptr(flags))
    return T.reshape(4, 4), flags


def ransac_select(a, b, samples, threshold):
    """geometry::EstimateRigidTransformationRANSAC with forced samples [iterations, 8] -> (winner, T 4x4 float32, inlier ids)"""
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
    s = np.ascontiguousarray(samples, np.int32).reshape(-1, 8)
    T = np.zeros(16, np.float32)
    flags = np.zeros(len(a), np.uint8)
    w = lib().orc_ransac_select(_ptr(a), _ptr(b), len(a), _ptr(s), len(s), threshold, _ptr(T), _ptr(flags), None)
    return int(w), T.reshape(4, 4), np.nonzero(flags)[0].astype(np.int32)
