"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/_ref/libopref_{f32,f64}.so -- the reference's own
translation units compiled unmodified (oracle/Makefile `make ref`).  Imported by tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() only; never by the product package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_f = C.c_float
c_d = C.c_double
c_i = C.c_int
c_l = C.c_long
c_p = C.c_void_p


def available(kind: str = "f32") -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libopref_{kind}.so"))


def lib(kind: str = "f32"):
    if kind in _LIBS:
        return _LIBS[kind]
    L = C.CDLL(os.path.join(_HERE, "_ref", f"libopref_{kind}.so"))
    L.ref_volume_create.restype = c_p
    L.ref_volume_create.argtypes = [c_f] * 4 + [c_i, c_i] + [c_f] * 5
    L.ref_volume_destroy.argtypes = [c_p]
    L.ref_volume_clear.argtypes = [c_p]
    L.ref_volume_integrate.restype = c_d
    L.ref_volume_integrate.argtypes = [c_p, c_p, c_i, c_p, c_p]
    L.ref_volume_bounding.argtypes = [c_p, c_p, c_i, c_p, c_p, c_p]
    L.ref_volume_prepare_cubes.restype = c_l
    L.ref_volume_prepare_cubes.argtypes = [c_p, c_p, c_i, c_p, c_p, c_l]
    L.ref_volume_num_cubes.restype = c_l
    L.ref_volume_num_cubes.argtypes = [c_p]
    L.ref_volume_download.argtypes = [c_p, c_p, c_p]
    L.ref_volume_upload.argtypes = [c_p, c_p, c_p, c_l]
    L.ref_volume_extract_mesh.restype = c_d
    L.ref_volume_extract_mesh.argtypes = [c_p, c_p, c_p]
    L.ref_volume_mesh_copy.argtypes = [c_p, c_p, c_p, c_p]
    L.ref_volume_write.restype = C.c_bool
    L.ref_volume_write.argtypes = [c_p, C.c_char_p]
    L.ref_volume_read.restype = C.c_bool
    L.ref_volume_read.argtypes = [c_p, C.c_char_p]
    L.ref_volume_transform.restype = c_p
    L.ref_volume_transform.argtypes = [c_p, c_p, c_i]
    L.ref_volume_merge.argtypes = [c_p, c_p]
    L.ref_volume_merge_transformed.argtypes = [c_p, c_p, c_p]
    L.ref_volume_resolution.restype = c_f
    L.ref_volume_resolution.argtypes = [c_p]
    L.ref_clustering_simplify.restype = c_d
    L.ref_clustering_simplify.argtypes = [c_p, c_p, c_p, c_p, c_p, c_f, c_i, c_p]
    L.ref_compute_normals.argtypes = [c_p, c_l, c_p, c_l, c_p]
    L.ref_write_ply.restype = C.c_bool
    L.ref_write_ply.argtypes = [C.c_char_p, c_p, c_p, c_p, c_l, c_p, c_l]
    L.ref_marching_cube_cell.restype = c_i
    L.ref_marching_cube_cell.argtypes = [c_p] * 5
    L.ref_frustum.argtypes = [c_f] * 4 + [c_i, c_i, c_p, c_f, c_f, c_p, c_p, c_l, c_p]
    L.ref_pose_inverse.argtypes = [c_p, c_p]
    L.ref_get_sdf.argtypes = [c_p, c_p, c_i, c_p, c_p, c_l, c_p]
    L.ref_icp.restype = c_d
    L.ref_icp.argtypes = [c_p, c_l, c_p, c_p, c_l, c_p, c_i, c_d, c_p, c_p, c_l, c_p, c_p]
    L.ref_icp_state_create.restype = c_p
    L.ref_icp_state_create.argtypes = [c_p, c_l, c_p, c_p, c_l]
    L.ref_icp_state_destroy.argtypes = [c_p]
    L.ref_icp_iteration.restype = c_l
    L.ref_icp_iteration.argtypes = [c_p, c_p, c_d, c_p, c_p, c_p, c_p, c_p, c_p]
    L.ref_se3_exp.argtypes = [c_p, c_p]
    L.ref_kabsch.argtypes = [c_p, c_p, c_l, c_p]
    L.ref_correspondences.restype = c_l
    L.ref_correspondences.argtypes = [c_p, c_p, c_i, c_i] + [c_f] * 4 + [c_p, c_p, c_l]
    L.ref_normalize_intensity.argtypes = [c_p, c_p, c_i, c_i, c_p, c_l]
    L.ref_multiscale.restype = c_l
    L.ref_multiscale.argtypes = [c_p, c_p, c_i, c_i] + [c_f] * 4 + [c_p, c_i, c_p, c_p, c_p, c_p, c_l, c_p, c_p, c_p]
    L.ref_single_iteration.restype = c_l
    L.ref_single_iteration.argtypes = [c_p, c_p, c_i, c_i, c_i] + [c_f] * 4 + [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_l]
    if kind == "f32":
        L.ref_load_from_depth.restype = c_l
        L.ref_load_from_depth.argtypes = [c_p, c_i, c_i, c_i] + [c_f] * 5 + [c_p]
        L.ref_downsample.restype = c_l
        L.ref_downsample.argtypes = [c_p, c_p, c_p, c_l, c_f, c_p, c_p, c_p]
        L.ref_kdtree_search.argtypes = [c_p, c_l, c_p, c_l, c_i, c_i, c_f, c_l, c_p, c_p, c_p]
        L.ref_fpfh.restype = c_d
        L.ref_fpfh.argtypes = [c_p, c_p, c_l, c_i, c_f, c_p]
        L.ref_feature_matching.restype = c_l
        L.ref_feature_matching.argtypes = [c_p, c_l, c_p, c_l, c_p]
        L.ref_reject_matches.restype = c_l
        L.ref_reject_matches.argtypes = [c_p, c_l, c_p, c_l, c_p, c_l, c_i, c_i, c_f]
        L.ref_ransac_hypothesis.restype = c_d
        L.ref_ransac_hypothesis.argtypes = [c_p, c_p, c_l, c_p, c_d, c_p]
        L.ref_ransac_registration.restype = c_d
        L.ref_ransac_registration.argtypes = [c_p, c_l, c_p, c_l, c_p, c_p, c_i, c_d, c_p, c_p, c_p]
        L.ref_estimate_normals.restype = c_d
        L.ref_estimate_normals.argtypes = [c_p, c_l, c_f, c_i, c_p]
    L.ref_set_quiet(1)
    _LIBS[kind] = L
    return L


def _ptr(a):
    return a.ctypes.data_as(c_p) if a is not None else None


def _pose_cm(pose):
    """4x4 (row-major numpy) -> column-major float32[16] as the C entry points take it."""
    return np.ascontiguousarray(np.asarray(pose, dtype=np.float32).T).reshape(16)


def _from_cm(a):
    return np.asarray(a, dtype=np.float64).reshape(4, 4).T.copy()


class RefVolume:
    """one_piece::integration::CubeHandler of the compiled reference."""

    def __init__(self, cam, voxel_resolution=0.01, truncation=0.1, near=0.5, far=5.0, kind="f32"):
        self.L = lib(kind)
        self.cam = cam
        self.h = self.L.ref_volume_create(cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.depth_scale,
                                          voxel_resolution, truncation, near, far)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_volume_destroy(self.h)
            self.h = None

    def clear(self):
        self.L.ref_volume_clear(self.h)

    def integrate(self, depth, bgr, pose) -> float:
        depth = np.ascontiguousarray(depth)
        bgr = np.ascontiguousarray(bgr)
        p = _pose_cm(pose)
        return self.L.ref_volume_integrate(self.h, _ptr(depth), int(depth.dtype == np.uint16), _ptr(bgr), _ptr(p))

    def bounding(self, depth, pose):
        depth = np.ascontiguousarray(depth)
        p = _pose_cm(pose)
        mx = np.zeros(3)
        mn = np.zeros(3)
        self.L.ref_volume_bounding(self.h, _ptr(depth), int(depth.dtype == np.uint16), _ptr(p), _ptr(mx), _ptr(mn))
        return mx, mn

    def prepare_cubes(self, depth, pose):
        depth = np.ascontiguousarray(depth)
        p = _pose_cm(pose)
        cap = 1 << 22
        ids = np.zeros((cap, 3), np.int32)
        n = self.L.ref_volume_prepare_cubes(self.h, _ptr(depth), int(depth.dtype == np.uint16), _ptr(p), _ptr(ids), cap)
        return ids[:n].copy()

    def num_cubes(self) -> int:
        return self.L.ref_volume_num_cubes(self.h)

    def download(self):
        """-> (ids [n,3] int32, voxels [n,512,5] float32) sorted lexicographically by cube id."""
        n = self.num_cubes()
        ids = np.zeros((n, 3), np.int32)
        vox = np.zeros((n, 512, 5), np.float32)
        self.L.ref_volume_download(self.h, _ptr(ids), _ptr(vox))
        order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))
        return ids[order], vox[order]

    def upload(self, ids, vox):
        ids = np.ascontiguousarray(ids, np.int32)
        vox = np.ascontiguousarray(vox, np.float32)
        self.L.ref_volume_upload(self.h, _ptr(ids), _ptr(vox), len(ids))

    @classmethod
    def _wrap(cls, L, cam, handle):
        o = cls.__new__(cls)
        o.L, o.cam, o.h = L, cam, handle
        return o

    def transform(self, trans, nearest: bool):
        """CubeHandler::Transform / TransformNearest -> new RefVolume"""
        p = _pose_cm(trans)
        return RefVolume._wrap(self.L, self.cam, self.L.ref_volume_transform(self.h, _ptr(p), int(nearest)))

    def merge(self, other, trans=None):
        if trans is None:
            self.L.ref_volume_merge(self.h, other.h)
        else:
            p = _pose_cm(trans)
            self.L.ref_volume_merge_transformed(self.h, other.h, _ptr(p))

    def resolution(self) -> float:
        return self.L.ref_volume_resolution(self.h)

    def extract_mesh(self):
        """-> (seconds, points [nv,3], colors [nv,3], triangles [nt,3])"""
        nv, nt = c_l(0), c_l(0)
        dt = self.L.ref_volume_extract_mesh(self.h, C.byref(nv), C.byref(nt))
        pts = np.zeros((nv.value, 3), np.float32)
        col = np.zeros((nv.value, 3), np.float32)
        tri = np.zeros((nt.value, 3), np.uint32)
        self.L.ref_volume_mesh_copy(self.h, _ptr(pts), _ptr(col), _ptr(tri))
        return dt, pts, col, tri

    def get_sdf(self, depth, pose, points):
        depth = np.ascontiguousarray(depth)
        points = np.ascontiguousarray(points, np.float32)
        out = np.zeros(len(points), np.float32)
        p = _pose_cm(pose)
        self.L.ref_get_sdf(self.h, _ptr(depth), int(depth.dtype == np.uint16), _ptr(p), _ptr(points), len(points), _ptr(out))
        return out

    def write(self, path):
        return self.L.ref_volume_write(self.h, str(path).encode())

    def read(self, path):
        return self.L.ref_volume_read(self.h, str(path).encode())


def marching_cube_cell(corners, sdf, colors, kind="f32"):
    corners = np.ascontiguousarray(corners, np.float32)
    sdf = np.ascontiguousarray(sdf, np.float32)
    colors = np.ascontiguousarray(colors, np.float32)
    xyz = np.zeros((15, 3), np.float32)
    rgb = np.zeros((15, 3), np.float32)
    n = lib(kind).ref_marching_cube_cell(_ptr(corners), _ptr(sdf), _ptr(colors), _ptr(xyz), _ptr(rgb))
    return xyz[:n], rgb[:n]


def frustum(cam, pose, far, near, points, kind="f32"):
    points = np.ascontiguousarray(points, np.float32)
    planes = np.zeros((6, 4))
    mask = np.zeros(len(points), np.uint8)
    p = _pose_cm(pose)
    lib(kind).ref_frustum(cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, _ptr(p), far, near, _ptr(planes),
                          _ptr(points), len(points), _ptr(mask))
    return planes, mask.astype(bool)


def pose_inverse(pose, kind="f32"):
    p = _pose_cm(pose)
    out = np.zeros(16)
    lib(kind).ref_pose_inverse(_ptr(p), _ptr(out))
    return _from_cm(out)


def icp(src, tgt, tgt_normals, init_T, max_iter=30, threshold=0.2, kind="f32"):
    """registration::PointToPlane (normals given) / PointToPoint (None).
    -> dict(T [4,4] f64, pairs [n,2] int32, rmse, seconds)"""
    src = np.ascontiguousarray(src, np.float32)
    tgt = np.ascontiguousarray(tgt, np.float32)
    nrm = np.ascontiguousarray(tgt_normals, np.float32) if tgt_normals is not None else None
    T0 = _pose_cm(init_T)
    out_T = np.zeros(16)
    cap = len(src)
    pairs = np.zeros((cap, 2), np.int32)
    n = c_l(0)
    rmse = c_d(0)
    dt = lib(kind).ref_icp(_ptr(src), len(src), _ptr(tgt), _ptr(nrm), len(tgt), _ptr(T0), max_iter, threshold,
                           _ptr(out_T), _ptr(pairs), cap, C.byref(n), C.byref(rmse))
    return dict(T=_from_cm(out_T), pairs=pairs[: n.value].copy(), rmse=rmse.value, seconds=dt)


class RefIcpState:
    """Teacher-forced single iterations of the reference's ICP loop body (ICP.cpp:175-205)."""

    def __init__(self, src, tgt, tgt_normals, kind="f32"):
        self.L = lib(kind)
        self.src = np.ascontiguousarray(src, np.float32)
        self.tgt = np.ascontiguousarray(tgt, np.float32)
        self.nrm = np.ascontiguousarray(tgt_normals, np.float32) if tgt_normals is not None else None
        self.h = self.L.ref_icp_state_create(_ptr(self.src), len(self.src), _ptr(self.tgt), _ptr(self.nrm), len(self.tgt))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_icp_state_destroy(self.h)
            self.h = None

    def iteration(self, T_in, threshold):
        Tin = np.ascontiguousarray(np.asarray(T_in, np.float64).T).reshape(16)
        nn = np.zeros(len(self.src), np.int32)
        JTJ = np.zeros(36)
        JTr = np.zeros(6)
        x = np.zeros(6)
        Tout = np.zeros(16)
        rmse = c_d(0)
        n = self.L.ref_icp_iteration(self.h, _ptr(Tin), threshold, _ptr(nn), _ptr(JTJ), _ptr(JTr), _ptr(x), _ptr(Tout),
                                     C.byref(rmse))
        return dict(n_inliers=n, nn=nn, JTJ=JTJ.reshape(6, 6), JTr=JTr, x=x, T=_from_cm(Tout), rmse=rmse.value)


def se3_exp(x, kind="f32"):
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros(16)
    lib(kind).ref_se3_exp(_ptr(x), _ptr(out))
    return _from_cm(out)


def kabsch(a, b, kind="f32"):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    out = np.zeros(16)
    lib(kind).ref_kabsch(_ptr(a), _ptr(b), len(a), _ptr(out))
    return _from_cm(out)


def load_from_depth(depth, cam):
    depth = np.ascontiguousarray(depth)
    out = np.zeros((depth.size, 3), np.float32)
    n = lib("f32").ref_load_from_depth(_ptr(depth), int(depth.dtype == np.uint16), cam.width, cam.height, cam.fx, cam.fy,
                                       cam.cx, cam.cy, cam.depth_scale, _ptr(out))
    return out[:n].copy()


def estimate_normals(xyz, radius=0.1, knn=30):
    xyz = np.ascontiguousarray(xyz, np.float32)
    out = np.zeros_like(xyz)
    dt = lib("f32").ref_estimate_normals(_ptr(xyz), len(xyz), radius, knn, _ptr(out))
    return out, dt


def correspondences(src_depth, tgt_depth, cam, T, kind="f32"):
    """odometry::ComputeCorrespondencePixelWise on NaN-masked float32 depth maps -> [n,4] (v_s,u_s,v_t,u_t)"""
    sd = np.ascontiguousarray(src_depth, np.float32)
    td = np.ascontiguousarray(tgt_depth, np.float32)
    h, w = sd.shape
    Tcm = np.ascontiguousarray(np.asarray(T, np.float64).T).reshape(16)
    pairs = np.zeros((w * h, 4), np.uint32)
    n = lib(kind).ref_correspondences(_ptr(sd), _ptr(td), w, h, cam.fx, cam.fy, cam.cx, cam.cy, _ptr(Tcm), _ptr(pairs), w * h)
    return pairs[:n].copy()


def normalize_intensity(src_gray, tgt_gray, pairs, kind="f32"):
    s = np.ascontiguousarray(src_gray, np.float32).copy()
    t = np.ascontiguousarray(tgt_gray, np.float32).copy()
    p = np.ascontiguousarray(pairs, np.uint32)
    lib(kind).ref_normalize_intensity(_ptr(s), _ptr(t), s.shape[1], s.shape[0], _ptr(p), len(p))
    return s, t


def multiscale(src_imgs, tgt_imgs, cam, init_T, term=0, kind="f32"):
    """Odometry::MultiScaleComputing over caller-supplied pyramids.  *_imgs[what][level] float32 arrays with
    what = 0 gray, 1 depth, 2 gray dx, 3 gray dy, 4 depth dx, 5 depth dy."""
    keep = []

    def table(imgs):
        arr = (c_p * 18)()
        for a in range(6):
            for l in range(3):
                m = np.ascontiguousarray(imgs[a][l], np.float32)
                keep.append(m)
                arr[a * 3 + l] = m.ctypes.data
        return arr
    h, w = src_imgs[0][0].shape
    sa, ta = table(src_imgs), table(tgt_imgs)
    T0 = np.ascontiguousarray(np.asarray(init_T, np.float64).T).reshape(16)
    Tout = np.zeros(16)
    rmse = c_d(0)
    ok = c_i(0)
    nit = c_i(0)
    pairs = np.zeros((w * h, 4), np.uint32)
    cpi = np.zeros(64, np.int64)
    tpi = np.zeros((64, 16))
    n = lib(kind).ref_multiscale(sa, ta, w, h, cam.fx, cam.fy, cam.cx, cam.cy, _ptr(T0), term, _ptr(Tout), C.byref(rmse), C.byref(ok),
                                 _ptr(pairs), w * h, _ptr(cpi), _ptr(tpi), C.byref(nit))
    k = nit.value
    return dict(T=_from_cm(Tout), rmse=rmse.value, success=bool(ok.value), pairs=pairs[:n].copy(), corr_per_iteration=cpi[:k].copy(),
                T_per_iteration=np.stack([_from_cm(t) for t in tpi[:k]]) if k else np.zeros((0, 4, 4)))


def single_iteration(src_imgs, tgt_imgs, cam, level, T, term=0, kind="f32"):
    """One teacher-forced iteration of the reference at a pyramid level -> dict(T, JTJ, JTr, r2, pairs)"""
    keep = []

    def table(imgs):
        arr = (c_p * 18)()
        for a in range(6):
            for l in range(3):
                m = np.ascontiguousarray(imgs[a][l], np.float32)
                keep.append(m)
                arr[a * 3 + l] = m.ctypes.data
        return arr
    h, w = src_imgs[0][0].shape
    sa, ta = table(src_imgs), table(tgt_imgs)
    T0 = np.ascontiguousarray(np.asarray(T, np.float64).T).reshape(16)
    Tout, JTJ, JTr, r2 = np.zeros(16), np.zeros(36), np.zeros(6), c_d(0)
    pairs = np.zeros(((w >> level) * (h >> level), 4), np.uint32)
    n = lib(kind).ref_single_iteration(sa, ta, level, w, h, cam.fx, cam.fy, cam.cx, cam.cy, _ptr(T0), term, _ptr(Tout), _ptr(JTJ),
                                       _ptr(JTr), C.byref(r2), _ptr(pairs), len(pairs))
    return dict(T=_from_cm(Tout), JTJ=JTJ.reshape(6, 6), JTr=JTr, r2=r2.value, pairs=pairs[:n].copy())


def clustering_simplify(points, colors, triangles, grid_len, with_normals=False, kind="f32"):
    """TriangleMesh::ClusteringSimplify of the compiled reference -> (points, colors, triangles, normals or None, seconds)"""
    pts = np.array(points, np.float32, copy=True).reshape(-1, 3)
    col = None if colors is None else np.array(colors, np.float32, copy=True).reshape(-1, 3)
    tri = np.array(triangles, np.uint32, copy=True).reshape(-1, 3)
    nrm = np.zeros_like(pts) if with_normals else None
    nv, nt = c_l(len(pts)), c_l(len(tri))
    dt = lib(kind).ref_clustering_simplify(_ptr(pts), _ptr(col), C.byref(nv), _ptr(tri), C.byref(nt), grid_len, int(with_normals), _ptr(nrm))
    return (pts[: nv.value].copy(), None if col is None else col[: nv.value].copy(), tri[: nt.value].copy(),
            None if nrm is None else nrm[: nv.value].copy(), dt)


def compute_normals(points, triangles, kind="f32"):
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    tri = np.ascontiguousarray(triangles, np.uint32).reshape(-1, 3)
    out = np.zeros_like(pts)
    lib(kind).ref_compute_normals(_ptr(pts), len(pts), _ptr(tri), len(tri), _ptr(out))
    return out


def write_ply(path, points, normals, colors, triangles, kind="f32"):
    """TriangleMesh::WriteToPLY of the compiled reference"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    nrm = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    col = None if colors is None else np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
    tri = np.ascontiguousarray(triangles, np.uint32).reshape(-1, 3)
    return lib(kind).ref_write_ply(str(path).encode(), _ptr(pts), _ptr(nrm), _ptr(col), len(pts), _ptr(tri), len(tri))


def downsample(points, colors, normals, grid_len):
    """PointCloud::DownSample of the compiled reference (float32 build)"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    col = None if colors is None else np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
    nrm = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    op = np.zeros_like(pts)
    oc = None if col is None else np.zeros_like(pts)
    on = None if nrm is None else np.zeros_like(pts)
    n = lib("f32").ref_downsample(_ptr(pts), _ptr(col), _ptr(nrm), len(pts), grid_len, _ptr(op), _ptr(oc), _ptr(on))
    return op[:n].copy(), None if oc is None else oc[:n].copy(), None if on is None else on[:n].copy()


def kdtree_search(points, queries, mode, k, radius=0.0):
    """geometry::KDTree<3> searches of the compiled reference (mode 0 Knn, 1 Radius, 2 KnnRadius)"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    qs = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
    idx = np.zeros((len(qs), k), np.int32)
    dist = np.zeros((len(qs), k), np.float32)
    cnt = np.zeros(len(qs), np.int32)
    lib("f32").ref_kdtree_search(_ptr(pts), len(pts), _ptr(qs), len(qs), mode, k, radius, k, _ptr(idx), _ptr(dist), _ptr(cnt))
    return idx, dist, cnt


def fpfh(points, normals, knn=100, radius=0.1):
    """registration::ComputeFPFHFeature of the compiled reference -> ([n,33], seconds)"""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    out = np.zeros((len(pts), 33), np.float32)
    dt = lib("f32").ref_fpfh(_ptr(pts), _ptr(nrm), len(pts), knn, radius, _ptr(out))
    return out, dt


def feature_matching(src_feat, tgt_feat):
    sf = np.ascontiguousarray(src_feat, np.float32).reshape(-1, 33)
    tf = np.ascontiguousarray(tgt_feat, np.float32).reshape(-1, 33)
    pairs = np.zeros((len(sf), 2), np.int32)
    m = lib("f32").ref_feature_matching(_ptr(sf), len(sf), _ptr(tf), len(tf), _ptr(pairs))
    return pairs[:m].copy()


def reject_matches(src_pts, tgt_pts, pairs, rounds=3, candidate_num=4, difference=0.1):
    s = np.ascontiguousarray(src_pts, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tgt_pts, np.float32).reshape(-1, 3)
    p = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2).copy()
    m = lib("f32").ref_reject_matches(_ptr(s), len(s), _ptr(t), len(t), _ptr(p), len(p), rounds, candidate_num, difference)
    return p[:m].copy()


def ransac_hypothesis(a, b, sample8, threshold):
    """TransformationModel over the eight pairs + Evaluate over all -> (inlier fraction, flags)"""
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
    s8 = np.ascontiguousarray(sample8, np.int32)
    flags = np.zeros(len(a), np.uint8)
    frac = lib("f32").ref_ransac_hypothesis(_ptr(a), _ptr(b), len(a), _ptr(s8), threshold, _ptr(flags))
    return frac, flags


def ransac_registration(src_pts, tgt_pts, src_feat, tgt_feat, max_iteration, threshold):
    """registration::RansacRegistration on precomputed features (randomly seeded inside GRANSAC) -> (T, n_inliers, rmse, seconds)"""
    s = np.ascontiguousarray(src_pts, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tgt_pts, np.float32).reshape(-1, 3)
    sf = np.ascontiguousarray(src_feat, np.float32).reshape(-1, 33)
    tf = np.ascontiguousarray(tgt_feat, np.float32).reshape(-1, 33)
    T = np.zeros(16, np.float64)
    n_in, rmse = C.c_long(0), C.c_double(0)
    dt = lib("f32").ref_ransac_registration(_ptr(s), len(s), _ptr(t), len(t), _ptr(sf), _ptr(tf), max_iteration, threshold, _ptr(T),
                                            C.byref(n_in), C.byref(rmse))
    return T.reshape(4, 4).T.copy(), n_in.value, rmse.value, dt
