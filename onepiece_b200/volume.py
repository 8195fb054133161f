"""Python mirror of one_piece::integration::CubeHandler (reference src/Integration/CubeHandler.h:24-366) over the
C-ABI.  Method names and argument meaning follow the reference class so that tests read like its examples
(example/ImageSequenceIntegration.cpp:20-53).  All computation happens in libonepiece_b200.so on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .scenes import Camera


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def pose_colmajor(pose) -> np.ndarray:
    """4x4 camera-to-world (numpy, row-major indexing) -> Eigen column-major float32[16]."""
    return np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(4, 4).T).reshape(16)


def depth_type_of(depth) -> int:
    if depth.dtype == np.float32:
        return capi.OPB_DEPTH_F32
    if depth.dtype == np.uint16:
        return capi.OPB_DEPTH_U16
    return -1  # rejected by the library like the reference's "Unknown depth image type"


class CubeHandler:
    def __init__(self, camera: Camera = Camera(), voxel_resolution: float = 0.01, truncation: float = 0.1,
                 near: float = 0.5, far: float = 5.0, max_cubes: int = 1 << 17, device: int = 0,
                 storage: int = capi.OPB_STORAGE_F32, shard=None, stream=None):
        d = capi.VolumeDesc()
        capi.lib.opb_volume_desc_default(C.byref(d))
        d.fx, d.fy, d.cx, d.cy = camera.fx, camera.fy, camera.cx, camera.cy
        d.width, d.height, d.depth_scale = camera.width, camera.height, camera.depth_scale
        d.voxel_resolution, d.truncation, d.near_plane, d.far_plane = voxel_resolution, truncation, near, far
        d.max_cubes, d.device, d.storage = max_cubes, device, storage
        if shard is not None:
            d.shard_rank, d.shard_world, d.shard_axis, d.shard_slab_cubes = shard
        d.stream = stream
        self.desc = d
        self.camera = camera
        self._h = C.c_void_p()
        capi.check(capi.lib.opb_volume_create(C.byref(d), C.byref(self._h)))

    # -- lifetime ---------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            capi.lib.opb_volume_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setters (CubeHandler.h:36,137,141,349-356) ---------------------------------------------------
    def _push_params(self):
        capi.check(capi.lib.opb_volume_set_params(self._h, C.byref(self.desc)))

    def SetVoxelResolution(self, r):
        self.desc.voxel_resolution = r
        self._push_params()

    def SetTruncation(self, t):
        self.desc.truncation = t
        self._push_params()

    def SetFarPlane(self, f):
        self.desc.far_plane = f
        self._push_params()

    def SetNearPlane(self, n):
        self.desc.near_plane = n
        self._push_params()

    def SetCamera(self, cam: Camera):
        d = self.desc
        d.fx, d.fy, d.cx, d.cy = cam.fx, cam.fy, cam.cx, cam.cy
        d.width, d.height, d.depth_scale = cam.width, cam.height, cam.depth_scale
        self.camera = cam
        self._push_params()

    def Clear(self):
        capi.check(capi.lib.opb_volume_clear(self._h))

    # -- integration ------------------------------------------------------------------------------------
    def IntegrateImage(self, depth, rgb, pose):
        """CubeHandler::IntegrateImage(depth, rgb, pose) with host arrays (synchronous)."""
        depth = np.ascontiguousarray(depth)
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        p = pose_colmajor(pose)
        capi.check(capi.lib.opb_volume_integrate(self._h, _ptr(depth), depth_type_of(depth), _ptr(rgb), _ptr(p)))

    def IntegrateCloud(self, cloud, pose):
        """CubeHandler::IntegrateImage with the depth + colour images a registration.DeviceCloud was loaded from (already in
        HBM); synchronous."""
        p = pose_colmajor(pose)
        capi.check(capi.lib.opb_volume_integrate_cloud(self._h, cloud._h, _ptr(p)))

    def IntegrateImageAsync(self, depth_ptr, depth_type, bgr_ptr, pose_cm):
        capi.check(capi.lib.opb_volume_integrate_async(self._h, _ptr(depth_ptr), depth_type, _ptr(bgr_ptr), _ptr(pose_cm)))

    def IntegrateImageDevice(self, d_depth_ptr, depth_type, d_bgr_ptr, pose_cm):
        capi.check(capi.lib.opb_volume_integrate_device(self._h, _ptr(d_depth_ptr), depth_type, _ptr(d_bgr_ptr), _ptr(pose_cm)))

    def Synchronize(self):
        capi.check(capi.lib.opb_volume_synchronize(self._h))

    def PrepareCubes(self, depth, pose):
        """CubeHandler::PrepareCubes(depth, pose, cube_id_list) (CubeHandler.cpp:147-196) -> [n, 3] int32 cube ids"""
        depth = np.ascontiguousarray(depth)
        p = pose_colmajor(pose)
        n = C.c_size_t(0)
        capi.check(capi.lib.opb_volume_prepare_cubes(self._h, _ptr(depth), depth_type_of(depth), _ptr(p), None, C.byref(n)))
        ids = np.zeros((max(n.value, 1), 3), np.int32)
        n = C.c_size_t(len(ids))
        capi.check(capi.lib.opb_volume_last_frame_cubes(self._h, _ptr(ids), C.byref(n)))
        return ids[: n.value].copy()

    def ComputeBounding(self, depth, pose):
        """CubeHandler::ComputeBounding(depth, pose, max_pos, min_pos) (CubeHandler.cpp:116-145) -> (max_pos, min_pos); read-only"""
        depth = np.ascontiguousarray(depth)
        p = pose_colmajor(pose)
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        capi.check(capi.lib.opb_volume_compute_bounding(self._h, _ptr(depth), depth_type_of(depth), _ptr(p), _ptr(mn), _ptr(mx)))
        return mx, mn

    def FrameStats(self) -> capi.FrameStats:
        s = capi.FrameStats()
        capi.check(capi.lib.opb_volume_frame_stats(self._h, C.byref(s)))
        return s

    def SetProfiling(self, on: bool):
        capi.check(capi.lib.opb_volume_set_profiling(self._h, int(on)))

    def ProfileRead(self, reset=True):
        a, b, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        capi.check(capi.lib.opb_volume_profile_read(self._h, C.byref(a), C.byref(b), C.byref(n), int(reset)))
        return a.value, b.value, n.value

    # -- content ----------------------------------------------------------------------------------------
    def NumCubes(self) -> int:
        n = C.c_size_t(0)
        capi.check(capi.lib.opb_volume_num_cubes(self._h, C.byref(n)))
        return n.value

    def GetCubeMap(self):
        """-> (ids [n,3] int32, voxels [n,512,5] float32 = sdf, weight, c0, c1, c2) sorted by cube id."""
        n = self.NumCubes()
        ids = np.zeros((n, 3), np.int32)
        vox = np.zeros((n, 512, 5), np.float32)
        cnt = C.c_size_t(n)
        capi.check(capi.lib.opb_volume_download(self._h, _ptr(ids), _ptr(vox), C.byref(cnt)))
        order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))
        return ids[order], vox[order]

    def SetCubeMap(self, ids, vox):
        ids = np.ascontiguousarray(ids, np.int32)
        vox = np.ascontiguousarray(vox, np.float32)
        capi.check(capi.lib.opb_volume_upload(self._h, _ptr(ids), _ptr(vox), len(ids)))

    # -- resampling and merging (CubeHandler.h:145-177,242-338) ----------------------------------------------
    @classmethod
    def _adopt(cls, handle, camera):
        o = cls.__new__(cls)
        o._h = handle
        o.camera = camera
        o.desc = capi.VolumeDesc()
        capi.check(capi.lib.opb_volume_get_desc(handle, C.byref(o.desc)))
        return o

    def _transform(self, trans, nearest, result_resolution, max_cubes):
        out = C.c_void_p()
        t = pose_colmajor(trans)
        capi.check(capi.lib.opb_volume_transform(self._h, _ptr(t), int(nearest), result_resolution, max_cubes, C.byref(out)))
        return CubeHandler._adopt(out, self.camera)

    def Transform(self, trans, max_cubes: int = 0):
        """CubeHandler::Transform(trans) -> new CubeHandler (trilinear resampling; c_para is copied)."""
        return self._transform(trans, False, 0.0, max_cubes)

    def TransformNearest(self, trans, max_cubes: int = 0, result_resolution: float = 0.01):
        """CubeHandler::TransformNearest(trans) -> new CubeHandler.  The reference forgets to copy c_para, so its result runs
        at CubePara's default VoxelResolution 0.01: that is the default here too (pass 0 for the source resolution)."""
        return self._transform(trans, True, result_resolution, max_cubes)

    def Merge(self, another, trans=None) -> bool:
        """CubeHandler::Merge(another[, trans]); False (and a warning, like the reference) when the resolutions differ."""
        if trans is not None:
            if self.desc.voxel_resolution != another.desc.voxel_resolution:   # checked before transforming (CubeHandler.h:170-174)
                print("[Warning]::[MergeVoxelHash]::Voxel resolution is not identical.")
                return False
            another = another.Transform(trans)
        rc = capi.lib.opb_volume_merge(self._h, another._h)
        if rc == capi.OPB_ERR_INVALID and b"MergeVoxelHash" in capi.lib.opb_last_error():
            print(capi.lib.opb_last_error().decode())
            return False
        capi.check(rc)
        return True

    # -- boundary-cube exchange for multi-GPU Marching Cubes (include/onepiece_b200.h, opb_volume_halo_*) -----------
    def HaloCount(self) -> int:
        n = C.c_size_t(0)
        capi.check(capi.lib.opb_volume_halo_export(self._h, None, None, 0, C.byref(n)))
        return n.value

    def HaloExport(self, ids, layers, cap: int) -> int:
        """ids: int32[cap,3], layers: float32[cap,5,64]; numpy arrays or raw host/device addresses.  -> cubes written"""
        n = C.c_size_t(0)
        capi.check(capi.lib.opb_volume_halo_export(self._h, _ptr(ids), _ptr(layers), cap, C.byref(n)))
        return n.value

    def HaloImport(self, ids, layers, n: int):
        capi.check(capi.lib.opb_volume_halo_import(self._h, _ptr(ids), _ptr(layers), n))

    def HaloClear(self):
        capi.check(capi.lib.opb_volume_halo_clear(self._h))

    # ... with the transport inside the kernels (peer memory over NVLink): see fusion.attach_halo_peers
    def HaloPeerBuffer(self, cap_cubes: int):
        """-> (device address of this volume's receive box, its 64-byte cudaIpc handle)"""
        buf = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        capi.check(capi.lib.opb_volume_halo_peer_buffer(self._h, cap_cubes, C.byref(buf), handle))
        return buf.value, bytes(handle)

    def HaloPeerAttach(self, dst_buffer, dst_cap_cubes: int, src_buffer):
        capi.check(capi.lib.opb_volume_halo_peer_attach(self._h, C.c_void_p(dst_buffer), dst_cap_cubes, C.c_void_p(src_buffer)))
        self._halo_peers = bool(dst_buffer)

    def HaloPeersAttached(self) -> bool:
        return getattr(self, "_halo_peers", False)

    def HaloExchangePeer(self):
        """collective: -> (boundary cubes sent to rank-1, ghost cubes imported from rank+1)"""
        ns, ni = C.c_size_t(0), C.c_size_t(0)
        capi.check(capi.lib.opb_volume_halo_exchange_peer(self._h, C.byref(ns), C.byref(ni)))
        return ns.value, ni.value

    def HaloExchangeBegin(self):
        capi.check(capi.lib.opb_volume_halo_exchange_begin(self._h))

    def HaloExchangeEnd(self):
        ns, ni = C.c_size_t(0), C.c_size_t(0)
        capi.check(capi.lib.opb_volume_halo_exchange_end(self._h, C.byref(ns), C.byref(ni)))
        return ns.value, ni.value

    # -- one frame uploaded in row bands by the ranks of a partitioned volume (opb_volume_frame_ring_*): fusion.attach_frame_ring
    def FrameRingBuffer(self):
        """-> (device address of this volume's frame ring, its 64-byte cudaIpc handle)"""
        buf = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        capi.check(capi.lib.opb_volume_frame_ring_buffer(self._h, C.byref(buf), handle))
        return buf.value, bytes(handle)

    def FrameRingAttach(self, rank: int, world: int, buffers):
        """buffers: the `world` ring addresses as mapped in this process (own ring at [rank]); None detaches"""
        if buffers is None:
            capi.check(capi.lib.opb_volume_frame_ring_attach(self._h, 0, 0, None))
            return
        arr = (C.c_void_p * world)(*[C.c_void_p(b) for b in buffers])
        capi.check(capi.lib.opb_volume_frame_ring_attach(self._h, rank, world, arr))

    def IntegrateRowsAsync(self, depth_rows_ptr, depth_type, bgr_rows_ptr, row0: int, n_rows: int, pose_cm):
        """collective, asynchronous: this rank's band of the frame from (pinned) host memory; see include/onepiece_b200.h"""
        capi.check(capi.lib.opb_volume_integrate_rows_async(self._h, C.c_void_p(depth_rows_ptr), depth_type, C.c_void_p(bgr_rows_ptr), row0,
                                                            n_rows, _ptr(pose_cm)))

    def FrameRingStatus(self):
        capi.check(capi.lib.opb_volume_frame_ring_status(self._h))

    def NumGhostCubes(self) -> int:
        n = C.c_size_t(0)
        capi.check(capi.lib.opb_volume_num_ghost_cubes(self._h, C.byref(n)))
        return n.value

    # -- Marching Cubes ---------------------------------------------------------------------------------
    def ExtractTriangleMesh(self, cube_order=None):
        """CubeHandler::ExtractTriangleMesh -> (points [nv,3] f32, colors [nv,3] f32, triangles [nt,3] u32).  cube_order (int32 [n,3],
        every cube once): emit the cubes in that sequence -- e.g. the iteration order of the reference's cube map -- instead of
        block-pool order (opb_volume_extract_mesh_ordered)."""
        xyz, rgb, tri = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nv, nt = C.c_size_t(0), C.c_size_t(0)
        if cube_order is None:
            capi.check(capi.lib.opb_volume_extract_mesh(self._h, C.byref(xyz), C.byref(rgb), C.byref(tri), C.byref(nv), C.byref(nt)))
        else:
            order = np.ascontiguousarray(cube_order, np.int32).reshape(-1, 3)
            capi.check(capi.lib.opb_volume_extract_mesh_ordered(self._h, _ptr(order), len(order), C.byref(xyz), C.byref(rgb), C.byref(tri),
                                                                C.byref(nv), C.byref(nt)))
        n = nv.value
        if n == 0:
            return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32)
        pts = np.ctypeslib.as_array(C.cast(xyz, C.POINTER(C.c_float)), shape=(n * 3,)).reshape(n, 3).copy()
        col = np.ctypeslib.as_array(C.cast(rgb, C.POINTER(C.c_float)), shape=(n * 3,)).reshape(n, 3).copy()
        t = np.ctypeslib.as_array(C.cast(tri, C.POINTER(C.c_uint32)), shape=(nt.value * 3,)).reshape(nt.value, 3).copy()
        for p in (xyz, rgb, tri):
            capi.lib.opb_free(p)
        return pts, col, t

    def ExtractTriangleMeshClustered(self, grid_len: float):
        """ExtractTriangleMesh followed by TriangleMesh::ClusteringSimplify(grid_len) (example/DenseFusion/DenseFusion.cpp:99-105)
        with the raw mesh staying on the device -> onepiece_b200.mesh.TriangleMesh"""
        from .mesh import TriangleMesh, _take
        xyz, rgb, tri = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nv, nt = C.c_size_t(0), C.c_size_t(0)
        capi.check(capi.lib.opb_volume_extract_mesh_clustered(self._h, grid_len, C.byref(xyz), C.byref(rgb), C.byref(tri), C.byref(nv), C.byref(nt)))
        return TriangleMesh(_take(xyz, C.c_float, nv.value * 3, (nv.value, 3)), _take(rgb, C.c_float, nv.value * 3, (nv.value, 3)),
                            _take(tri, C.c_uint32, nt.value * 3, (nt.value, 3)))

    # -- the reference's .cubes stream (CubeHandler.h:40-69,113-128) ---------------------------------------
    def WriteToFile(self, filename: str) -> bool:
        from . import formats
        ids, vox = self.GetCubeMap()
        formats.write_cubes(filename, ids, vox)
        return True

    def ReadFromFile(self, filename: str) -> bool:
        from . import formats
        got = formats.read_cubes(filename)
        if got is None:
            return False
        self.SetCubeMap(*got)
        return True

    def CountMesh(self):
        nv, nt = C.c_size_t(0), C.c_size_t(0)
        capi.check(capi.lib.opb_volume_count_mesh(self._h, C.byref(nv), C.byref(nt)))
        return nv.value, nt.value
