"""Sub-volume fusion across the GPUs of one box: one process per GPU, `torch.distributed` for the
plumbing.  The reference is single-process; this is the host side of the partitioned path the north star asks for:

  * integration needs no data-path collective: every rank sees the frame and updates the cubes it owns
    (`opb_volume_desc.shard_*`, slabs of `slab` cubes along `axis`, owner = floor(id/slab) mod world);
  * before Marching Cubes each rank sends the first voxel layer of the cubes at the low face of its slabs to the owner
    of the previous slab (`exchange_halo`), because GenerateMeshByCube (reference src/Integration/CubeHandler.cpp:83-99)
    reads the +x/+y/+z neighbour cubes;
  * the mesh is the concatenation of the per-rank meshes (TriangleMesh::LoadFromMeshes semantics,
    src/Geometry/TriangleMesh.cpp:73-94): `gather_mesh`.

Everything here is transport and bookkeeping; the arithmetic is in libonepiece_b200.so.  The functions take any "volume"
object with HaloCount / HaloExport / HaloImport, so the N>1 host logic is testable with `gloo` on CPU tensors."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

LAYER_FLOATS = 5 * 64  # sdf, weight, c0, c1, c2 of the 64 voxels of one cube layer


def owner_of(cube_ids, axis: int, slab: int, world: int):
    """Rank owning each cube: floor_mod(floor_div(id[axis], slab), world) -- the rule of select_kernel (csrc/opb_volume.cu)."""
    c = np.asarray(cube_ids)[..., axis]
    return np.mod(np.floor_divide(c, slab), world)


def halo_peers(rank: int, world: int):
    """(destination of this rank's boundary layer, source of the layer this rank needs)."""
    return (rank - 1) % world, (rank + 1) % world


def shard_range(n: int, rank: int, world: int):
    """Contiguous split of n items (source points of an ICP, image rows of the odometry) into world nearly equal parts."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def attach_halo_peers(volume, rank: int, world: int, cap_cubes: int, device_index: int, group=None):
    """Collective, once per volume: every rank creates its receive box for `cap_cubes` boundary cubes, the cudaIpc handles go
    round the group, and each rank maps the boxes of rank-1 (it exports into it) and rank+1 (it acknowledges there).  After
    this `exchange_halo` runs entirely on the device.  Returns the mapped pointers (hand them to `detach_halo_peers`)."""
    import ctypes as C

    from . import capi
    if world <= 1:
        return []
    own, handle = volume.HaloPeerBuffer(cap_cubes)
    handles = [None] * world
    dist.all_gather_object(handles, (handle, cap_cubes), group=group)
    dst, src = halo_peers(rank, world)
    mapped, opened = {}, []
    for r in {dst, src}:
        p = C.c_void_p()
        capi.check(capi.lib.opb_ipc_open(device_index, (C.c_ubyte * 64).from_buffer_copy(handles[r][0]), C.byref(p)))
        mapped[r] = p.value
        opened.append(p)
    volume.HaloPeerAttach(mapped[dst], handles[dst][1], mapped[src])
    dist.barrier(group=group)  # nobody exports before every box is mapped
    return opened


def detach_halo_peers(volume, opened, device_index: int, group=None):
    from . import capi
    if not opened:
        return
    volume.HaloPeerAttach(None, 0, None)
    dist.barrier(group=group)  # peers may still be writing into this rank's box until they detached too
    for p in opened:
        capi.lib.opb_ipc_close(device_index, p)


def attach_frame_ring(volume, rank: int, world: int, device_index: int, group=None):
    """Collective, once per volume: every rank creates its frame ring, the cudaIpc handles go round the group and every rank maps
    all of them.  After this `volume.IntegrateRowsAsync` uploads only this rank's band of a frame and the bands meet over NVLink.
    Returns the mapped pointers (hand them to `detach_frame_ring`)."""
    import ctypes as C

    from . import capi
    own, handle = volume.FrameRingBuffer()
    if world <= 1:
        volume.FrameRingAttach(0, 1, [own])
        return []
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    bufs, opened = [], []
    for r in range(world):
        if r == rank:
            bufs.append(own)
            continue
        p = C.c_void_p()
        capi.check(capi.lib.opb_ipc_open(device_index, (C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p)))
        opened.append(p)
        bufs.append(p.value)
    volume.FrameRingAttach(rank, world, bufs)
    dist.barrier(group=group)  # nobody scatters before every ring is mapped (and zeroed) everywhere
    return opened


def detach_frame_ring(volume, opened, device_index: int, group=None):
    from . import capi
    volume.FrameRingAttach(0, 0, None)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)  # peers may still be writing into this rank's ring until they detached too
    for p in opened:
        capi.lib.opb_ipc_close(device_index, p)


def exchange_halo(volume, rank: int, world: int, device="cpu", group=None) -> int:
    """Boundary-cube exchange before Marching Cubes.  Returns the number of ghost cubes imported.  Collective: every rank
    of the group must call it.  A volume with attached peer boxes (`attach_halo_peers`) does it in two kernel launches over
    peer memory; otherwise the buffers travel through torch.distributed send / recv (the path the gloo tests exercise)."""
    if world <= 1:
        return 0
    if getattr(volume, "HaloPeersAttached", None) and volume.HaloPeersAttached():
        return volume.HaloExchangePeer()[1]
    dst, src = halo_peers(rank, world)
    n_send = volume.HaloCount()
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    mine = torch.tensor([n_send], dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    n_recv = int(counts[src].item())
    ids_s = torch.empty((max(n_send, 1), 3), dtype=torch.int32, device=device)
    lay_s = torch.empty((max(n_send, 1), LAYER_FLOATS), dtype=torch.float32, device=device)
    if n_send:
        got = volume.HaloExport(ids_s.data_ptr(), lay_s.data_ptr(), n_send)
        assert got == n_send
    ids_r = torch.empty((max(n_recv, 1), 3), dtype=torch.int32, device=device)
    lay_r = torch.empty((max(n_recv, 1), LAYER_FLOATS), dtype=torch.float32, device=device)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    ops = []
    if n_send:
        ops += [dist.P2POp(dist.isend, ids_s[:n_send], dst, group), dist.P2POp(dist.isend, lay_s[:n_send], dst, group)]
    if n_recv:
        ops += [dist.P2POp(dist.irecv, ids_r[:n_recv], src, group), dist.P2POp(dist.irecv, lay_r[:n_recv], src, group)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    if n_recv:
        volume.HaloImport(ids_r.data_ptr(), lay_r.data_ptr(), n_recv)
    return n_recv


def gather_mesh(points, colors, rank: int, world: int, device="cpu", dst: int = 0, group=None):
    """Concatenation of the per-rank meshes on rank `dst` (three fresh vertices per triangle, so concatenating vertex
    arrays concatenates triangles).  Returns (points, colors) on dst, (None, None) elsewhere."""
    pts = torch.from_numpy(np.ascontiguousarray(points, np.float32).reshape(-1, 3))
    col = torch.from_numpy(np.ascontiguousarray(colors, np.float32).reshape(-1, 3))
    if world <= 1:
        return pts.numpy(), col.numpy()
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, torch.tensor([len(pts)], dtype=torch.int64, device=device), group=group)
    counts = [int(c) for c in counts.tolist()]
    payload = torch.cat([pts, col], 1).to(device)  # [n, 6]
    if rank != dst:
        if len(payload):
            dist.send(payload, dst, group=group)
        return None, None
    parts = []
    for r in range(world):
        if r == dst:
            parts.append(payload)
        elif counts[r]:
            buf = torch.empty((counts[r], 6), dtype=torch.float32, device=device)
            dist.recv(buf, r, group=group)
            parts.append(buf)
    allv = torch.cat(parts, 0).cpu().numpy()
    return np.ascontiguousarray(allv[:, :3]), np.ascontiguousarray(allv[:, 3:])


class ShardedCubeHandler:
    """One rank's part of a CubeHandler partitioned over the process group: same method names as the reference class
    (src/Integration/CubeHandler.h) for the calls the fusion mains make."""

    def __init__(self, camera, voxel_resolution=0.01, truncation=0.1, max_cubes=1 << 17, axis=0, slab=4, device_index=0,
                 group=None, halo_cubes=None, **kw):
        """halo_cubes: capacity of this rank's receive box for boundary cubes (default max_cubes / 4; 0: no peer boxes, the
        exchange then goes through torch.distributed send / recv).  Collective when the group has more than one rank."""
        from .volume import CubeHandler
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.device = torch.device("cuda", device_index)
        self.device_index = device_index
        self.volume = CubeHandler(camera, voxel_resolution, truncation, max_cubes=max_cubes, device=device_index,
                                  shard=(self.rank, self.world, axis, slab), **kw)
        cap = max(1024, max_cubes // 4) if halo_cubes is None else int(halo_cubes)
        self._halo_maps = attach_halo_peers(self.volume, self.rank, self.world, cap, device_index, group) if cap > 0 else []

    def close(self):
        """Collective: unmaps the neighbours' boxes, then releases the volume."""
        if self.volume is not None:
            if hasattr(self, "_ring_maps"):
                detach_frame_ring(self.volume, self._ring_maps, self.device_index, self.group)
                del self._ring_maps
            detach_halo_peers(self.volume, self._halo_maps, self.device_index, self.group)
            self._halo_maps = []
            self.volume.close()
            self.volume = None

    def IntegrateImage(self, depth, rgb, pose):
        """Every rank is given the same frame and pose (the host program broadcasts or loads it per rank)."""
        self.volume.IntegrateImage(depth, rgb, pose)

    def IntegrateImageRows(self, depth, rgb, pose):
        """Every rank holds the frame in host memory but uploads only its band of rows; the bands are exchanged over NVLink
        (opb_volume_integrate_rows_async).  Collective and asynchronous: `self.volume.Synchronize()` waits for the frame.  The
        host arrays must stay alive (and should be pinned) until then."""
        from .volume import depth_type_of, pose_colmajor
        if not hasattr(self, "_ring_maps"):
            self._ring_maps = attach_frame_ring(self.volume, self.rank, self.world, self.device_index, self.group)
        depth = np.ascontiguousarray(depth)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        lo, hi = shard_range(depth.shape[0], self.rank, self.world)
        self.volume.IntegrateRowsAsync(depth[lo:hi].ctypes.data, depth_type_of(depth), rgb[lo:hi].ctypes.data, lo, hi - lo, pose_colmajor(pose))

    def IntegrateImageBroadcast(self, depth, rgb, pose, src: int = 0):
        """The frame lives on rank `src` only (the others pass None): depth, colour and pose are broadcast over NCCL into
        device buffers and integrated from there -- no host copy on the receiving ranks."""
        from . import capi
        from .volume import depth_type_of, pose_colmajor
        cam = self.volume.camera
        meta = torch.zeros(17, dtype=torch.float32, device=self.device)  # depth type + pose
        if self.rank == src:
            depth = np.ascontiguousarray(depth)
            meta[0] = float(depth_type_of(depth))
            meta[1:] = torch.from_numpy(pose_colmajor(pose))
        if self.world > 1:
            dist.broadcast(meta, src, group=self.group)
        dtype = int(meta[0].item())
        np_dt = np.uint16 if dtype == capi.OPB_DEPTH_U16 else np.float32
        if self.rank == src:
            d_depth = torch.from_numpy(depth.view(np.int16) if np_dt is np.uint16 else depth).to(self.device)
            d_rgb = torch.from_numpy(np.ascontiguousarray(rgb, np.uint8)).to(self.device)
        else:
            d_depth = torch.empty((cam.height, cam.width), dtype=torch.int16 if np_dt is np.uint16 else torch.float32, device=self.device)
            d_rgb = torch.empty((cam.height, cam.width, 3), dtype=torch.uint8, device=self.device)
        if self.world > 1:
            dist.broadcast(d_depth.view(torch.uint8), src, group=self.group)  # raw bytes: NCCL has no 16-bit integer type
            dist.broadcast(d_rgb, src, group=self.group)
        torch.cuda.synchronize(self.device)   # the volume runs on its own stream
        self.volume.IntegrateImageDevice(d_depth.data_ptr(), dtype, d_rgb.data_ptr(), np.ascontiguousarray(meta[1:].cpu().numpy()))
        self.volume.Synchronize()             # the broadcast buffers may be released after this

    def ExtractTriangleMesh(self, dst: int = 0):
        """Halo exchange, per-rank Marching Cubes, concatenation on rank dst -> (points, colors, triangles) or Nones."""
        exchange_halo(self.volume, self.rank, self.world, self.device, self.group)
        pts, col, _ = self.volume.ExtractTriangleMesh()
        self.volume.HaloClear()
        p, c = gather_mesh(pts, col, self.rank, self.world, self.device, dst, self.group)
        if p is None:
            return None, None, None
        tri = np.arange(len(p), dtype=np.uint32).reshape(-1, 3)
        return p, c, tri


class SplitICP:
    """registration::PointToPlane / PointToPoint (reference src/Registration/ICP.cpp:31-224) with the source cloud split
    across the ranks of the process group: each rank searches and reduces its share, the 30-scalar packets meet in the
    reduction kernel over peer memory (include/onepiece_b200.h, opb_icp_comm_*).  Construction and every call are collective."""

    def __init__(self, device_index: int, group=None, stream=None):
        import ctypes as C

        from . import capi
        self._C, self._capi = C, capi
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device_index = device_index
        self.ws = C.c_void_p()
        capi.check(capi.lib.opb_icp_create(device_index, C.c_void_p(stream) if stream else None, C.byref(self.ws)))
        self._opened = []
        if self.world > 1:
            own = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            capi.check(capi.lib.opb_icp_comm_buffer(self.ws, C.byref(own), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            bufs = (C.c_void_p * self.world)()
            for r, h in enumerate(handles):
                if r == self.rank:
                    bufs[r] = own.value
                else:
                    p = C.c_void_p()
                    capi.check(capi.lib.opb_ipc_open(device_index, (C.c_ubyte * 64).from_buffer_copy(h), C.byref(p)))
                    self._opened.append(p)
                    bufs[r] = p.value
            capi.check(capi.lib.opb_icp_comm_attach(self.ws, self.rank, self.world, bufs))
            dist.barrier(group=group)  # nobody starts exchanging before every mailbox is mapped everywhere

    def close(self):
        capi = self._capi
        if self.ws:
            if self.world > 1:
                capi.lib.opb_icp_comm_detach(self.ws)
                dist.barrier(group=self.group)  # peers may still be reading this mailbox until they detached too
            for p in self._opened:
                capi.lib.opb_ipc_close(self.device_index, p)
            capi.lib.opb_icp_destroy(self.ws)
            self.ws = None

    def _run(self, source, target, init_T, icp_para, plane, gather_pairs):
        from . import registration as reg
        lo, hi = shard_range(len(source.points), self.rank, self.world)
        part = reg.PointCloud(source.points[lo:hi])
        r = reg._run(part, target, init_T, icp_para, plane, self.device_index, self.ws, want_pairs=gather_pairs)
        idx = r.correspondence_set_index  # empty when gather_pairs is False: the inliers then stay on the device
        idx[:, 0] += lo
        if gather_pairs and self.world > 1:
            parts = [None] * self.world
            dist.all_gather_object(parts, idx, group=self.group)
            idx = np.concatenate(parts)  # rank order == ascending source index, as the reference's push_back loop
        r.correspondence_set_index = idx
        r.correspondence_set = (source.points[idx[:, 0]], target.points[idx[:, 1]])
        return r

    def PointToPlane(self, source, target, init_T=np.eye(4), icp_para=None, gather_pairs=True):
        from . import registration as reg
        return self._run(source, target, init_T, icp_para or reg.ICPParameter(), True, gather_pairs)

    def PointToPoint(self, source, target, init_T=np.eye(4), icp_para=None, gather_pairs=True):
        from . import registration as reg
        return self._run(source, target, init_T, icp_para or reg.ICPParameter(), False, gather_pairs)
