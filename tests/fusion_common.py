"""Helpers shared by the multi-rank fusion tests: a CPU "volume" with the halo interface of
onepiece_b200.volume.CubeHandler built on the oracle (test infrastructure), and the gloo worker."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _view(ptr, ctype, count):
    return np.ctypeslib.as_array((ctype * count).from_address(int(ptr)))


def layer_of(vox, axis):
    """[n,512,5] AoS voxels -> [n,5,64] planes of the layer with axis coordinate 0 (element e = u + 8w as in opb_halo.cu)."""
    e = np.arange(64)
    u, w = e & 7, e >> 3
    vid = (u << 3) + (w << 6) if axis == 0 else (u + (w << 6) if axis == 1 else u + (w << 3))
    return np.ascontiguousarray(vox[:, vid, :].transpose(0, 2, 1)), vid


class OracleShard:
    """The cubes of a finished oracle volume that `rank` owns, plus ghosts received through the halo interface."""

    def __init__(self, cam, res, ids, vox, rank, world, axis, slab):
        from onepiece_b200.fusion import owner_of
        self.cam, self.res, self.axis, self.slab = cam, res, axis, slab
        mine = owner_of(ids, axis, slab, world) == rank
        self.ids, self.vox = ids[mine].copy(), vox[mine].copy()
        self.ghost_ids = np.zeros((0, 3), np.int32)
        self.ghost_vox = np.zeros((0, 512, 5), np.float32)

    def _boundary(self):
        return np.mod(self.ids[:, self.axis], self.slab) == 0

    def HaloCount(self):
        return int(self._boundary().sum())

    def HaloExport(self, ids_ptr, layers_ptr, cap):
        b = self._boundary()
        n = int(b.sum())
        assert n <= cap
        _view(ids_ptr, C.c_int32, n * 3)[:] = self.ids[b].reshape(-1)
        lay, _ = layer_of(self.vox[b], self.axis)
        _view(layers_ptr, C.c_float, n * 320)[:] = lay.reshape(-1)
        return n

    def HaloImport(self, ids_ptr, layers_ptr, n):
        ids = _view(ids_ptr, C.c_int32, n * 3).reshape(n, 3).copy()
        lay = _view(layers_ptr, C.c_float, n * 320).reshape(n, 5, 64).copy()
        vox = np.zeros((n, 512, 5), np.float32)
        vox[:, :, 0], vox[:, :, 1], vox[:, :, 2:] = 999.0, 0.0, -1.0  # TSDFVoxel defaults
        _, vid = layer_of(vox, self.axis)
        vox[:, vid, :] = lay.transpose(0, 2, 1)
        self.ghost_ids, self.ghost_vox = ids, vox

    def mesh(self):
        from oracle import oracleapi
        ov = oracleapi.OracleVolume(self.cam, self.res)
        ov.upload(np.concatenate([self.ids, self.ghost_ids]), np.concatenate([self.vox, self.ghost_vox]))
        return ov.extract_mesh()


def small_scene(n_frames=2):
    """S1 wall at quarter resolution with two poses -> finished oracle volume (ids, vox) and its mesh."""
    from onepiece_b200 import scenes
    from oracle import oracleapi
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0)
    res = 0.02
    ov = oracleapi.OracleVolume(cam, res)
    frames = []
    T = scenes.se3_exp([0.05, -0.02, 0.03, 0.02, -0.03, 0.01]).astype(np.float32)
    for k in range(n_frames):
        d, c = scenes.wavy_wall(cam, k)
        pose = np.eye(4, dtype=np.float32) if k == 0 else T
        ov.integrate(d, c, pose)
        frames.append((d, c, pose))
    ids, vox = ov.download()
    return cam, res, frames, ids, vox, ov.extract_mesh()


def gloo_worker(rank, world, port, axis, slab, out_path):
    """One rank of the CPU test: owned part of the oracle volume -> exchange_halo over gloo -> oracle MC -> gather."""
    import torch.distributed as dist

    from onepiece_b200 import fusion
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cam, res, _, ids, vox, _ = small_scene()
        shard = OracleShard(cam, res, ids, vox, rank, world, axis, slab)
        n_ghost = fusion.exchange_halo(shard, rank, world, "cpu")
        pts, col = shard.mesh()
        P, Cc = fusion.gather_mesh(pts, col, rank, world, "cpu", 0)
        stats = [None] * world
        dist.all_gather_object(stats, dict(rank=rank, owned=len(shard.ids), sent=shard.HaloCount(), ghosts=n_ghost, verts=len(pts)))
        if rank == 0:
            np.savez(out_path, points=P, colors=Cc, stats=np.array([[s["owned"], s["sent"], s["ghosts"], s["verts"]] for s in stats]))
    finally:
        dist.destroy_process_group()
