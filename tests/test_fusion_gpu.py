"""Partitioned fusion on the device: halo exchange + Marching Cubes over sharded volumes, and the ICP whose
source points are split across ranks with the 6x6 system exchanged through peer memory.  The single-device tests run two
ranks on one GPU (two volumes / two solver workspaces on their own streams); the torchrun tests need >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal, canon_triangles
from fusion_common import small_scene

pytestmark = pytest.mark.gpu


def _n_gpus():
    from onepiece_b200 import capi
    return capi.lib.opb_device_count()


@pytest.mark.parametrize("world,axis,slab", [(2, 0, 2), (3, 1, 1), (2, 2, 4)])
def test_halo_exchange_between_shards_on_one_device(world, axis, slab):
    from onepiece_b200.volume import CubeHandler
    cam, res, frames, ids, vox, (opts, ocol) = small_scene()
    full = CubeHandler(cam, res, max_cubes=4096)
    shards = [CubeHandler(cam, res, max_cubes=4096, shard=(r, world, axis, slab)) for r in range(world)]
    for d, c, pose in frames:
        full.IntegrateImage(d, c, pose)
        for s in shards:
            s.IntegrateImage(d, c, pose)
    fpts, fcol, _ = full.ExtractTriangleMesh()
    assert np.array_equal(canon_triangles(fpts, fcol), canon_triangles(opts, ocol))
    # without the exchange the seams between slabs are missing
    bare = sum(s.CountMesh()[0] for s in shards)
    assert bare < len(fpts)
    # export -> (transport) -> import into the owner of the previous slab
    packets = []
    for s in shards:
        n = s.HaloCount()
        bi, bl = np.zeros((max(n, 1), 3), np.int32), np.zeros((max(n, 1), 5, 64), np.float32)
        assert s.HaloExport(bi, bl, n) == n
        packets.append((bi[:n], bl[:n]))
    for r, s in enumerate(shards):
        bi, bl = packets[(r + 1) % world]
        s.HaloImport(np.ascontiguousarray(bi), np.ascontiguousarray(bl), len(bi))
        assert s.NumGhostCubes() == len(bi)
    parts = [s.ExtractTriangleMesh() for s in shards]
    P = np.concatenate([p[0] for p in parts])
    Cc = np.concatenate([p[1] for p in parts])
    assert np.array_equal(canon_triangles(P, Cc), canon_triangles(fpts, fcol))
    # ghosts are not part of the volume: the union of the downloads is still exactly the unsharded volume ...
    gi = np.concatenate([s.GetCubeMap()[0] for s in shards])
    gv = np.concatenate([s.GetCubeMap()[1] for s in shards])
    order = np.lexsort((gi[:, 2], gi[:, 1], gi[:, 0]))
    fi, fv = full.GetCubeMap()
    assert np.array_equal(gi[order], fi)
    assert_bit_equal(gv[order], fv, "sharded union after halo import")
    # ... and integrating another frame drops them and stays bit-exact
    d, c, pose = frames[-1]
    full.IntegrateImage(d, c, pose)
    for s in shards:
        s.IntegrateImage(d, c, pose)
        assert s.NumGhostCubes() == 0
    gi = np.concatenate([s.GetCubeMap()[0] for s in shards])
    gv = np.concatenate([s.GetCubeMap()[1] for s in shards])
    order = np.lexsort((gi[:, 2], gi[:, 1], gi[:, 0]))
    fi, fv = full.GetCubeMap()
    assert np.array_equal(gi[order], fi)
    assert_bit_equal(gv[order], fv, "sharded union after a further frame")
    assert sum(s.CountMesh()[0] for s in shards) < full.CountMesh()[0]


@pytest.mark.parametrize("world,axis,slab", [(2, 0, 2), (3, 1, 1)])
def test_halo_exchange_through_peer_boxes_on_one_device(world, axis, slab):
    """The exchange with the transport inside the kernels: every shard exports straight into the receive box of the previous
    shard and imports from its own (plain device pointers here, cudaIpc mappings between processes).  Same mesh as the
    unsharded volume, twice in a row (the acknowledgement lets the second export overwrite the box)."""
    from onepiece_b200 import capi
    from onepiece_b200.volume import CubeHandler
    cam, res, frames, ids, vox, (opts, ocol) = small_scene()
    full = CubeHandler(cam, res, max_cubes=4096)
    shards = [CubeHandler(cam, res, max_cubes=4096, shard=(r, world, axis, slab)) for r in range(world)]
    boxes = [s.HaloPeerBuffer(2048)[0] for s in shards]
    for r, s in enumerate(shards):
        s.HaloPeerAttach(boxes[(r - 1) % world], 2048, boxes[(r + 1) % world])
        assert s.HaloPeersAttached()
    for rounds in range(2):
        d, c, pose = frames[rounds] if rounds < len(frames) else frames[-1]
        full.IntegrateImage(d, c, pose)
        for s in shards:
            s.IntegrateImage(d, c, pose)
            assert s.NumGhostCubes() == 0
        expect = [s.HaloCount() for s in shards]
        for s in shards:
            s.HaloExchangeBegin()          # one host thread drives all ranks: enqueue everywhere, then collect
        got = [s.HaloExchangeEnd() for s in shards]
        for r in range(world):
            assert got[r][0] == expect[r]
            assert got[r][1] == expect[(r + 1) % world] == shards[r].NumGhostCubes()
        fpts, fcol, _ = full.ExtractTriangleMesh()
        parts = [s.ExtractTriangleMesh() for s in shards]
        P = np.concatenate([p[0] for p in parts])
        Cc = np.concatenate([p[1] for p in parts])
        assert np.array_equal(canon_triangles(P, Cc), canon_triangles(fpts, fcol))
    # a box that is too small is an error, not a truncated mesh
    small = [CubeHandler(cam, res, max_cubes=4096, shard=(r, 2, 0, 1)) for r in range(2)]
    sb = [s.HaloPeerBuffer(4)[0] for s in small]
    for r, s in enumerate(small):
        s.HaloPeerAttach(sb[1 - r], 4, sb[1 - r])
        s.IntegrateImage(*frames[0])
    for s in small:
        s.HaloExchangeBegin()
    for s in small:
        with pytest.raises(capi.OpbError) as e:
            s.HaloExchangeEnd()
        assert e.value.code == capi.OPB_ERR_CAPACITY


@pytest.mark.parametrize("world", [2, 3])
def test_frame_uploaded_in_row_bands_and_gathered_over_peer_memory(world):
    """opb_volume_integrate_rows_async: every shard uploads only its band of rows, the bands meet in every shard's frame ring
    (plain device pointers here, cudaIpc mappings between processes).  The union of the shards equals the unsharded volume bit for
    bit, over more frames than the ring has slots (the consumed flags gate the reuse of a slot)."""
    from onepiece_b200 import capi
    from onepiece_b200.fusion import shard_range
    from onepiece_b200.volume import CubeHandler, depth_type_of, pose_colmajor
    cam, res, frames, ids, vox, _ = small_scene(5)
    full = CubeHandler(cam, res, max_cubes=4096)
    shards = [CubeHandler(cam, res, max_cubes=4096, shard=(r, world, 0, 2)) for r in range(world)]
    rings = [s.FrameRingBuffer()[0] for s in shards]
    for r, s in enumerate(shards):
        s.FrameRingAttach(r, world, rings)
    for d, c, pose in frames:
        full.IntegrateImage(d, c, pose)
        d = np.ascontiguousarray(d); c = np.ascontiguousarray(c, np.uint8)
        for r, s in enumerate(shards):      # one host thread drives all ranks: every call only enqueues
            lo, hi = shard_range(d.shape[0], r, world)
            s.IntegrateRowsAsync(d[lo:hi].ctypes.data, depth_type_of(d), c[lo:hi].ctypes.data, lo, hi - lo, pose_colmajor(pose))
        for s in shards:
            s.Synchronize()                 # (the host arrays of this frame are released after this)
    for s in shards:
        s.FrameRingStatus()
    gi = np.concatenate([s.GetCubeMap()[0] for s in shards])
    gv = np.concatenate([s.GetCubeMap()[1] for s in shards])
    order = np.lexsort((gi[:, 2], gi[:, 1], gi[:, 0]))
    fi, fv = full.GetCubeMap()
    assert np.array_equal(gi[order], fi)
    assert_bit_equal(gv[order], fv, "shards fed by row bands")
    # a band that does not fit the image is refused before anything is enqueued
    with pytest.raises(capi.OpbError) as e:
        shards[0].IntegrateRowsAsync(d.ctypes.data, depth_type_of(d), c.ctypes.data, d.shape[0] - 1, 2, pose_colmajor(pose))
    assert e.value.code == capi.OPB_ERR_INVALID
    for s in shards:
        s.FrameRingAttach(0, 0, None)


def test_halo_capacity_and_unsharded_volume():
    from onepiece_b200 import capi
    from onepiece_b200.volume import CubeHandler
    cam, res, frames, *_ = small_scene(1)
    v = CubeHandler(cam, res, max_cubes=4096)
    v.IntegrateImage(*frames[0])
    assert v.HaloCount() == 0                       # world 1: nothing lives elsewhere
    s = CubeHandler(cam, res, max_cubes=v.NumCubes(), shard=(0, 2, 0, 1))
    s.IntegrateImage(*frames[0])
    n = s.HaloCount()
    assert n == s.NumCubes() > 0                    # slab of one cube: every cube is a boundary cube
    bi, bl = np.zeros((n, 3), np.int32), np.zeros((n, 5, 64), np.float32)
    with pytest.raises(capi.OpbError) as e:
        s.HaloExport(bi, bl, n - 1)
    assert e.value.code == capi.OPB_ERR_CAPACITY
    s.HaloExport(bi, bl, n)
    big = np.zeros((4096, 3), np.int32)
    big[:, 1] = np.arange(4096) + 1000
    with pytest.raises(capi.OpbError) as e:
        s.HaloImport(big, np.zeros((4096, 5, 64), np.float32), 4096)
    assert e.value.code == capi.OPB_ERR_CAPACITY


def _torchrun(nproc, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", script), *args]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:]
    return r.stdout


def test_torchrun_two_gpus_fusion_and_split_icp():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    out = _torchrun(2, "mgpu_worker.py")
    assert "MGPU OK" in out, out[-4000:]


def _icp_inputs(step=3, quarter=True):
    from onepiece_b200 import scenes
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0) if quarter else c0
    d0, _, _, n0 = scenes.room(cam, 0, with_normals=True)
    d1, _, _ = scenes.room(cam, step)
    tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)
    nrm = np.ascontiguousarray(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)])
    return src, tgt, nrm


@pytest.mark.parametrize("world,plane", [(2, True), (3, True), (2, False)])
def test_split_icp_workspaces_on_one_device(world, plane):
    """Source points split over `world` solver workspaces that exchange their packets through each other's mailboxes (plain
    device pointers here; cudaIpc mappings between processes): same inlier pairs, same pose as the unsplit call."""
    import ctypes as C
    import threading

    from onepiece_b200 import capi, registration as reg
    from onepiece_b200.fusion import shard_range
    src, tgt, nrm = _icp_inputs()
    par = reg.ICPParameter(8, 0.05, 1.0)
    run = reg.PointToPlane if plane else reg.PointToPoint
    target = reg.PointCloud(tgt, nrm if plane else None)
    whole = run(reg.PointCloud(src), target, np.eye(4), par)
    ws, bufs = [], (C.c_void_p * world)()
    for r in range(world):
        h, b = C.c_void_p(), C.c_void_p()
        capi.check(capi.lib.opb_icp_create(0, None, C.byref(h)))
        capi.check(capi.lib.opb_icp_comm_buffer(h, C.byref(b), None))
        capi.check(capi.lib.opb_icp_reserve(h, len(src), len(tgt)))  # no allocation inside the collective calls (ranks share the device)
        ws.append(h)
        bufs[r] = b.value
    for r in range(world):
        capi.check(capi.lib.opb_icp_comm_attach(ws[r], r, world, bufs))
    out, errs = [None] * world, []

    def rank_main(r):
        try:
            lo, hi = shard_range(len(src), r, world)
            for _ in range(2):  # two collective calls in a row: the mailbox epochs carry over
                out[r] = (lo, run(reg.PointCloud(src[lo:hi]), target, np.eye(4), par, workspace=ws[r]))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(60) for t in th]
    assert not errs, errs
    pairs = []
    for lo, r in out:
        assert np.array_equal(r.T, out[0][1].T) and np.array_equal(r.T_iterated, out[0][1].T_iterated)  # identical on every rank
        assert r.rmse == out[0][1].rmse
        p = r.correspondence_set_index.copy()
        p[:, 0] += lo
        pairs.append(p)
    assert np.array_equal(np.concatenate(pairs), whole.correspondence_set_index)
    assert np.abs(out[0][1].T_iterated - whole.T_iterated).max() < 1e-6
    assert np.abs(out[0][1].T - whole.T).max() < 1e-6
    # (the whole call sums exact products in the persistent loop's 8x8 accumulation, the split one float-rounded products:
    # the iterated pose may differ in its last float bit, the rmse with it)
    assert abs(out[0][1].rmse - whole.rmse) < 2e-6 * whole.rmse
    # a collective call that a peer never makes fails after the time limit instead of hanging
    if world == 2 and plane:
        with pytest.raises(capi.OpbError) as e:
            run(reg.PointCloud(src[:100]), target, np.eye(4), reg.ICPParameter(1, 0.05, 1.0), workspace=ws[0])
        assert e.value.code == capi.OPB_ERR_CUDA
    for h in ws:
        capi.lib.opb_icp_comm_detach(h)
        capi.lib.opb_icp_destroy(h)
