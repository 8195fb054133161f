"""torchrun worker of tests/test_fusion_gpu.py (one process per GPU, NCCL): partitioned fusion + halo exchange + Marching
Cubes against the unsharded volume, and the split ICP against the single-GPU call.  Prints "MGPU OK" on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    from conftest import assert_bit_equal, canon_triangles
    from fusion_common import small_scene
    from onepiece_b200 import fusion, registration as reg
    from onepiece_b200.volume import CubeHandler
    from test_fusion_gpu import _icp_inputs
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cam, res, frames, ids, vox, (opts, ocol) = small_scene()
    # 1. partitioned fusion: every rank integrates the same frames into the cubes it owns; halo; MC; gather
    sh = fusion.ShardedCubeHandler(cam, res, max_cubes=4096, axis=0, slab=2, device_index=local)
    for d, c, pose in frames:
        sh.IntegrateImage(d, c, pose)
    P, Cc, tri = sh.ExtractTriangleMesh(0)
    owned = torch.tensor([sh.volume.NumCubes()], device="cuda")
    dist.all_reduce(owned)
    if rank == 0:
        assert int(owned.item()) == len(ids), (int(owned.item()), len(ids))
        assert np.array_equal(canon_triangles(P, Cc), canon_triangles(opts, ocol)), "sharded mesh differs from the oracle mesh"
        assert len(tri) * 3 == len(P)
    assert sh.volume.HaloPeersAttached()   # the exchange above went through the peer boxes, not through send / recv
    # ... and the same exchange through torch.distributed send / recv imports the same number of ghosts
    n_peer = fusion.exchange_halo(sh.volume, rank, world, sh.device)
    sh.volume.HaloClear()
    maps, sh._halo_maps = sh._halo_maps, []
    fusion.detach_halo_peers(sh.volume, maps, local)
    n_nccl = fusion.exchange_halo(sh.volume, rank, world, sh.device)
    assert n_peer == n_nccl > 0, (n_peer, n_nccl)
    sh.close()
    # 2. split ICP over peer memory vs the single-GPU call on this rank's device
    src, tgt, nrm = _icp_inputs(quarter=False)
    par = reg.ICPParameter(10, 0.05, 1.0)
    whole = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), par, device=local)
    sp = fusion.SplitICP(local)
    for _ in range(2):
        r = sp.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), par)
    assert np.array_equal(r.correspondence_set_index, whole.correspondence_set_index), "split ICP pairs differ"
    assert np.abs(r.T_iterated - whole.T_iterated).max() < 1e-6 and np.abs(r.T - whole.T).max() < 1e-6
    Ts = [None] * world
    dist.all_gather_object(Ts, r.T_iterated.tobytes())
    assert all(t == Ts[0] for t in Ts), "ranks disagree on the pose"
    # timing of the split call vs the whole call (reported, not asserted)
    import time
    for name, fn in (("whole", lambda: reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), par, device=local)),
                     ("split", lambda: sp.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), par, gather_pairs=False))):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            print(f"ICP {name}: {(time.perf_counter() - t0) * 100:.3f} ms per call ({len(src)} source points, world {world})")
    sp.close()
    # 3. BASELINE.json config 4 at one quarter of its GPU count: 1280x960 depth, 2 mm voxels, volume partitioned over the ranks
    #    by slabs of 8 cubes, boundary-cube exchange over NCCL, Marching Cubes per rank -- the vertex count must be the
    #    unsharded volume's (north_star: "Marching-Cubes mesh vertex count identical")
    from onepiece_b200 import scenes
    c0 = scenes.Camera()
    big = scenes.Camera(2 * c0.fx, 2 * c0.fy, 2 * c0.cx, 2 * c0.cy, 1280, 960, 1000.0)
    sh4 = fusion.ShardedCubeHandler(big, 0.002, max_cubes=1 << 18, axis=0, slab=8, device_index=local)
    whole = CubeHandler(big, 0.002, max_cubes=1 << 18, device=local) if rank == 0 else None
    for k in range(2):
        d, c = scenes.wavy_wall(big, k)
        sh4.IntegrateImage(d, c, np.eye(4, dtype=np.float32))
        if whole is not None:
            whole.IntegrateImage(d, c, np.eye(4, dtype=np.float32))
    bare = sh4.volume.CountMesh()[0]
    n_ghost = fusion.exchange_halo(sh4.volume, rank, world, sh4.device)
    mine = sh4.volume.CountMesh()[0]
    tot = torch.tensor([sh4.volume.NumCubes(), bare, mine, n_ghost], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)
    if rank == 0:
        cubes, bare_sum, full_sum, ghosts = (int(x) for x in tot.tolist())
        assert cubes == whole.NumCubes(), (cubes, whole.NumCubes())
        assert full_sum == whole.CountMesh()[0] > bare_sum, (full_sum, whole.CountMesh()[0], bare_sum)
        print(f"config 4 slice: {cubes} cubes over {world} ranks, {ghosts} boundary cubes exchanged "
              f"({ghosts * 1292 / 1e6:.1f} MB instead of {ghosts * 10240 / 1e6:.1f} MB), {full_sum} mesh vertices = unsharded")
    sh4.close()
    dist.barrier()
    # 4. BASELINE.json config 5 in miniature: a DenseFusion-style loop -- per frame one ICP split over the ranks (6x6 packet
    #    exchanged over peer memory), the pose chained, the frame integrated into the partitioned volume -- then halo exchange
    #    and Marching Cubes.  Every rank must chain the identical pose; with those poses the partitioned volume and its mesh
    #    must be exactly what one GPU produces.
    cam5 = scenes.Camera(c0.fx / 2, c0.fy / 2, c0.cx / 2, c0.cy / 2, 320, 240, 1000.0)
    sh5 = fusion.ShardedCubeHandler(cam5, 0.01, max_cubes=1 << 15, axis=0, slab=4, device_index=local)
    whole5 = CubeHandler(cam5, 0.01, max_cubes=1 << 15, device=local) if rank == 0 else None
    sp5 = fusion.SplitICP(local)
    pose = np.eye(4)
    prev = None
    for k in range(5):
        d, c, T_true, n = scenes.room(cam5, 2 * k, with_normals=True)
        cloud = scenes.backproject(d, cam5)
        nrm5 = np.ascontiguousarray(n.reshape(-1, 3)[(d > 0).reshape(-1)])
        if prev is not None:
            r5 = sp5.PointToPlane(reg.PointCloud(cloud), reg.PointCloud(prev[0], prev[1]), np.eye(4), reg.ICPParameter(10, 0.05, 1.0),
                                  gather_pairs=False)
            pose = pose @ r5.T_iterated.astype(np.float64)
        prev = (cloud, nrm5)
        p32 = pose.astype(np.float32)
        if k % 3 == 1:
            sh5.IntegrateImage(d, c, p32)               # every rank loaded the frame itself
        elif k % 3 == 2:
            sh5.IntegrateImageRows(d, c, p32)           # every rank uploads its band of rows, the bands meet over NVLink
            sh5.volume.Synchronize()
            sh5.volume.FrameRingStatus()
        else:
            sh5.IntegrateImageBroadcast(d if rank == 0 else None, c if rank == 0 else None, p32 if rank == 0 else None, src=0)
        if whole5 is not None:
            whole5.IntegrateImage(d, c, p32)
    poses = [None] * world
    dist.all_gather_object(poses, pose.tobytes())
    assert all(p == poses[0] for p in poses), "ranks chained different poses"
    err = np.linalg.norm(pose[:3, 3] - T_true[:3, 3])
    P5, C5, _ = sh5.ExtractTriangleMesh(0)
    gi, gv = sh5.volume.GetCubeMap()
    parts = [None] * world
    dist.all_gather_object(parts, (gi, gv))
    if rank == 0:
        ai = np.concatenate([p[0] for p in parts]); av = np.concatenate([p[1] for p in parts])
        order = np.lexsort((ai[:, 2], ai[:, 1], ai[:, 0]))
        wi, wv = whole5.GetCubeMap()
        assert np.array_equal(ai[order], wi)
        assert_bit_equal(av[order], wv, "partitioned volume after the fusion loop")
        wp, wc, _ = whole5.ExtractTriangleMesh()
        assert np.array_equal(canon_triangles(P5, C5), canon_triangles(wp, wc)), "partitioned mesh differs"
        print(f"config 5 slice: 5 frames tracked with split ICP (drift {1e3 * err:.2f} mm vs ground truth), {len(wi)} cubes, "
              f"{len(wp)} mesh vertices identical to the single-GPU pipeline")
    sp5.close()
    sh5.close()
    dist.barrier()
    if rank == 0:
        print("MGPU OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
