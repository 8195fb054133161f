"""Parity of the CUDA TSDF volume (through the C-ABI) with the oracle: bit-exact voxels, cube sets, meshes."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_bit_equal, canon_triangles, sha
from onepiece_b200 import capi, scenes
from onepiece_b200.volume import CubeHandler
from oracle import oracleapi

pytestmark = pytest.mark.gpu

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "integrate_*.npz")))


def small_camera():
    c = scenes.Camera()
    return scenes.Camera(c.fx / 4, c.fy / 4, c.cx / 4, c.cy / 4, 160, 120, 1000.0)


def random_pose(rng, scale=1.0):
    return scenes.se3_exp(rng.normal(0, 1, 6) * np.array([1, 1, 1, .5, .5, .5]) * scale).astype(np.float32)


def compare_volumes(gpu: CubeHandler, ov: oracleapi.OracleVolume):
    gi, gv = gpu.GetCubeMap()
    oi, ovx = ov.download()
    assert np.array_equal(gi, oi), f"cube sets differ: {len(gi)} vs {len(oi)}"
    assert_bit_equal(gv, ovx, "voxels")


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_golden_fixtures(path):
    g = np.load(path)
    c = g["cam"]
    cam = scenes.Camera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5]), float(c[6]))
    gpu = CubeHandler(cam, float(g["res"]), float(g["trunc"]), max_cubes=4096)
    for d, col, T in zip(g["depth"], g["bgr"], g["poses"]):
        gpu.IntegrateImage(d, col, T)
    ids, vox = gpu.GetCubeMap()
    assert np.array_equal(ids, g["ids"])
    assert_bit_equal(vox[:16], g["voxels_head"], "first cubes")
    assert sha(vox) == str(g["voxels_sha"])
    pts, colr, tri = gpu.ExtractTriangleMesh()
    assert len(pts) == int(g["n_vertices"]) and len(tri) == int(g["n_triangles"])
    canon = canon_triangles(pts, colr)
    assert_bit_equal(canon[:64], g["mesh_head"], "first triangles")
    assert sha(canon) == str(g["mesh_sha"])


@pytest.mark.parametrize("u16", [False, True])
def test_small_sequence_vs_oracle(u16):
    cam = small_camera()
    rng = np.random.default_rng(11)
    gpu, ov = CubeHandler(cam, 0.02, max_cubes=8192), oracleapi.OracleVolume(cam, 0.02)
    for k in range(6):
        d, c = scenes.wavy_wall(cam, k)
        if u16:
            d = np.clip(np.rint(d * 1000), 0, 65535).astype(np.uint16)
        T = random_pose(rng, 0.15 if k else 0.0)
        gpu.IntegrateImage(d, c, T)
        n = ov.integrate(d, c, T)
        st = gpu.FrameStats()
        assert st.frame_cubes == n and st.overflow == 0
        mx, mn = ov.bounding(d, T)
        assert_bit_equal(np.array(st.bbox_max[:], np.float32), mx, "bbox max")
        assert_bit_equal(np.array(st.bbox_min[:], np.float32), mn, "bbox min")
    compare_volumes(gpu, ov)
    pts, col, tri = gpu.ExtractTriangleMesh()
    op, oc = ov.extract_mesh()
    assert len(pts) == len(op) and len(tri) * 3 == len(pts)
    assert np.array_equal(tri.reshape(-1), np.arange(len(pts), dtype=np.uint32))
    assert_bit_equal(canon_triangles(pts, col), canon_triangles(op, oc), "mesh")
    assert gpu.CountMesh() == (len(pts), len(tri))


def test_full_size_config1_vs_oracle():
    """BASELINE config 1 shape: 640x480, 5 mm voxels, identity poses (3 frames keep the CPU oracle to ~1 s)."""
    cam = scenes.Camera()
    gpu, ov = CubeHandler(cam, 0.005, max_cubes=1 << 15), oracleapi.OracleVolume(cam, 0.005)
    I = np.eye(4, dtype=np.float32)
    for k in range(3):
        d, c = scenes.wavy_wall(cam, k)
        gpu.IntegrateImage(d, c, I)
        ov.integrate(d, c, I)
    st = gpu.FrameStats()
    assert st.total_cubes == ov.num_cubes()
    compare_volumes(gpu, ov)
    nv, nt = gpu.CountMesh()
    op, _ = ov.extract_mesh()
    assert nv == len(op) and nt * 3 == nv


def test_full_size_properties_50_frames():
    """Size-independent properties at BASELINE's full workload (50 frames, too slow for the CPU oracle):
    weights are integers <= frames, weight sum == total updated voxels, re-running is deterministic."""
    cam = scenes.Camera()
    I = np.eye(4, dtype=np.float32)
    frames = [scenes.wavy_wall(cam, k) for k in range(50)]
    results = []
    for _ in range(2):
        gpu = CubeHandler(cam, 0.005, max_cubes=1 << 15)
        upd = 0
        for d, c in frames:
            gpu.IntegrateImage(d, c, I)
            upd += gpu.FrameStats().updated_voxels
        ids, vox = gpu.GetCubeMap()
        w = vox[..., 1]
        assert np.array_equal(w, np.round(w)) and w.max() == 50 and w.min() == 0
        assert int(w.sum(dtype=np.float64)) == upd
        valid = w > 0
        assert np.all(np.abs(vox[..., 0][valid]) < 0.1)
        assert np.all((vox[..., 2:][valid] >= 0) & (vox[..., 2:][valid] <= 1))
        assert np.all(vox[..., 0][~valid] == 999) and np.all(vox[..., 2:][~valid] == -1)
        results.append((ids, vox))
        gpu.close()
    assert np.array_equal(results[0][0], results[1][0])
    assert_bit_equal(results[0][1], results[1][1], "run-to-run determinism")


def test_empty_and_degenerate_inputs():
    cam = small_camera()
    gpu, ov = CubeHandler(cam, 0.02, max_cubes=4096), oracleapi.OracleVolume(cam, 0.02)
    I = np.eye(4, dtype=np.float32)
    _, c = scenes.wavy_wall(cam, 0)
    # all-zero depth: no point passes z > 0, nothing is allocated
    gpu.IntegrateImage(np.zeros((cam.height, cam.width), np.float32), c, I)
    assert gpu.NumCubes() == 0 and gpu.FrameStats().frame_cubes == 0
    assert gpu.ExtractTriangleMesh()[0].shape == (0, 3)
    # NaN / negative / beyond-far depth
    d, _ = scenes.wavy_wall(cam, 0)
    d2 = d.copy()
    d2[::7, ::5] = np.nan
    d2[3::11, 1::3] = -1.0
    d2[:, :20] = 7.5  # beyond the far plane: never inside the frustum, but still integrated if listed
    gpu.IntegrateImage(d2, c, I)
    ov.integrate(d2, c, I)
    compare_volumes(gpu, ov)
    # camera looking away: surface entirely outside
    T = scenes.se3_exp([0, 0, 0, 0, np.pi, 0]).astype(np.float32)
    gpu.IntegrateImage(d, c, T)
    ov.integrate(d, c, T)
    compare_volumes(gpu, ov)


def test_truncation_above_one_overwrites_large_sdf():
    """TSDFVoxel::IsValid treats sdf >= 1 as invalid (TSDFVoxel.h:75-78): with truncation > 1 such voxels are
    overwritten instead of averaged.  The CUDA path must reproduce that quirk."""
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 8, c0.fy / 8, c0.cx / 8, c0.cy / 8, 80, 60, 1000.0)
    gpu, ov = CubeHandler(cam, 0.08, truncation=1.5, max_cubes=8192), oracleapi.OracleVolume(cam, 0.08, 1.5)
    I = np.eye(4, dtype=np.float32)
    for k in range(3):
        d, c = scenes.wavy_wall(cam, k)
        gpu.IntegrateImage(d, c, I)
        ov.integrate(d, c, I)
    compare_volumes(gpu, ov)


def test_prepare_cubes_matches_oracle_set():
    cam = small_camera()
    gpu, ov = CubeHandler(cam, 0.02, max_cubes=4096), oracleapi.OracleVolume(cam, 0.02)
    d, _ = scenes.wavy_wall(cam, 2)
    T = scenes.se3_exp([0.1, 0, -0.1, 0.05, 0.1, 0]).astype(np.float32)
    a = gpu.PrepareCubes(d, T)
    b = ov.prepare_cubes(d, T)
    key = lambda x: x[np.lexsort((x[:, 2], x[:, 1], x[:, 0]))]
    assert np.array_equal(key(a), key(b))
    assert gpu.NumCubes() == ov.num_cubes()


def test_upload_download_roundtrip_and_clear():
    cam = small_camera()
    gpu, ov = CubeHandler(cam, 0.02, max_cubes=4096), oracleapi.OracleVolume(cam, 0.02)
    I = np.eye(4, dtype=np.float32)
    d, c = scenes.wavy_wall(cam, 0)
    ov.integrate(d, c, I)
    ids, vox = ov.download()
    perm = np.random.default_rng(0).permutation(len(ids))
    gpu.SetCubeMap(ids[perm], vox[perm])
    compare_volumes(gpu, ov)
    # continue integrating on top of uploaded content
    d1, c1 = scenes.wavy_wall(cam, 1)
    gpu.IntegrateImage(d1, c1, I)
    ov.integrate(d1, c1, I)
    compare_volumes(gpu, ov)
    gpu.Clear()
    assert gpu.NumCubes() == 0
    gpu.IntegrateImage(d, c, I)
    ov.clear()
    ov.integrate(d, c, I)
    compare_volumes(gpu, ov)


def test_setters_take_effect():
    cam = small_camera()
    gpu = CubeHandler(cam)  # reference defaults 0.01 / 0.1
    gpu.SetVoxelResolution(0.04)
    gpu.SetTruncation(0.2)
    gpu.SetNearPlane(1.8)
    gpu.SetFarPlane(2.2)
    ov = oracleapi.OracleVolume(cam, 0.04, 0.2, 1.8, 2.2)
    d, c = scenes.wavy_wall(cam, 0)
    gpu.IntegrateImage(d, c, np.eye(4))
    ov.integrate(d, c, np.eye(4))
    compare_volumes(gpu, ov)


def test_a_full_pool_grows_instead_of_dropping_cubes():
    """The reference's cube map is unbounded (CubeHandler.cpp:181-191).  A pool that is far too small grows under the
    synchronous calls and the frame is completed for the cubes that found no slot: same volume as with a roomy pool."""
    cam = small_camera()
    rng = np.random.default_rng(5)
    small = CubeHandler(cam, 0.02, max_cubes=64)  # ~1000 cubes needed
    roomy = CubeHandler(cam, 0.02, max_cubes=8192)
    ov = oracleapi.OracleVolume(cam, 0.02)
    for k in range(3):
        d, c = scenes.wavy_wall(cam, k)
        T = random_pose(rng, 0.1 if k else 0.0)
        small.IntegrateImage(d, c, T)
        roomy.IntegrateImage(d, c, T)
        n = ov.integrate(d, c, T)
        st = small.FrameStats()
        assert st.frame_cubes == n == roomy.FrameStats().frame_cubes
        assert st.updated_voxels == roomy.FrameStats().updated_voxels
    assert small.FrameStats().overflow >= 4 and roomy.FrameStats().overflow == 0   # 64 -> 128 -> ... -> 1024+
    a, b = small.GetCubeMap(), roomy.GetCubeMap()
    oa, ob = np.lexsort(a[0].T[::-1]), np.lexsort(b[0].T[::-1])
    assert np.array_equal(a[0][oa], b[0][ob])
    assert_bit_equal(a[1][oa], b[1][ob], "voxels of the grown volume")
    oi, ovx = ov.download()
    oo = np.lexsort(oi.T[::-1])
    assert np.array_equal(a[0][oa], oi[oo])
    assert_bit_equal(a[1][oa], ovx[oo], "grown volume vs oracle")
    # PrepareCubes and SetCubeMap grow as well
    tiny = CubeHandler(cam, 0.02, max_cubes=16)
    d, c = scenes.wavy_wall(cam, 0)
    ids = tiny.PrepareCubes(d, np.eye(4))
    assert len(ids) == len(roomy.PrepareCubes(d, np.eye(4))) and len(ids) > 16
    tiny2 = CubeHandler(cam, 0.02, max_cubes=16)
    tiny2.SetCubeMap(*b)
    c2 = tiny2.GetCubeMap()
    assert np.array_equal(c2[0], b[0]) and np.array_equal(c2[1].view(np.uint32), b[1].view(np.uint32))


def test_compute_bounding_is_read_only():
    cam = small_camera()
    d, c = scenes.wavy_wall(cam, 0)
    T = scenes.se3_exp([0.05, -0.02, 0.03, 0.02, -0.03, 0.01]).astype(np.float32)
    gpu = CubeHandler(cam, 0.02, max_cubes=4096)
    ov = oracleapi.OracleVolume(cam, 0.02)
    mx, mn = gpu.ComputeBounding(d, T)
    omx, omn = ov.bounding(d, T)
    assert_bit_equal(mx, omx, "bbox max")
    assert_bit_equal(mn, omn, "bbox min")
    assert gpu.NumCubes() == 0        # the reference's ComputeBounding allocates nothing (CubeHandler.cpp:116-145)
    gpu.IntegrateImage(d, c, T)
    n = gpu.NumCubes()
    gpu.ComputeBounding(d, np.eye(4))
    assert gpu.NumCubes() == n


def test_errors_are_loud():
    cam = small_camera()
    d, c = scenes.wavy_wall(cam, 0)
    gpu = CubeHandler(cam, 0.02, max_cubes=64)
    # asynchronous frames cannot be re-run once their inputs are gone: a full pool is an error at the next synchronisation
    import torch
    pd, pc = torch.from_numpy(d).pin_memory(), torch.from_numpy(c).pin_memory()
    I = np.ascontiguousarray(np.eye(4, dtype=np.float32)).reshape(16)
    gpu.IntegrateImageAsync(pd.numpy(), capi.OPB_DEPTH_F32, pc.numpy(), I)
    with pytest.raises(capi.OpbError) as e:
        gpu.Synchronize()
    assert e.value.code == capi.OPB_ERR_CAPACITY
    gpu.Synchronize()                                 # reported once; the volume stays usable
    assert gpu.NumCubes() == 64
    gpu.IntegrateImage(d, c, np.eye(4))               # ... and the synchronous call completes it
    big = CubeHandler(cam, 0.02, max_cubes=4096)
    big.IntegrateImage(d, c, np.eye(4))
    assert gpu.NumCubes() == big.NumCubes()
    # unknown depth type: the reference exit(1)s (ImageProcessing.cpp:86-90); the C-ABI returns an error
    with pytest.raises(capi.OpbError) as e:
        gpu.IntegrateImage(d.astype(np.float64), c, np.eye(4))
    assert e.value.code == capi.OPB_ERR_INVALID
    h = C.c_void_p()
    bad = capi.VolumeDesc()
    capi.lib.opb_volume_desc_default(C.byref(bad))
    bad.voxel_resolution = 0.0
    assert capi.lib.opb_volume_create(C.byref(bad), C.byref(h)) == capi.OPB_ERR_INVALID


def test_async_pipeline_equals_sync():
    cam = small_camera()
    I = np.ascontiguousarray(np.eye(4, dtype=np.float32)).reshape(16)
    frames = [scenes.wavy_wall(cam, k) for k in range(8)]
    a = CubeHandler(cam, 0.02, max_cubes=4096)
    for d, c in frames:
        a.IntegrateImage(d, c, np.eye(4))
    b = CubeHandler(cam, 0.02, max_cubes=4096)
    # pinned ring of 3 host frames, enqueue-only calls
    ring = []
    for _ in range(3):
        pd, pc = C.c_void_p(), C.c_void_p()
        capi.check(capi.lib.opb_host_alloc(C.byref(pd), cam.width * cam.height * 4))
        capi.check(capi.lib.opb_host_alloc(C.byref(pc), cam.width * cam.height * 3))
        ring.append((pd, pc))
    for k, (d, c) in enumerate(frames):
        pd, pc = ring[k % 3]
        if k >= 3:
            b.Synchronize()  # simplest safe reuse rule for the test
        C.memmove(pd, d.ctypes.data, d.nbytes)
        C.memmove(pc, c.ctypes.data, c.nbytes)
        b.IntegrateImageAsync(pd.value, capi.OPB_DEPTH_F32, pc.value, I)
    b.Synchronize()
    ai, av = a.GetCubeMap()
    bi, bv = b.GetCubeMap()
    assert np.array_equal(ai, bi)
    assert_bit_equal(av, bv, "async vs sync")
    for pd, pc in ring:
        capi.lib.opb_host_free(pd)
        capi.lib.opb_host_free(pc)


def test_sharded_volumes_partition_the_cubes():
    """Sub-volume ownership: W shards together hold exactly the unsharded volume."""
    cam = small_camera()
    I = np.eye(4, dtype=np.float32)
    full = CubeHandler(cam, 0.02, max_cubes=4096)
    W = 4
    shards = [CubeHandler(cam, 0.02, max_cubes=4096, shard=(r, W, 0, 2)) for r in range(W)]
    for k in range(3):
        d, c = scenes.wavy_wall(cam, k)
        full.IntegrateImage(d, c, I)
        for s in shards:
            s.IntegrateImage(d, c, I)
    fi, fv = full.GetCubeMap()
    parts = [s.GetCubeMap() for s in shards]
    assert sum(len(p[0]) for p in parts) == len(fi)
    assert min(len(p[0]) for p in parts) > 0
    for r, (pi, _) in enumerate(parts):
        assert np.all((pi[:, 0] // 2) % W == r)
    ids = np.concatenate([p[0] for p in parts])
    vox = np.concatenate([p[1] for p in parts])
    order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))
    assert np.array_equal(ids[order], fi)
    assert_bit_equal(vox[order], fv, "sharded union")


def test_long_sequence_weights_up_to_120():
    """Exercises the shared-reciprocal quotient of the blend for every integer weight W = 2..120 on real data."""
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 8, c0.fy / 8, c0.cx / 8, c0.cy / 8, 80, 60, 1000.0)
    gpu, ov = CubeHandler(cam, 0.04, max_cubes=4096), oracleapi.OracleVolume(cam, 0.04)
    rng = np.random.default_rng(5)
    for k in range(120):
        d, c = scenes.wavy_wall(cam, k)
        c = rng.integers(0, 256, c.shape, dtype=np.uint8)
        T = scenes.se3_exp(rng.normal(0, 0.01, 6)).astype(np.float32)
        gpu.IntegrateImage(d, c, T)
        ov.integrate(d, c, T)
    compare_volumes(gpu, ov)
    assert gpu.GetCubeMap()[1][..., 1].max() >= 100


def test_wild_depth_values_take_the_ieee_division_path():
    """Depth values far outside any sensor's range (1e-9 m, 1e9 m) make the frame 'wild': the update kernel must
    switch to IEEE division and still match the oracle bit for bit.  (+inf depth is excluded: it back-projects to
    NaN points, and the reference's running std::max/std::min then depends on the raster position of that pixel.)"""
    cam = small_camera()
    gpu, ov = CubeHandler(cam, 0.02, max_cubes=8192), oracleapi.OracleVolume(cam, 0.02)
    I = np.eye(4, dtype=np.float32)
    d, c = scenes.wavy_wall(cam, 0)
    gpu.IntegrateImage(d, c, I)
    ov.integrate(d, c, I)
    d2 = d.copy()
    d2[10:20, 10:20] = 1e-9
    d2[30:40, 50:60] = 1e9
    d2[60:70, 60:90] = 3e-7
    gpu.IntegrateImage(d2, c, I)
    ov.integrate(d2, c, I)
    compare_volumes(gpu, ov)


def test_uploaded_wild_values_take_the_ieee_division_path():
    cam = small_camera()
    gpu, ov = CubeHandler(cam, 0.02, max_cubes=4096), oracleapi.OracleVolume(cam, 0.02)
    I = np.eye(4, dtype=np.float32)
    d, c = scenes.wavy_wall(cam, 0)
    ov.integrate(d, c, I)
    ids, vox = ov.download()
    rng = np.random.default_rng(1)
    vox = vox.copy()
    n = vox.shape[0]
    vox[: n // 4, ::3, 0] = 1e-30          # denormal-range quotients
    vox[: n // 4, ::5, 1] = 2.5            # non-integer weights
    vox[n // 4: n // 2, ::7, 2:] = 1e25    # huge colours
    vox[n // 2:, ::11, 1] = 3e7            # weights above 2^24
    vox[n // 2:, ::13, 0] = -1e-41         # denormal sdf
    gpu.SetCubeMap(ids, vox)
    ov.upload(ids, vox)
    for k in range(1, 4):
        d, c = scenes.wavy_wall(cam, k)
        gpu.IntegrateImage(d, c, I)
        ov.integrate(d, c, I)
    compare_volumes(gpu, ov)
    # Clear() makes the volume tame again
    gpu.Clear()
    ov.clear()
    gpu.IntegrateImage(d, c, I)
    ov.integrate(d, c, I)
    compare_volumes(gpu, ov)


def test_wild_pose_takes_the_ieee_division_path():
    cam = small_camera()
    gpu, ov = CubeHandler(cam, 0.02, max_cubes=8192), oracleapi.OracleVolume(cam, 0.02)
    d, c = scenes.wavy_wall(cam, 0)
    T = np.eye(4, dtype=np.float32)
    T[0, 1] = 1e-9  # shear far below the tame range
    T[2, 3] = 1e-8
    gpu.IntegrateImage(d, c, T)
    ov.integrate(d, c, T)
    compare_volumes(gpu, ov)
