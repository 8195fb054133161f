"""OPB_STORAGE_PACKED16 (8-byte voxels: half sdf, half weight, rgb8), the throughput mode north_star sketches as "fp16 voxel
writes".  It is outside the parity contract by construction (half precision is coarser than the 1e-4 relative gate), so it carries
its own gate: the deviation from the float volume -- which is bit-exact with the reference (tests/test_volume_gpu.py) -- on
BASELINE.json config 1 (50 frames of the S1 sequence at 640x480, 5 mm voxels, identity poses)."""
import numpy as np
import pytest

from onepiece_b200 import capi, scenes
from onepiece_b200.volume import CubeHandler

pytestmark = pytest.mark.gpu

# the gate of the packed mode, against the float path
MAX_SDF_ERR_M = 5e-4        # |sdf_packed - sdf_float| over every observed voxel: half rounding (<= 3e-5 m for a value below 0.125 m) of a running average, re-rounded every frame (measured 2.9e-4 m after 29 frames of the bench stream)
MAX_COLOR_ERR = 4.0 / 255   # colour channels are stored as bytes and re-rounded every frame (measured 2.6 / 255 after 29 frames)
MAX_VERTEX_DELTA = 2e-3     # relative difference of the Marching-Cubes vertex count


def test_config1_50_frames_against_the_float_volume():
    cam = scenes.Camera()
    f32 = CubeHandler(cam, 0.005, max_cubes=1 << 16)
    p16 = CubeHandler(cam, 0.005, max_cubes=1 << 16, storage=capi.OPB_STORAGE_PACKED16)
    I = np.eye(4, dtype=np.float32)
    for k in range(50):
        d, c = scenes.wavy_wall(cam, k)
        f32.IntegrateImage(d, c, I)
        p16.IntegrateImage(d, c, I)
        assert p16.FrameStats().updated_voxels == f32.FrameStats().updated_voxels  # same projection, same band test
    fi, fv = f32.GetCubeMap()
    pi, pv = p16.GetCubeMap()
    assert np.array_equal(fi, pi), "the two storage modes select the same cubes"
    seen = fv[..., 1] > 0
    assert np.array_equal(seen, pv[..., 1] > 0)
    assert np.array_equal(fv[..., 1], pv[..., 1]), "weights up to 2048 are exact in half precision"
    d_sdf = np.abs(fv[..., 0] - pv[..., 0])[seen].max()
    d_col = np.abs(fv[..., 2:] - pv[..., 2:])[seen].max()
    # never-observed voxels keep the TSDFVoxel defaults in both modes
    assert np.array_equal(pv[~seen], fv[~seen])
    nv_f, nv_p = f32.CountMesh()[0], p16.CountMesh()[0]
    rel = abs(nv_f - nv_p) / nv_f
    print(f"packed16 vs f32 after 50 frames: max |d sdf| {d_sdf:.3e} m, max |d colour| {d_col:.4f}, "
          f"Marching Cubes vertices {nv_p} vs {nv_f} (relative delta {rel:.2e}), {len(fi)} cubes")
    assert d_sdf < MAX_SDF_ERR_M
    assert d_col < MAX_COLOR_ERR
    assert rel < MAX_VERTEX_DELTA
    pts, col, tri = p16.ExtractTriangleMesh()
    assert len(pts) == nv_p and np.isfinite(pts).all() and col.min() >= 0 and col.max() <= 1


def test_packed_pool_grows_and_refuses_what_it_cannot_do():
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0)
    I = np.eye(4, dtype=np.float32)
    d, c = scenes.wavy_wall(cam, 0)
    big = CubeHandler(cam, 0.01, max_cubes=1 << 14, storage=capi.OPB_STORAGE_PACKED16)
    small = CubeHandler(cam, 0.01, max_cubes=64, storage=capi.OPB_STORAGE_PACKED16)  # grows under the synchronous call
    for v in (big, small):
        v.IntegrateImage(d, c, I)
        v.IntegrateImage(d, c, I)
    bi, bv = big.GetCubeMap()
    si, sv = small.GetCubeMap()
    ob, os_ = np.lexsort((bi[:, 2], bi[:, 1], bi[:, 0])), np.lexsort((si[:, 2], si[:, 1], si[:, 0]))
    assert np.array_equal(bi[ob], si[os_]) and np.array_equal(bv[ob], sv[os_])
    assert big.CountMesh() == small.CountMesh()
    big.Clear()
    assert big.NumCubes() == 0
    with pytest.raises(capi.OpbError) as e:
        big.SetCubeMap(si, sv)
    assert e.value.code == capi.OPB_ERR_UNSUPPORTED
    with pytest.raises(capi.OpbError) as e:
        CubeHandler(cam, 0.01, storage=capi.OPB_STORAGE_PACKED16, shard=(0, 2, 0, 4))
    assert e.value.code == capi.OPB_ERR_UNSUPPORTED
    with pytest.raises(capi.OpbError) as e:
        small.Transform(np.eye(4))
    assert e.value.code == capi.OPB_ERR_UNSUPPORTED
