"""Device-resident point clouds (opb_cloud): PointCloud::LoadFromDepth on the device against the host restatement of the same
arithmetic, ICP on borrowed device buffers against the host-buffer call, integration of the frame a cloud was loaded from."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from onepiece_b200 import registration as reg
from onepiece_b200 import scenes
from onepiece_b200.volume import CubeHandler

pytestmark = pytest.mark.gpu


def small_camera():
    c = scenes.Camera()
    return scenes.Camera(c.fx / 2, c.fy / 2, c.cx / 2, c.cy / 2, 320, 240, 1000.0)


@pytest.mark.parametrize("u16", [True, False])
def test_load_from_depth_is_bit_identical_and_ordered(u16):
    cam = small_camera()
    d, _, _ = scenes.room(cam, 0)
    d = d.copy()
    rng = np.random.default_rng(3)
    d[rng.random(d.shape) < 0.05] = 0          # dropped pixels: the cloud is compacted in raster order (PointCloud.cpp:84-96)
    if not u16:
        d = (d.astype(np.float32) / np.float32(1000.0)).astype(np.float32)
        d[5, 7] = -1.0                          # z > 0 only
    cloud = reg.DeviceCloud().LoadFromDepth(d, cam)
    host = scenes.backproject(d, cam)
    assert cloud.size == len(host) < d.size
    assert_bit_equal(cloud.Download(), host, "LoadFromDepth")
    # an empty image
    z = np.zeros_like(d)
    assert reg.DeviceCloud().LoadFromDepth(z, cam).size == 0


@pytest.mark.parametrize("mode", ["plane", "point"])
def test_icp_on_device_clouds_equals_the_host_buffer_call(mode):
    cam = small_camera()
    d0, b0, _, n0 = scenes.room(cam, 0, with_normals=True)
    d1, b1, _ = scenes.room(cam, 2)
    tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)
    nrm = np.ascontiguousarray(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)])
    par = reg.ICPParameter(12, 0.05, 1.0)
    ct = reg.DeviceCloud().LoadFromDepth(d0, cam, b0)
    cs = reg.DeviceCloud().LoadFromDepth(d1, cam, b1)
    if mode == "plane":
        ct.SetNormals(nrm)
        a = reg.PointToPlaneClouds(cs, ct, np.eye(4), par)
        b = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), par)
    else:
        a = reg.PointToPointClouds(cs, ct, np.eye(4), par)
        b = reg.PointToPoint(reg.PointCloud(src), reg.PointCloud(tgt), np.eye(4), par)
    assert np.array_equal(a.correspondence_set_index, b.correspondence_set_index)
    assert np.array_equal(a.T.view(np.uint32), b.T.view(np.uint32)) and np.array_equal(a.T_iterated.view(np.uint32), b.T_iterated.view(np.uint32))
    assert a.rmse == b.rmse
    # the clouds are untouched and can be used again (source of one registration, target of the next)
    assert_bit_equal(cs.Download(), src, "source after the call")
    # PointToPoint with scaling works on scaled copies
    if mode == "point":
        a2 = reg.PointToPointClouds(cs, ct, np.eye(4), reg.ICPParameter(3, 0.1, 2.0))
        b2 = reg.PointToPoint(reg.PointCloud(src), reg.PointCloud(tgt), np.eye(4), reg.ICPParameter(3, 0.1, 2.0))
        assert np.array_equal(a2.T.view(np.uint32), b2.T.view(np.uint32))
        assert_bit_equal(cs.Download(), src, "source after a scaled call")
    # point-to-plane without normals: the reference's error path
    if mode == "plane":
        bare = reg.DeviceCloud().LoadFromDepth(d0, cam)
        assert not reg.PointToPlaneClouds(cs, bare, np.eye(4), par).ok


def test_integrate_the_frame_a_cloud_was_loaded_from():
    cam = small_camera()
    T = scenes.se3_exp([0.05, -0.02, 0.03, 0.02, -0.03, 0.01]).astype(np.float32)
    a, b = CubeHandler(cam, 0.02, max_cubes=1 << 15), CubeHandler(cam, 0.02, max_cubes=64)
    for k, pose in enumerate([np.eye(4, dtype=np.float32), T]):
        d, c, _ = scenes.room(cam, k)
        a.IntegrateImage(d, c, pose)
        b.IntegrateCloud(reg.DeviceCloud().LoadFromDepth(d, cam, c), pose)    # the small pool grows under it
    ai, av = a.GetCubeMap()
    bi, bv = b.GetCubeMap()
    oa, ob = np.lexsort(ai.T[::-1]), np.lexsort(bi.T[::-1])
    assert np.array_equal(ai[oa], bi[ob])
    assert_bit_equal(av[oa], bv[ob], "voxels")
