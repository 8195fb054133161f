"""Parity on the very workload bench.py's headline is quoted on (BASELINE.json config 2): a 640x480 S2 frame pair, 307,200
points each, registration::PointToPlane with 30 iterations and threshold 0.05, then CubeHandler::IntegrateImage at 5 mm with
the ICP pose.  CUDA path through the C-ABI against the compiled reference (oracle/_ref, float64 and float32 builds: it travels
with the snapshot); the volume against the oracle.  (The oracle's own ICP searches by brute force: 9e10 distances per iteration at
this size, so the ICP leg is checked against the compiled reference only.)"""
import numpy as np
import pytest

from conftest import assert_bit_equal
from onepiece_b200 import registration as reg
from onepiece_b200 import scenes
from onepiece_b200.volume import CubeHandler
from oracle import oracleapi, refapi

pytestmark = pytest.mark.gpu

ITERS, THR, VOXEL = 30, 0.05, 0.005


def pose_delta(A, B):
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    R = A[:3, :3].T @ B[:3, :3]
    ang = np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return float(np.linalg.norm(A[:3, 3] - B[:3, 3])), float(ang)


@pytest.fixture(scope="module")
def stream():
    """frames 0..3 of the bench stream (rank 0), exactly as bench.py's make_stream builds them"""
    cam = scenes.Camera()
    frames = []
    for k in range(4):
        d, c, T, n = scenes.room(cam, k, with_normals=True)
        frames.append(dict(depth=d, bgr=c, pose=T.astype(np.float64), cloud=scenes.backproject(d, cam),
                           normals=np.ascontiguousarray(n.reshape(-1, 3)[(d > 0).reshape(-1)])))
    return cam, frames


@pytest.fixture(scope="module")
def registrations(stream):
    cam, frames = stream
    out = []
    for k in range(3):
        a, b = frames[k], frames[k + 1]
        assert len(b["cloud"]) == 307200 and len(a["cloud"]) == 307200
        out.append(reg.PointToPlane(reg.PointCloud(b["cloud"]), reg.PointCloud(a["cloud"], a["normals"]), np.eye(4),
                                    reg.ICPParameter(ITERS, THR, 1.0)))
    return out


def test_full_size_point_to_plane_vs_the_float64_reference(stream, registrations):
    """north_star tolerance: 1e-5 m / 1e-4 rad against the reference (its -DUSING_FLOAT64 build is the truth; the float32 build's
    own deviation is printed as the noise floor).  Inlier pairs must be the float64 reference's."""
    if not refapi.available("f64"):
        pytest.skip("oracle/_ref (the compiled reference) did not travel")
    cam, frames = stream
    a, b = frames[0], frames[1]
    g = registrations[0]
    r64 = refapi.icp(b["cloud"], a["cloud"], a["normals"], np.eye(4), ITERS, THR, "f64")
    r32 = refapi.icp(b["cloud"], a["cloud"], a["normals"], np.eye(4), ITERS, THR, "f32")
    dt, dr = pose_delta(g.T, r64["T"])
    ft, fr = pose_delta(r32["T"], r64["T"])
    print(f"\nfull-size PointToPlane vs float64 reference: CUDA {dt:.3e} m / {dr:.3e} rad; float32 reference itself {ft:.3e} m / {fr:.3e} rad; "
          f"inliers CUDA {len(g.correspondence_set_index)} ref64 {len(r64['pairs'])} ref32 {len(r32['pairs'])}; "
          f"rmse CUDA {g.rmse:.9g} ref64 {r64['rmse']:.9g} ref32 {r32['rmse']:.9g}")
    assert dt < 1e-5 and dr < 1e-4, (dt, dr)
    # Inlier pairs: the float64 reference's, except where its pose -- which differs from the float32 pose of this path in the
    # eighth digit -- flips an exactly borderline decision (two target points equally near, a pair at the inlier radius);
    # the float32 reference's own list differs from the float64 one in the same way.
    mine, ref = set(map(tuple, g.correspondence_set_index)), set(map(tuple, r64["pairs"]))
    ref32 = set(map(tuple, r32["pairs"]))
    n_diff, n_diff32 = len(mine ^ ref), len(ref32 ^ ref)
    print(f"pairs differing from the float64 reference: CUDA {n_diff}, float32 reference {n_diff32} of {len(ref)}")
    assert n_diff <= max(3, 2 * n_diff32), (n_diff, n_diff32)
    assert abs(g.rmse - r64["rmse"]) < 1e-7
    # the float32 reference differs from its float64 build by more than the CUDA path does, or by nothing at all
    assert dt <= ft + 1e-7


def test_three_frames_integrated_with_the_icp_poses_vs_the_oracle(stream, registrations):
    """bench steps 0..2: the CUDA volume fed with the CUDA ICP poses against the oracle volume fed with the same poses -- cube
    set and voxels bit-equal, Marching-Cubes vertex count identical"""
    cam, frames = stream
    gpu = CubeHandler(cam, VOXEL, max_cubes=1 << 17)
    ov = oracleapi.OracleVolume(cam, VOXEL)
    for k in range(3):
        pose = (frames[k]["pose"] @ registrations[k].T.astype(np.float64)).astype(np.float32)
        gpu.IntegrateImage(frames[k + 1]["depth"], frames[k + 1]["bgr"], pose)
        n = ov.integrate(frames[k + 1]["depth"], frames[k + 1]["bgr"], pose)
        st = gpu.FrameStats()
        assert st.frame_cubes == n and st.overflow == 0
    gi, gv = gpu.GetCubeMap()
    oi, ovx = ov.download()
    assert np.array_equal(gi, oi), f"cube sets differ: {len(gi)} vs {len(oi)}"
    assert_bit_equal(gv, ovx, "voxels")
    nv, nt = gpu.CountMesh()
    op, _ = ov.extract_mesh()
    assert nv == len(op) and nt * 3 == nv
