"""The reference's caller code (example/ImageSequenceIntegration.cpp / example/ICPTest.cpp style) compiled
unchanged against the drop-in C++ classes in onepiece_b200/cpp, run on the GPU and compared with the oracle.
The binary is built in the build container (tests/cpp/Makefile needs the reference's headers) and travels to the
GPU box with the snapshot."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal
from onepiece_b200 import scenes
from oracle import oracleapi

BIN = os.path.join(ROOT, "tests", "cpp", "dropin_main.bin")


def test_dropin_sources_cite_and_mirror_the_reference_api():
    hdr = open(os.path.join(ROOT, "onepiece_b200", "cpp", "Integration", "CubeHandler.h")).read()
    for name in ("SetVoxelResolution", "SetTruncation", "SetCamera", "SetFarPlane", "SetNearPlane", "IntegrateImage",
                 "ExtractTriangleMesh", "GetCubeMap", "SetCubeMap", "HasCube", "Clear", "GetPointCloud", "WriteToFile",
                 "ReadFromFile", "PrepareCubes", "ComputeBounding"):
        assert name in hdr, name
    assert "namespace integration" in hdr and "class CubeHandler" in hdr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="tests/cpp/dropin_main.bin not built (needs the reference headers)")
def test_reference_caller_code_runs_on_the_gpu(tmp_path):
    c0 = scenes.Camera()
    cam = scenes.Camera(c0.fx / 4, c0.fy / 4, c0.cx / 4, c0.cy / 4, 160, 120, 1000.0)
    ov = oracleapi.OracleVolume(cam, 0.02)
    poses = []
    n_frames = 3
    for k in range(n_frames):
        d, c = scenes.wavy_wall(cam, k)
        T = scenes.se3_exp(np.array([0.02, -0.01, 0.01, 0.01, -0.02, 0.005]) * k).astype(np.float32)
        d.tofile(tmp_path / f"depth{k}.bin")
        c.tofile(tmp_path / f"bgr{k}.bin")
        poses.append(T)
        ov.integrate(d, c, T)
    np.stack(poses).astype(np.float32).tofile(tmp_path / "poses.bin")
    d0, _, _, n0 = scenes.room(cam, 0, with_normals=True)
    d1, _, _ = scenes.room(cam, 3)
    tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)
    nrm = np.ascontiguousarray(n0.reshape(-1, 3))
    src.tofile(tmp_path / "src.bin"); tgt.tofile(tmp_path / "tgt.bin"); nrm.tofile(tmp_path / "nrm.bin")
    # example/DenseOdometry.cpp inputs: two S2 frames (uint16 depth, like a sensor)
    _, c0_bgr, _ = scenes.room(cam, 0)
    _, c1_bgr, _ = scenes.room(cam, 3)
    d0.tofile(tmp_path / "odo_depth0.bin"); c0_bgr.tofile(tmp_path / "odo_bgr0.bin")
    d1.tofile(tmp_path / "odo_depth1.bin"); c1_bgr.tofile(tmp_path / "odo_bgr1.bin")
    out = subprocess.run([BIN, str(tmp_path), str(n_frames)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "dropin ok" in out.stdout
    # the reference prints one line per IntegrateImage call; the drop-in keeps that behaviour
    assert out.stdout.count("Finish image integration") == n_frames + 1  # + the pre-filtered frame at the end
    ids = np.fromfile(tmp_path / "ids.bin", np.int32).reshape(-1, 3)
    vox = np.fromfile(tmp_path / "voxels.bin", np.float32).reshape(-1, 512, 5)
    order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))
    oi, ovx = ov.download()
    assert np.array_equal(ids[order], oi)
    assert_bit_equal(vox[order], ovx, "voxels through the C++ drop-in")
    mesh = np.fromfile(tmp_path / "mesh_points.bin", np.float32).reshape(-1, 3)
    op, _ = ov.extract_mesh()
    assert len(mesh) == len(op)
    icp = np.fromfile(tmp_path / "icp.bin", np.float64)
    o = oracleapi.icp(src, tgt, nrm, np.eye(4), 10, 0.05)
    assert np.linalg.norm(icp[:16].reshape(4, 4)[:3, 3] - o["T"][:3, 3]) < 1e-5
    assert int(icp[17]) == len(o["pairs"])
    # PointCloud::EstimateNormals through the drop-in (example/ICPTest.cpp:27-29)
    en = np.fromfile(tmp_path / "estimated_normals.bin", np.float32).reshape(-1, 3)
    assert_bit_equal(en, oracleapi.estimate_normals(tgt), "EstimateNormals through the drop-in")
    # registration::ComputeFPFHFeature through the drop-in, called the way DenseSlam's submap back end calls it
    fp = np.fromfile(tmp_path / "fpfh_points.bin", np.float32).reshape(-1, 3)
    ff = np.fromfile(tmp_path / "fpfh.bin", np.float32).reshape(-1, 33)
    down = oracleapi.downsample(tgt, None, None, 0.05)[0]
    assert_bit_equal(fp, down, "DownSample through the drop-in")
    of = oracleapi.fpfh(down, oracleapi.estimate_normals(down, 0.1, 30), 100, 0.25)
    assert len(ff) == len(of) > 100 and ((ff.view(np.uint32) == of.view(np.uint32)) | (np.isnan(ff) & np.isnan(of))).all()
    # the rest of RansacRegistration: matching + three rejection passes through the drop-ins are the oracle's index for index;
    # the estimator (drop-in too) draws fresh samples on every call like the reference's GRANSAC, so its pose is checked against
    # the true camera motion, and its inlier list against the oracle's evaluation of the pose it returned
    sdown = oracleapi.downsample(src, None, None, 0.05)[0]
    assert_bit_equal(np.fromfile(tmp_path / "fpfh_source_points.bin", np.float32).reshape(-1, 3), sdown, "source DownSample")
    sf = oracleapi.fpfh(sdown, oracleapi.estimate_normals(sdown, 0.1, 30), 100, 0.25)
    m0 = oracleapi.feature_matching(sf, of)
    assert np.array_equal(np.fromfile(tmp_path / "matches_initial.bin", np.int32).reshape(-1, 2), m0)
    m3 = oracleapi.reject_matches(sdown, down, m0, 3)
    assert np.array_equal(np.fromfile(tmp_path / "matches_kept.bin", np.int32).reshape(-1, 2), m3) and 8 < len(m3) < len(m0)
    rr = np.fromfile(tmp_path / "ransac.bin", np.float64)
    T_est, n_inl, n_corr = rr[:16].reshape(4, 4), int(rr[16]), int(rr[17])
    T_true = np.linalg.inv(scenes.room_pose(0)) @ scenes.room_pose(3)
    assert n_corr == len(m3) and n_inl > 0.3 * n_corr
    ids = rr[18:18 + n_inl].astype(np.int64)
    a3, b3 = sdown[m3[:, 0]], down[m3[:, 1]]
    err = np.linalg.norm(a3 @ T_est[:3, :3].T + T_est[:3, 3] - b3, axis=1)
    assert (np.diff(ids) > 0).all() and (err[ids] < 0.1 + 1e-5).all() and (np.delete(err, ids) > 0.1 - 1e-5).all()
    assert np.linalg.norm(T_est[:3, 3] - T_true[:3, 3]) < 0.1 and np.abs(T_est[:3, :3] - T_true[:3, :3]).max() < 0.1
    # Odometry::DenseTracking through the drop-in: two chained calls on the same RGBDFrames, then the cv::Mat overload
    odo = np.fromfile(tmp_path / "odometry.bin", np.float64)
    S, T = oracleapi.OracleFrame(c1_bgr, d1), oracleapi.OracleFrame(c0_bgr, d0)
    for call in range(2):
        o = oracleapi.dense_tracking_frames(S, T, cam, np.eye(4), 0)
        rec = odo[call * 21:(call + 1) * 21]
        assert np.abs(rec[:16].reshape(4, 4) - o["T"]).max() < 1e-6, f"DenseTracking call {call}"
        assert abs(rec[16] - o["rmse"]) < 1e-7 and int(rec[17]) == len(o["pairs"]) and bool(rec[18]) == o["success"]
        assert int(rec[19]) == len(o["pairs"]) and rec[20] == 1.0
    om = oracleapi.dense_tracking(c1_bgr, c0_bgr, d1, d0, cam, np.eye(4), 0)
    assert np.abs(odo[42:58].reshape(4, 4) - om["T"]).max() < 1e-6 and int(odo[59]) == len(om["pairs"])
    # TransformNearest (+ its forgotten c_para), Transform and Merge through the drop-in
    rs = np.fromfile(tmp_path / "resample.bin", np.float64)
    mid = poses[n_frames // 2]
    on = ov.transform(mid, True)
    ot = ov.transform(mid, False)
    om = oracleapi.OracleVolume(cam, 0.02)
    om.merge(ot)
    om.merge(ov)
    assert int(rs[0]) == on.num_cubes() and int(rs[1]) == len(on.extract_mesh()[0])
    assert int(rs[2]) == ot.num_cubes() and int(rs[3]) == om.num_cubes()
    assert rs[4] == float(om.download()[1][:, :, 1].astype(np.float64).sum())
    # ClusteringSimplify -> ComputeNormals -> DownSample through the drop-in definitions (onepiece_b200/cpp/Geometry/MeshPost.cpp)
    post = np.fromfile(tmp_path / "meshpost.bin", np.float32)
    # (vertex clustering depends on the triangle order, so the oracle is fed the mesh in the order the drop-in extracted it)
    op = mesh
    oc = np.fromfile(tmp_path / "mesh_colors.bin", np.float32).reshape(-1, 3)
    cp, cc, ct = oracleapi.clustering_simplify(op, oc, np.arange(len(op), dtype=np.uint32).reshape(-1, 3), 0.02)
    cn = oracleapi.compute_normals(cp, ct)
    dp, dc, dn = oracleapi.downsample(cp, cc, cn, 0.05)
    assert (int(post[0]), int(post[1]), int(post[2])) == (len(cp), len(ct), len(dp))
    body = post[3:3 + 6 * len(cp)].reshape(-1, 3, 2)
    assert_bit_equal(body[:, :, 0], cp, "clustered points through the drop-in")
    assert_bit_equal(body[:, :, 1], cn, "normals through the drop-in")
    rest = post[3 + 6 * len(cp):].reshape(-1, 3, 3)
    assert_bit_equal(rest[:, :, 0], dp, "down-sampled points through the drop-in")
    assert_bit_equal(rest[:, :, 1], dn, "down-sampled normals")
    assert_bit_equal(rest[:, :, 2], dc, "down-sampled colours")
    # tool::ConvertDepthTo32F + tool::BilateralFilter through the drop-in, then IntegrateImage of the filtered depth
    oc = oracleapi.convert_depth_32f(d1, cam.depth_scale)
    of = oracleapi.bilateral_filter(oc)
    assert_bit_equal(np.fromfile(tmp_path / "refined_depth.bin", np.float32).reshape(d1.shape), oc, "ConvertDepthTo32F drop-in")
    assert_bit_equal(np.fromfile(tmp_path / "filtered_depth.bin", np.float32).reshape(d1.shape), of, "BilateralFilter drop-in")
    ovf = oracleapi.OracleVolume(cam, 0.02)
    ovf.integrate(of, c1_bgr, np.eye(4, dtype=np.float32))
    assert int(np.fromfile(tmp_path / "filtered_cubes.bin", np.float64)[0]) == ovf.num_cubes()
    # the .cubes file the drop-in wrote has the reference's layout: [u32 n_cubes] then per cube 3 id floats ... -2
    raw = np.fromfile(tmp_path / "volume.cubes", np.float32)
    assert raw[:1].view(np.uint32)[0] == len(oi) and (raw == -2.0).sum() >= len(oi)
