"""Generates tests/golden/*.npz from the reference's own translation units (oracle/_ref, built by
`make -C oracle ref` from /root/reference).  Run in the build container:

    python tests/golden/gen_golden.py

Each fixture stores the INPUTS (so the tests do not depend on numpy's sin/cos implementation of the machine
they run on) and the reference's outputs: cube ids, a SHA-256 of the sorted voxel array, the first cubes raw,
Marching Cubes vertex counts and a SHA-256 of the sorted triangle list."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import canon_triangles, sha  # noqa: E402
from onepiece_b200 import scenes  # noqa: E402
from oracle import refapi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def small_camera():
    c = scenes.Camera()
    return scenes.Camera(c.fx / 4, c.fy / 4, c.cx / 4, c.cy / 4, 160, 120, 1000.0)


def integrate_fixture(name, cam, res, trunc, frames, depth_u16=False):
    ref = refapi.RefVolume(cam, res, trunc)
    depths, bgrs, poses, lists = [], [], [], []
    for k, pose in frames:
        d, c = scenes.wavy_wall(cam, k)
        if depth_u16:
            d = np.clip(np.rint(d * cam.depth_scale), 0, 65535).astype(np.uint16)
        ref.integrate(d, c, pose)
        depths.append(d); bgrs.append(c); poses.append(np.asarray(pose, np.float32))
    ids, vox = ref.download()
    _, pts, col, tri = ref.extract_mesh()
    canon = canon_triangles(pts, col)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        cam=np.array([cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.depth_scale], np.float64),
        res=np.float32(res), trunc=np.float32(trunc),
        depth=np.stack(depths), bgr=np.stack(bgrs), poses=np.stack(poses),
        ids=ids, voxels_sha=sha(vox), voxels_head=vox[:16],
        weight_sum=np.float64(vox[..., 1].sum(dtype=np.float64)),
        n_vertices=np.int64(len(pts)), n_triangles=np.int64(len(tri)), mesh_sha=sha(canon), mesh_head=canon[:64])
    print(name, "cubes", len(ids), "verts", len(pts), "file", os.path.getsize(os.path.join(OUT, name + ".npz")))


def main():
    cam = small_camera()
    I = np.eye(4, dtype=np.float32)
    T1 = scenes.se3_exp([0.05, -0.02, 0.03, 0.02, -0.03, 0.01]).astype(np.float32)
    T2 = scenes.se3_exp([-0.1, 0.04, 0.2, -0.05, 0.08, 0.3]).astype(np.float32)
    integrate_fixture("integrate_small_f32", cam, 0.02, 0.1, [(0, I), (1, T1), (2, I), (3, T2)])
    integrate_fixture("integrate_small_u16", cam, 0.02, 0.1, [(0, I), (1, T1), (2, T2)], depth_u16=True)
    integrate_fixture("integrate_small_trunc", cam, 0.04, 0.3, [(0, I), (1, T2)])


if __name__ == "__main__" and "--icp" not in sys.argv and "--odometry" not in sys.argv:
    main()


def icp_fixture():
    """S2 frame pair at 160x120: inputs + the float64 reference's result (truth) + the float32 reference's result
    (its own deviation from float64 is the noise floor pose parity is gated on)."""
    cam = small_camera()
    d0, _, _, n0 = scenes.room(cam, 0, with_normals=True)
    d1, _, _, _ = scenes.room(cam, 3, with_normals=True)
    tgt = scenes.backproject(d0, cam)
    src = scenes.backproject(d1, cam)
    nrm = np.ascontiguousarray(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)])
    out = dict(src=src, tgt=tgt, nrm=nrm, max_iter=np.int32(10), threshold=np.float64(0.05))
    for mode, normals in (("plane", nrm), ("point", None)):
        r64 = refapi.icp(src, tgt, normals, np.eye(4), 10, 0.05, "f64")
        r32 = refapi.icp(src, tgt, normals, np.eye(4), 10, 0.05, "f32")
        out[f"{mode}_T64"] = r64["T"]
        out[f"{mode}_T32"] = r32["T"]
        out[f"{mode}_rmse64"] = np.float64(r64["rmse"])
        out[f"{mode}_pairs64"] = r64["pairs"]
        out[f"{mode}_pairs32_equal"] = np.bool_(np.array_equal(r64["pairs"], r32["pairs"]))
    np.savez_compressed(os.path.join(OUT, "icp_small.npz"), **out)
    print("icp_small", os.path.getsize(os.path.join(OUT, "icp_small.npz")))


if __name__ == "__main__" and "--icp" in sys.argv:
    icp_fixture()


def cv2_frame_images(bgr, depth_u16, depth_scale):
    """The reference's per-frame pre-processing (Odometry.cpp:609-620,436-449) run through REAL OpenCV (python cv2, the
    only OpenCV in this container; the README pins 3.4) -> img[what][level]"""
    import cv2
    gray = cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY).astype(np.float32) / np.float32(255.0)
    d = depth_u16.astype(np.float32)
    ok = (d > 0.5 * depth_scale) & (d < 4 * depth_scale)
    d32 = np.where(ok, d / np.float32(depth_scale), np.float32(np.nan)).astype(np.float32)
    base = [cv2.GaussianBlur(gray, (3, 3), 0), cv2.GaussianBlur(d32, (3, 3), 0)]
    img = [[None] * 3 for _ in range(6)]
    for a in range(2):
        img[a][0] = base[a]
        for l in (1, 2):
            img[a][l] = cv2.pyrDown(img[a][l - 1])
        for l in range(3):
            img[2 + 2 * a][l] = cv2.Sobel(img[a][l], cv2.CV_32F, 1, 0)
            img[3 + 2 * a][l] = cv2.Sobel(img[a][l], cv2.CV_32F, 0, 1)
    return img


def odometry_fixture():
    """S2 frame pair at 160x120 for the dense RGB-D odometry path.  Stores the raw inputs, the pre-processed images as
    real OpenCV produces them (tolerance check of the restated filters), and the compiled reference's outputs on the
    ORACLE-built images (integer-exact C, identical on every machine): teacher-forced single iterations at every
    level / term (float32 build: correspondence lists; float64 build: next pose = truth for the solve) and the
    whole coarse-to-fine run of both builds."""
    from oracle import oracleapi
    cam = small_camera()
    d0, c0, _ = scenes.room(cam, 0)
    d1, c1, _ = scenes.room(cam, 3)
    S = oracleapi.OracleFrame(c1, d1).preprocess(cam.depth_scale)
    T = oracleapi.OracleFrame(c0, d0).preprocess(cam.depth_scale)
    si, ti = S.images(), T.images()
    out = dict(cam=np.array([cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.depth_scale], np.float64),
               src_bgr=c1, src_depth=d1, tgt_bgr=c0, tgt_depth=d0)
    cvi = cv2_frame_images(c0, d0, cam.depth_scale)
    for a in range(6):
        for l in range(3):
            out[f"cv2_tgt_{a}_{l}"] = cvi[a][l]
            out[f"sha_src_{a}_{l}"] = sha(si[a][l])
            out[f"sha_tgt_{a}_{l}"] = sha(ti[a][l])
    rng = np.random.default_rng(7)
    cases = []
    for level in (2, 1, 0):
        for term in (0, 1, 2):
            T0 = scenes.se3_exp(rng.normal(0, 3e-3, 6)).astype(np.float32)
            r32 = refapi.single_iteration(si, ti, cam, level, T0, term, "f32")
            r64 = refapi.single_iteration(si, ti, cam, level, T0, term, "f64")
            k = len(cases)
            cases.append((level, term))
            out[f"tf{k}_T0"] = T0
            out[f"tf{k}_pairs32"] = r32["pairs"].astype(np.uint16)
            out[f"tf{k}_pairs64_equal"] = np.bool_(np.array_equal(r32["pairs"], r64["pairs"]))
            out[f"tf{k}_T32"], out[f"tf{k}_T64"] = r32["T"], r64["T"]
            out[f"tf{k}_JTJ64"], out[f"tf{k}_JTr64"], out[f"tf{k}_r2_64"] = r64["JTJ"], r64["JTr"], np.float64(r64["r2"])
    out["tf_cases"] = np.array(cases, np.int32)
    # the whole coarse-to-fine run (Odometry.cpp:526-608) from the identity.  The RGBDFrame overload normalises the level-0 gray
    # in place first: let the oracle do that (NormalizeIntensity is checked against the reference separately below)
    n0 = refapi.correspondences(si[1][0], ti[1][0], cam, np.eye(4), "f32")
    sg, tg = refapi.normalize_intensity(si[0][0], ti[0][0], n0, "f32")
    out["norm_pairs"] = np.int64(len(n0))
    out["norm_src_sha"], out["norm_tgt_sha"] = sha(sg), sha(tg)
    si[0][0], ti[0][0] = sg, tg
    for term in (0, 1, 2):
        for kind in ("f32", "f64"):
            r = refapi.multiscale(si, ti, cam, np.eye(4), term, kind)
            out[f"ms{term}_T_{kind}"] = r["T"]
            out[f"ms{term}_rmse_{kind}"] = np.float64(r["rmse"])
            out[f"ms{term}_ok_{kind}"] = np.bool_(r["success"])
            out[f"ms{term}_corr_{kind}"] = r["corr_per_iteration"]
            out[f"ms{term}_Tit_{kind}"] = r["T_per_iteration"]
            out[f"ms{term}_npairs_{kind}"] = np.int64(len(r["pairs"]))
    np.savez_compressed(os.path.join(OUT, "odometry_small.npz"), **out)
    print("odometry_small", os.path.getsize(os.path.join(OUT, "odometry_small.npz")))


if __name__ == "__main__" and "--odometry" in sys.argv:
    odometry_fixture()
