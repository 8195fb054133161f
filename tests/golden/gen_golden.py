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


if __name__ == "__main__" and "--icp" not in sys.argv:
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
