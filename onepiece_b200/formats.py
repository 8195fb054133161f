"""On-disk formats of the reference that sit either side of the hot path; host-side byte shuffling,
no arithmetic.

  .cubes        CubeHandler::WriteToFile / ReadFromFile (reference src/Integration/CubeHandler.h:40-69,113-128) with
                VoxelCube::WriteToBuffer / ReadFromBuffer (src/Integration/VoxelCube.h:128-167): one float stream --
                [uint32 cube count, bit-cast] then per cube 3 id floats, a 6-float record (voxel index, sdf, weight, c0, c1, c2)
                for every voxel with |sdf| < 1 and weight != 0, and the sentinel -2.0
  trajectory    tool::ReadImageSequenceWithPose's trajectory.txt (src/Tool/IO.cpp:81-108): one line per frame holding the 16
                values of the row-major 4x4 camera-to-world matrix
The PLY writer lives in onepiece_b200/mesh.py (write_ply)."""
from __future__ import annotations

import numpy as np


def write_cubes(filename, ids, vox):
    ids = np.asarray(ids, np.int32).reshape(-1, 3)
    vox = np.asarray(vox, np.float32).reshape(-1, 512, 5)
    parts = [np.array([len(ids)], np.uint32).view(np.float32)]
    for cid, cube in zip(ids, vox):
        keep = (np.abs(cube[:, 0]) < 1) & (cube[:, 1] != 0)
        rec = np.concatenate([np.nonzero(keep)[0].astype(np.float32)[:, None], cube[keep]], 1)
        parts += [cid.astype(np.float32), rec.reshape(-1), np.array([-2.0], np.float32)]
    np.concatenate(parts).astype(np.float32).tofile(filename)


def read_cubes(filename):
    """-> (ids [n,3] int32, voxels [n,512,5] float32) or None if the file cannot be read.  Voxels without a record keep the
    TSDFVoxel defaults (999, 0, -1, -1, -1)."""
    try:
        buf = np.fromfile(filename, np.float32)
    except OSError:
        return None
    n = int(buf[:1].view(np.uint32)[0])
    ids = np.zeros((n, 3), np.int32)
    vox = np.zeros((n, 512, 5), np.float32)
    vox[:, :, 0], vox[:, :, 2:] = 999.0, -1.0
    ptr = 1
    for c in range(n):
        ids[c] = buf[ptr:ptr + 3].astype(np.int32)
        ptr += 3
        end = ptr
        while buf[end] != -2.0:   # records are 6 floats wide, so the sentinel is only looked for at record starts (as ReadFromBuffer does)
            end += 6
        rec = buf[ptr:end].reshape(-1, 6)
        vox[c, rec[:, 0].astype(np.int64)] = rec[:, 1:]
        ptr = end + 1
    return ids, vox


def write_trajectory(filename, poses):
    with open(filename, "w") as f:
        for T in poses:
            f.write(" ".join(repr(float(x)) for x in np.asarray(T, np.float64).reshape(16)) + "\n")


def read_trajectory(filename):
    """-> list of 4x4 float64 camera-to-world matrices (one per non-empty line, as the reference's getline loop reads them)"""
    poses = []
    for line in open(filename):
        vals = line.split()
        if len(vals) >= 16:
            poses.append(np.array(vals[:16], np.float64).reshape(4, 4))
    return poses
