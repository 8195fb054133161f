"""Ad-hoc GPU probe (not a test): CUDA volume vs the compiled reference on scene S1."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from onepiece_b200 import scenes
from onepiece_b200.volume import CubeHandler
from oracle import refapi

cam = scenes.Camera()
ref = refapi.RefVolume(cam, 0.005)
gpu = CubeHandler(cam, voxel_resolution=0.005, max_cubes=1 << 16)
I = np.eye(4, dtype=np.float32)
T = scenes.se3_exp([0.05, -0.02, 0.03, 0.02, -0.03, 0.01]).astype(np.float32)
for k, pose in enumerate([I, I, T, I]):
    d, c = scenes.wavy_wall(cam, k)
    mx, mn = ref.bounding(d, pose)
    ref.integrate(d, c, pose)
    gpu.IntegrateImage(d, c, pose)
    st = gpu.FrameStats()
    print("frame", k, "ref bbox", mn, mx)
    print("        gpu bbox", np.array(st.bbox_min[:]), np.array(st.bbox_max[:]), "cand", st.candidate_cubes,
          "frame cubes", st.frame_cubes, "total", st.total_cubes, "upd", st.updated_voxels, "ovf", st.overflow)
    print("        ref cubes", ref.num_cubes())
rid, rv = ref.download()
gid, gv = gpu.GetCubeMap()
print("ids equal:", rid.shape, gid.shape, np.array_equal(rid, gid))
if np.array_equal(rid, gid):
    same = rv.view(np.uint32) == gv.view(np.uint32)
    print("bit-exact voxels:", same.all(), "mismatching floats:", (~same).sum(), "of", same.size)
    if not same.all():
        bad = np.argwhere(~same)
        for b in bad[:10]:
            print(b, rv[tuple(b)], gv[tuple(b)])
        print("max abs diff", np.nanmax(np.abs(rv - gv)))
