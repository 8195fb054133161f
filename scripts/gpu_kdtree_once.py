"""One full-frame EstimateNormals, one DenseSlam-style FPFH extraction and one RansacRegistration against a second view (for ncu
launch lists and wall-clock timing)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onepiece_b200 import registration as reg, scenes  # noqa: E402

cam = scenes.Camera()
d0, _, _ = scenes.room(cam, 0)
cloud = scenes.backproject(d0, cam)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    pc = reg.PointCloud(cloud)
    t0 = time.perf_counter()
    pc.EstimateNormals()                       # PointCloud.cpp:102-144 with the defaults (0.1, 30)
    t1 = time.perf_counter()
    down = pc.DownSample(0.05)                 # DenseSlam.h:49-56: voxel 0.05, normals (0.1, 30), FPFH (100, 0.25)
    t2 = time.perf_counter()
    down.EstimateNormals(0.1, 30)
    t3 = time.perf_counter()
    f = reg.ComputeFPFHFeature(down, 100, 0.25)
    t4 = time.perf_counter()
# the rest of the submap registration chain (GlobalRegistration.cpp:219-267) against a second view
d1, _, _ = scenes.room(cam, 12)
pc1 = reg.PointCloud(scenes.backproject(d1, cam)).DownSample(0.05)
pc1.EstimateNormals(0.1, 30)
f1 = reg.ComputeFPFHFeature(pc1, 100, 0.25)
for _ in range(reps):
    t5 = time.perf_counter()
    m = reg.FeatureMatching3D(f, f1)
    t6 = time.perf_counter()
    res = reg.RansacRegistration(down, pc1, f, f1, reg.RANSACParameter(max_iteration=40000, threshold=0.1), seed=1)
    t7 = time.perf_counter()
print(f"FeatureMatching3D {len(f)} x {len(f1)} {1e3 * (t6 - t5):.2f} ms | RansacRegistration (matching + 3 rejections + 40,000 hypotheses) "
      f"{1e3 * (t7 - t6):.2f} ms, {len(res.correspondence_set_index)} inliers")
print(f"EstimateNormals {len(cloud)} pts {1e3 * (t1 - t0):.2f} ms | DownSample -> {len(down.points)} pts {1e3 * (t2 - t1):.2f} ms | "
      f"EstimateNormals {1e3 * (t3 - t2):.2f} ms | ComputeFPFHFeature {1e3 * (t4 - t3):.2f} ms | finite rows {int(np.isfinite(f).all(1).sum())}")
