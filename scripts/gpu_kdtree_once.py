"""One full-frame EstimateNormals and one DenseSlam-style FPFH extraction (for ncu launch lists and wall-clock timing)."""
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
print(f"EstimateNormals {len(cloud)} pts {1e3 * (t1 - t0):.2f} ms | DownSample -> {len(down.points)} pts {1e3 * (t2 - t1):.2f} ms | "
      f"EstimateNormals {1e3 * (t3 - t2):.2f} ms | ComputeFPFHFeature {1e3 * (t4 - t3):.2f} ms | finite rows {int(np.isfinite(f).all(1).sum())}")
