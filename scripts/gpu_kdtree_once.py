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
# Optimizer::FastBA over a synthetic 24-pose graph (25 frame pairs x 5,000 point pairs), five iterations
from onepiece_b200 import optimization as opt  # noqa: E402
rng = np.random.default_rng(24)
true = [np.eye(4, dtype=np.float32)] + [scenes.se3_exp(rng.normal(0, 0.15, 6)).astype(np.float32) for _ in range(23)]
links = [(i, i + 1) for i in range(23)] + [(0, 23), (1, 22)]
cs = []
for s_, t_ in links:
    w = rng.uniform(-1, 1, (5000, 3)) + [0, 0, 3]
    inv_s, inv_t = np.linalg.inv(true[s_].astype(np.float64)), np.linalg.inv(true[t_].astype(np.float64))
    cs.append(opt.Correspondence(s_, t_, (w @ inv_s[:3, :3].T + inv_s[:3, 3]).astype(np.float32), (w @ inv_t[:3, :3].T + inv_t[:3, 3]).astype(np.float32)))
start = np.stack([true[0]] + [(scenes.se3_exp(rng.normal(0, 0.03, 6)) @ T).astype(np.float32) for T in true[1:]])
for _ in range(reps):
    t8 = time.perf_counter()
    refined = opt.Optimizer().FastBA(cs, start, 5)
    t9 = time.perf_counter()
print(f"FastBA 24 poses, 25 x 5000 point pairs, 5 iterations {1e3 * (t9 - t8):.2f} ms | max pose error {np.abs(refined - np.stack(true)).max():.2e}")
print(f"EstimateNormals {len(cloud)} pts {1e3 * (t1 - t0):.2f} ms | DownSample -> {len(down.points)} pts {1e3 * (t2 - t1):.2f} ms | "
      f"EstimateNormals {1e3 * (t3 - t2):.2f} ms | ComputeFPFHFeature {1e3 * (t4 - t3):.2f} ms | finite rows {int(np.isfinite(f).all(1).sum())}")
