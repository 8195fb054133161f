"""Full-size timing of PointCloud::EstimateNormals: device vs the compiled reference (nanoflann + Eigen on one core)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from onepiece_b200 import registration as reg, scenes
from oracle import refapi
cam = scenes.Camera()
d, _, _ = scenes.room(cam, 0)
pts = scenes.backproject(d, cam)
pc = reg.PointCloud(pts)
pc.EstimateNormals()
t0 = time.perf_counter(); pc.EstimateNormals(); t1 = time.perf_counter()
print(f"device: {1e3 * (t1 - t0):.1f} ms for {len(pts)} points (host to host)")
if refapi.available("f32"):
    n, dt = refapi.estimate_normals(pts)
    agree = np.abs((n * pc.normals).sum(1))
    print(f"reference CPU: {1e3 * dt:.0f} ms; bit-identical normals {np.mean((n.view(np.uint32) == pc.normals.view(np.uint32)).all(1)):.3f}, |dot| > 0.999: {np.mean(agree > 0.999):.4f}")
