"""Full-size timings of the kernels either side of the hot path: depth pre-filter, volume resampling /
merging, mesh post-processing -- device path vs the CPU implementation available on the box (cv2 for the OpenCV filter,
oracle/_ref = the compiled reference for the rest).  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from onepiece_b200 import imageproc, scenes  # noqa: E402
from onepiece_b200.mesh import TriangleMesh  # noqa: E402
from onepiece_b200.volume import CubeHandler  # noqa: E402
from oracle import oracleapi, refapi  # noqa: E402


def best(fn, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


out = {}
cam = scenes.Camera()
d, c, T = scenes.room(cam, 0)
pf = imageproc.DepthPrefilter(cam.width, cam.height)
pf.run(d, cam.depth_scale)
out["prefilter_ms_host_to_host"] = best(lambda: pf.run(d, cam.depth_scale))
try:
    import cv2
    f32 = oracleapi.convert_depth_32f(d, cam.depth_scale)
    cv2.setNumThreads(0)
    out["prefilter_ms_cv2_all_threads"] = best(lambda: cv2.bilateralFilter(f32, 7, 0.03, 4.5))
    cv2.setNumThreads(1)
    out["prefilter_ms_cv2_one_thread"] = best(lambda: cv2.bilateralFilter(f32, 7, 0.03, 4.5))
except ImportError:
    pass
out["prefilter_ms_oracle"] = best(lambda: oracleapi.bilateral_filter(oracleapi.convert_depth_32f(d, cam.depth_scale)), 2)

gv = CubeHandler(cam, 0.005, max_cubes=1 << 17)
rv = refapi.RefVolume(cam, 0.005) if refapi.available("f32") else None
for k in range(2):
    d, c, T = scenes.room(cam, 4 * k)
    gv.IntegrateImage(d, c, T)
    if rv is not None:
        rv.integrate(d, c, T)
out["volume_cubes"] = gv.NumCubes()
Tm = scenes.se3_exp([0.03, -0.02, 0.05, 0.1, -0.2, 0.15]).astype(np.float32)
out["transform_nearest_ms"] = best(lambda: gv.TransformNearest(Tm, result_resolution=0.0).close(), 3)
out["transform_trilinear_ms"] = best(lambda: gv.Transform(Tm).close(), 3)
other = gv.Transform(Tm)
dst = CubeHandler(cam, 0.005, max_cubes=1 << 17)
dst.Merge(gv)
out["merge_ms"] = best(lambda: dst.Merge(other), 3)
if rv is not None:
    t0 = time.perf_counter(); rt = rv.transform(Tm, False); out["transform_trilinear_ms_reference_cpu"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); rv.merge(rt); out["merge_ms_reference_cpu"] = (time.perf_counter() - t0) * 1e3

pts, col, tri = gv.ExtractTriangleMesh()
out["mesh_triangles"] = len(tri)
mesh = TriangleMesh(pts, col, tri)
mesh.ClusteringSimplify(0.005)
out["clustering_ms_host_to_host"] = best(lambda: mesh.ClusteringSimplify(0.005), 3)
out["extract_mesh_ms"] = best(lambda: gv.ExtractTriangleMesh(), 3)
out["extract_clustered_fused_ms"] = best(lambda: gv.ExtractTriangleMeshClustered(0.005), 3)
m2 = gv.ExtractTriangleMeshClustered(0.005)
out["clustered_vertices"] = len(m2.points)
out["normals_ms_host_to_host"] = best(lambda: m2.ComputeNormals(), 3)
if rv is not None:
    out["clustering_ms_reference_cpu"] = refapi.clustering_simplify(pts, col, tri, 0.005)[4] * 1e3
    t0 = time.perf_counter(); refapi.compute_normals(m2.points, m2.triangles); out["normals_ms_reference_cpu"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(out))
