"""Diagnostic: how many queries each ICP pass has to search when nearest-neighbour certificates are on."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onepiece_b200 import capi, registration as reg, scenes  # noqa: E402

cam = scenes.Camera()
d0, _, _, n0 = scenes.room(cam, 0, with_normals=True)
d1, _, _ = scenes.room(cam, 1)
tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)
nrm = np.ascontiguousarray(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)])
prev = np.eye(4)
steps = []
for it in (1, 2, 3, 4, 6, 8, 12, 16, 20, 30):
    r = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), reg.ICPParameter(it, 0.05, 1.0))
    steps.append((it, r.T_iterated.copy()))
p = np.array([0.5, 0.4, 3.0, 1.0])
last = np.eye(4)
for it, T in steps:
    print(f"after {it:2d} iterations: point moved {1e3 * np.linalg.norm((T - last) @ p):.4f} mm since the previous probe")
    last = T
tr = np.zeros(64, np.uint32)
capi.check(capi.lib.opb_icp_last_search_trace(reg._Workspace.get(0), tr.ctypes.data_as(C.c_void_p), 64))
print("guard", os.environ.get("OPB_ICP_GUARD", "default"), "searched per pass (of", len(src), "):", tr[:31].tolist())
