"""One full-size point-to-plane ICP call (for ncu launch lists)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onepiece_b200 import registration as reg, scenes  # noqa: E402

cam = scenes.Camera()
d0, _, _, n0 = scenes.room(cam, 0, with_normals=True)
d1, _, _ = scenes.room(cam, 1)
tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)
nrm = np.ascontiguousarray(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    r = reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), reg.ICPParameter(30, 0.05, 1.0))
print(r.T_iterated)
