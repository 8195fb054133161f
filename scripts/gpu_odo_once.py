"""A few chained DenseTracking calls at 640x480 (for ncu launch lists)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onepiece_b200 import scenes  # noqa: E402
from onepiece_b200.odometry import Odometry  # noqa: E402

cam = scenes.Camera()
odo = Odometry(cam)
frames = [scenes.room(cam, k)[:2] for k in range(4)]
prev = odo.Frame(frames[0][1], frames[0][0])
for k in range(1, 4):
    cur = odo.Frame(frames[k][1], frames[k][0])
    r = odo.DenseTracking(cur, prev, np.eye(4), 0, want_correspondences=False)
    prev = cur
import ctypes as C
from onepiece_b200 import capi
ph = (C.c_uint64 * 4)()
capi.check(capi.lib.opb_odometry_last_phases(odo.handle, ph))
print("phases us (28 iterations): candidates %.1f  barrier %.1f  reduce %.1f  release-wait %.1f" % tuple(x / 1e3 for x in ph))
print(r.T)
