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
if os.environ.get("OPB_ODO_PERSISTENT", "1") == "1":
    print("phases us (%d iterations): candidates+acceptance+rows %.1f  (unused %.1f)  publish+barrier+sum %.1f  solve %.1f" % ((r.iterations,) + tuple(x / 1e3 for x in ph)))
else:
    print("phases us (%d iterations): candidates %.1f  barrier %.1f  reduce %.1f  release-wait %.1f" % ((r.iterations,) + tuple(x / 1e3 for x in ph)))
ms, tail = C.c_float(0), C.c_float(0)
capi.check(capi.lib.opb_odometry_set_profiling(odo.handle, 1))
cur = odo.Frame(frames[1][1], frames[1][0])
tgt = odo.Frame(frames[0][1], frames[0][0])
for _ in range(3):
    r = odo.DenseTracking(cur, tgt, np.eye(4), 0, want_correspondences=False)
    capi.check(capi.lib.opb_odometry_last_timing(odo.handle, C.byref(ms), C.byref(tail)))
    print("tracking call %.3f ms on the device" % ms.value)
print(r.T)
