"""Dense-odometry micro-benchmark (frames resident on the device) for profiling: one DenseTracking per step."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from onepiece_b200 import scenes
from onepiece_b200.odometry import Odometry
cam = scenes.Camera()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
odo = Odometry(cam)
odo.set_profiling(True)
imgs = [scenes.room(cam, k) for k in range(4)]
for want in (False, True):
    ts, ds = [], []
    for k in range(n):
        a, b = imgs[k % 3], imgs[k % 3 + 1]
        S, T = odo.Frame(b[1], b[0]).preprocess(), odo.Frame(a[1], a[0]).preprocess()
        t0 = time.perf_counter()
        r = odo.DenseTracking(S, T, np.eye(4), 0, want_correspondences=want)
        ts.append((time.perf_counter() - t0) * 1e3)
        ds.append(odo.last_tracking_ms())
    print(f"pairs downloaded={want}: call median {np.median(ts):.3f} ms, device {np.median(ds):.3f} ms, iterations {r.iterations}, "
          f"correspondences {r.corr_per_iteration[-1]}, success {r.tracking_success}, solve tail {odo.last_solve_tail_us:.2f} us/iteration")
# pre-processing alone
t0 = time.perf_counter()
for k in range(20):
    odo.Frame(imgs[0][1], imgs[0][0]).preprocess().image(0, 2)
print(f"frame upload + preprocess + 1 small download: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
# per-level cost of one solver iteration (candidates + iteration kernels + launch gaps): difference of two schedules
if "--levels" in sys.argv:
    for lvl in (2, 1, 0):
        res = []
        for reps in (8, 24):
            it = [0, 0, 0]
            it[lvl] = reps
            o2 = Odometry(cam, levels=3, iterations=tuple(it))
            o2.set_profiling(True)
            S, T = o2.Frame(imgs[1][1], imgs[1][0]).preprocess(), o2.Frame(imgs[0][1], imgs[0][0]).preprocess()
            d = []
            for k in range(6):
                o2.DenseTracking(S, T, np.eye(4), 0, want_correspondences=False)
                d.append(o2.last_tracking_ms())
            res.append(np.median(d))
        print(f"level {lvl}: {(res[1] - res[0]) / 16 * 1e3:.1f} us per iteration (8 its {res[0]:.3f} ms, 24 its {res[1]:.3f} ms)")
