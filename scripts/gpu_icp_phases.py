"""Per-pass phase times of the persistent ICP loop kernel on the bench workload (640x480 S2 pair, 30 iterations): where a pass
spends its time.  Prints a table and writes profiles/<name>.json when given a path."""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from onepiece_b200 import capi, registration as reg, scenes  # noqa: E402

cam = scenes.Camera()
d0, _, _, n0 = scenes.room(cam, 0, with_normals=True)
d1, _, _ = scenes.room(cam, 1)
tgt, src = scenes.backproject(d0, cam), scenes.backproject(d1, cam)
nrm = np.ascontiguousarray(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)])
par = reg.ICPParameter(30, 0.05, 1.0)
ws = reg._Workspace.get(0)
capi.lib.opb_icp_set_profiling(ws, 1)
rows = []
for rep in range(6):
    reg.PointToPlane(reg.PointCloud(src), reg.PointCloud(tgt, nrm), np.eye(4), par, want_pairs=False)
    st = np.zeros((48, 8), np.uint64)
    capi.check(capi.lib.opb_icp_last_stamps(ws, st.ctypes.data_as(C.c_void_p), 48))
    a, b = C.c_float(0), C.c_float(0)
    capi.lib.opb_icp_last_timing(ws, C.byref(a), C.byref(b))
    rows.append((st.astype(np.int64), a.value, b.value))
st, grid_ms, loop_ms = rows[-1]
n_pass = 31
us = lambda x: x / 1e3
out = {"grid_build_ms": grid_ms, "loop_ms": loop_ms, "passes": []}
print(f"grid build {grid_ms:.3f} ms, loop + finaliser {loop_ms:.3f} ms")
print("pass  total   search  accum+publish  wait_for_release | last CTA: arrive->last_known  sum  solve  release_seen")
for p in range(n_pass):
    s = st[p]
    nxt = st[p + 1][0] if p + 1 < n_pass else s[2]
    row = {"pass": p, "total_us": us(nxt - s[0]), "search_us": us(s[1] - s[0]), "accumulate_publish_us": us(s[2] - s[1]),
           "wait_release_us": us(s[3] - s[2]) if p + 1 < n_pass else 0.0,
           "cta0_done_to_last_known_us": us(s[4] - s[2]), "sum_us": us(s[5] - s[4]), "solve_us": us(s[6] - s[5]),
           "solve_to_release_seen_us": us(s[3] - s[6]) if p + 1 < n_pass else 0.0}
    out["passes"].append(row)
    print(f"{p:3d} {row['total_us']:7.1f} {row['search_us']:7.1f} {row['accumulate_publish_us']:10.1f} {row['wait_release_us']:12.1f}"
          f"   | {row['cta0_done_to_last_known_us']:10.1f} {row['sum_us']:6.1f} {row['solve_us']:6.1f} {row['solve_to_release_seen_us']:8.1f}")
steady = out["passes"][3:30]
for k in ("total_us", "search_us", "accumulate_publish_us", "wait_release_us", "cta0_done_to_last_known_us", "sum_us", "solve_us",
          "solve_to_release_seen_us"):
    out["steady_" + k] = float(np.mean([r[k] for r in steady]))
print("steady-state pass (3..29):", {k: round(v, 2) for k, v in out.items() if k.startswith("steady_")})
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
