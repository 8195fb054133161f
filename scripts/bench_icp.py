"""ICP micro-benchmark (device-resident clouds) for profiling."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from onepiece_b200 import scenes, capi, registration as reg
cam = scenes.Camera()
d0, _, T0, n0 = scenes.room(cam, 0, with_normals=True)
d1, _, T1, _ = scenes.room(cam, 1, with_normals=True)
tgt = torch.from_numpy(scenes.backproject(d0, cam)).cuda(); src = torch.from_numpy(scenes.backproject(d1, cam)).cuda()
nrm = torch.from_numpy(n0.reshape(-1, 3)[(d0 > 0).reshape(-1)].copy()).cuda()
ws = reg._Workspace.get(0)
capi.lib.opb_icp_set_profiling(ws, 1)
par = capi.IcpParams(30, 0.05, 1.0); res = capi.IcpResult()
I = np.ascontiguousarray(np.eye(4, dtype=np.float32)).reshape(16)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for k in range(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    capi.check(capi.lib.opb_icp_point_to_plane(ws, C.c_void_p(src.data_ptr()), len(src), C.c_void_p(tgt.data_ptr()), C.c_void_p(nrm.data_ptr()), len(tgt), I.ctypes.data_as(C.c_void_p), C.byref(par), C.byref(res), None, 0))
    dt = time.perf_counter() - t0
    a, b = C.c_float(0), C.c_float(0); capi.lib.opb_icp_last_timing(ws, C.byref(a), C.byref(b))
    print("call %.3f ms  grid %.3f ms  loop %.3f ms  inliers %d rmse %.6g" % (dt * 1e3, a.value, b.value, res.n_inliers, res.rmse))
