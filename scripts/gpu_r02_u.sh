#!/bin/bash
# Round-2 GPU pass U (1 GPU): ICP tests (async pair list), bench
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_icp_gpu.py tests/test_headline_gpu.py tests/test_cloud_gpu.py -m gpu -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02u_pytest.log )
tail -4 gpurun_out/r02u_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02u_bench.err
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02u_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], "pairs inside", b["e2e"]["pairs_inside_the_call"]["value"], "pose only", b["e2e"]["pose_only"]["value"], "refsig", b["e2e"]["reference_signature"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"], "icp traffic", b["roofline_icp"]["traffic"])
    o = b.get("dense_odometry", {})
    print("odometry", o.get("value"), o.get("device_ms_per_frame"), o.get("e2e_pose_only", {}).get("value"))
    print("parity", b.get("parity_check", {}).get("ok"), b.get("parity_check", {}).get("dt_m_vs_float64_reference"))
    print("config4 N=1", {k: v for k, v in b.get("partitioned_fusion", {}).items() if k in ("frames_per_s", "e2e_frames_per_s", "error")})
except Exception as e:
    print("bench parse failed", e)
PY
