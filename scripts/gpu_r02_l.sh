#!/bin/bash
# Round-2 GPU pass L (1 GPU): all GPU tests, ICP phases, bench
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_pytest_gpu.log )
tail -8 gpurun_out/r02l_pytest_gpu.log | cut -c1-300
timeout 300 python scripts/gpu_icp_phases.py gpurun_out/r02l_icp_phases.json > gpurun_out/r02l_icp_phases.log 2>&1; echo "phases exit $?"
head -8 gpurun_out/r02l_icp_phases.log; tail -3 gpurun_out/r02l_icp_phases.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02l_bench.err
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02l_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"])
    o = b.get("dense_odometry", {})
    print("odometry", o.get("value"), o.get("device_ms_per_frame"))
    print("parity", b.get("parity_check", {}).get("ok"), b.get("parity_check", {}).get("dt_m_vs_float64_reference"))
    print("config4 N=1", {k: v for k, v in b.get("partitioned_fusion", {}).items() if k in ("frames_per_s", "e2e_frames_per_s", "error")})
except Exception as e:
    print("bench parse failed", e)
PY
