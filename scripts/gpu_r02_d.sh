#!/bin/bash
# Round-2 GPU pass D: parity tests, ICP phases after the search window / shorter load chain / fused finaliser, bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest_gpu.log )
tail -8 gpurun_out/r02d_pytest_gpu.log
timeout 300 python scripts/gpu_icp_phases.py gpurun_out/r02d_icp_phases.json > gpurun_out/r02d_icp_phases.log 2>&1; echo "phases exit $?"
head -8 gpurun_out/r02d_icp_phases.log; tail -4 gpurun_out/r02d_icp_phases.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-odometry --no-partitioned > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02d_bench.err
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02d_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], "pose_only", b["e2e"]["pose_only"]["value"], "refsig", b["e2e"]["reference_signature"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"])
    print("parity", b.get("parity_check"))
except Exception as e:
    print("bench parse failed", e)
PY
