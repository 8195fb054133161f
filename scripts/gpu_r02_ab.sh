#!/bin/bash
# Round-2 GPU pass AB (1 GPU): ICP search with prefetched row ranges -- ICP tests, phases, bench (headline only)
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_icp_gpu.py tests/test_headline_gpu.py tests/test_fusion_gpu.py tests/test_cloud_gpu.py -m gpu -q > gpurun_out/r02ab_pytest_icp.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ab_pytest_icp.log )
tail -6 gpurun_out/r02ab_pytest_icp.log | cut -c1-300
timeout 300 python scripts/gpu_icp_phases.py gpurun_out/r02ab_icp_phases.json > gpurun_out/r02ab_icp_phases.log 2>&1; echo "phases exit $?"
head -8 gpurun_out/r02ab_icp_phases.log; tail -3 gpurun_out/r02ab_icp_phases.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-odometry --no-partitioned > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02ab_bench.err
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02ab_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"])
    print("parity", b.get("parity_check", {}).get("ok"), b.get("parity_check", {}).get("dt_m_vs_float64_reference"))
except Exception as e:
    print("bench parse failed", e)
PY
