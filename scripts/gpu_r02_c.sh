#!/bin/bash
# Round-2 GPU pass C: parity tests, ICP phases of the reworked loop, the rewritten bench (both arms)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest_gpu.log )
tail -8 gpurun_out/r02c_pytest_gpu.log
timeout 300 python scripts/gpu_icp_phases.py gpurun_out/r02c_icp_phases.json > gpurun_out/r02c_icp_phases.log 2>&1; echo "phases exit $?"
head -8 gpurun_out/r02c_icp_phases.log; tail -4 gpurun_out/r02c_icp_phases.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02c_bench.err
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02c_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], "pose_only", b["e2e"]["pose_only"]["value"], "refsig", b["e2e"]["reference_signature"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"])
    print("parity", b.get("parity_check"))
    print("config4@1", {k: v for k, v in b.get("partitioned_fusion", {}).items() if k not in ("what", "e2e_note")})
    print("odo", b.get("dense_odometry", {}).get("value"), b.get("dense_odometry", {}).get("e2e"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/r02c_bench_ref.json
