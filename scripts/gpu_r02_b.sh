#!/bin/bash
# Round-2 GPU pass B: parity tests, DMMA micro-benchmark, ICP per-pass phase times of the new loop kernel (and the old one), bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest_gpu.log )
tail -15 gpurun_out/r02b_pytest_gpu.log
timeout 120 scripts/micro/dmma_rate.bin > gpurun_out/r02b_dmma_rate.log 2>&1; echo "dmma exit $?"; cat gpurun_out/r02b_dmma_rate.log
timeout 300 python scripts/gpu_icp_phases.py gpurun_out/r02b_icp_phases.json > gpurun_out/r02b_icp_phases.log 2>&1; echo "phases exit $?"
head -8 gpurun_out/r02b_icp_phases.log; tail -6 gpurun_out/r02b_icp_phases.log
OPB_ICP_PERSISTENT=2 timeout 300 python scripts/gpu_icp_phases.py gpurun_out/r02b_icp_phases_old.json > gpurun_out/r02b_icp_phases_old.log 2>&1; echo "phases(old) exit $?"
tail -3 gpurun_out/r02b_icp_phases_old.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:icp_loop2_kernel -c 1 -f -o gpurun_out/r02b_full_icp_loop2_kernel \
    python scripts/gpu_icp_once.py > gpurun_out/r02b_ncu_icp.log 2>&1; echo "ncu icp exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-odometry > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02b_bench.json"))
    print({k: b[k] for k in ("value", "ms_per_step")}, b["e2e"]["value"], b["config"]["step_breakdown_ms"], b["roofline"]["frac"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r02b_bench.err").read()[-2000:])
PY
