#!/bin/bash
# Round-2 final GPU pass (1 GPU): all GPU tests, smoke, bench + reference arm, ncu launch list and full captures of the hot kernels
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02fin_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02fin_pytest_gpu.log )
tail -5 gpurun_out/r02fin_pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fin_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02fin_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02fin_bench.json 2> gpurun_out/r02fin_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02fin_bench_ref.json 2> gpurun_out/r02fin_bench_ref.err; echo "bench ref exit $?"
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02fin_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], "pairs inside", b["e2e"]["pairs_inside_the_call"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"], "clocks", b["clocks"])
    o = b.get("dense_odometry", {})
    print("odometry", o.get("value"), o.get("device_ms_per_frame"))
    print("parity", b.get("parity_check", {}).get("ok"))
    r = json.load(open("gpurun_out/r02fin_bench_ref.json"))
    print("reference arm", r["value"], r["steps"], r["cpu_baseline"]["cores"], "same config", r["config"] == b["config"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02fin_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-odometry --no-partitioned --no-cpu-baseline > gpurun_out/r02fin_ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:odo_loop2_kernel -c 1 -f -o gpurun_out/r02fin_full_odo_loop2_kernel \
    python scripts/gpu_odo_once.py > gpurun_out/r02fin_ncu_odo.log 2>&1; echo "ncu odo exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:icp_loop2_kernel -c 1 -f -o gpurun_out/r02fin_full_icp_loop2_kernel \
    python scripts/gpu_icp_once.py > gpurun_out/r02fin_ncu_icp.log 2>&1; echo "ncu icp exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:integrate_packed_kernel -s 10 -c 1 -f -o gpurun_out/r02fin_full_integrate_packed_kernel \
    python bench.py --steps 3 --warmup 3 --no-odometry --no-cpu-baseline > gpurun_out/r02fin_ncu_packed.log 2>&1; echo "ncu packed exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:integrate_pipelined_kernel -s 12 -c 1 -f -o gpurun_out/r02fin_full_integrate_pipelined_kernel \
    python bench.py --steps 3 --warmup 3 --no-odometry --no-partitioned --no-cpu-baseline > gpurun_out/r02fin_ncu_pipe.log 2>&1; echo "ncu pipelined exit $?"
ls -la gpurun_out/r02fin*.ncu-rep
