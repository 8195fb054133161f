#!/bin/bash
# Round-2 closing GPU pass (1 GPU): every GPU test, smoke, the bench line and the reference arm
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02ver_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ver_pytest_gpu.log )
tail -4 gpurun_out/r02ver_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ver_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02ver_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02ver_bench.json 2> gpurun_out/r02ver_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02ver_bench_ref.json 2> gpurun_out/r02ver_bench_ref.err; echo "bench ref exit $?"
python - <<'PY'
import json
b = json.load(open("gpurun_out/r02ver_bench.json"))
print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], "pairs inside", b["e2e"]["pairs_inside_the_call"]["value"])
print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"], "clocks", b["clocks"])
o = b.get("dense_odometry", {})
print("odometry", o.get("value"), o.get("device_ms_per_frame"))
print("parity", b.get("parity_check", {}).get("ok"))
r = json.load(open("gpurun_out/r02ver_bench_ref.json"))
print("reference arm", r["value"], r["steps"], r["cpu_baseline"]["cores"], "same config", r["config"] == b["config"])
PY
