#!/bin/bash
# Round-2 GPU final pass Z (1 GPU): all GPU tests, smoke, bench, reference arm
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02z_pytest_gpu.log )
tail -6 gpurun_out/r02z_pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02z_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench exit $?"
tail -c 400 gpurun_out/r02z_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err; echo "bench ref exit $?"
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02z_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], "pairs inside", b["e2e"]["pairs_inside_the_call"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"], "clocks", b["clocks"])
    o = b.get("dense_odometry", {})
    print("odometry", o.get("value"), o.get("device_ms_per_frame"))
    print("parity", b.get("parity_check", {}).get("ok"))
    print("config4 N=1", {k: v for k, v in b.get("partitioned_fusion", {}).items() if k in ("frames_per_s", "e2e_frames_per_s", "error")})
    print("packed", b.get("packed16_voxels", {}).get("kernel_speedup"), b.get("packed16_voxels", {}).get("deviation_from_f32"))
    r = json.load(open("gpurun_out/r02z_bench_ref.json"))
    print("reference arm", r["value"], r["steps"], r["cpu_baseline"]["cores"], "same config", r["config"] == b["config"])
except Exception as e:
    print("bench parse failed", e)
PY
