#!/bin/bash
# Round-2 GPU pass W (1 GPU): two-level candidate pruning in select_kernel -- all GPU tests, A/B of the selection, bench
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02w_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02w_pytest_gpu.log )
tail -6 gpurun_out/r02w_pytest_gpu.log | cut -c1-300
for pr in 0 1; do
OPB_SELECT_PRUNE=$pr timeout 600 python bench.py --steps 20 --warmup 5 --no-odometry --no-cpu-baseline > gpurun_out/r02w_bench_prune$pr.json 2> gpurun_out/r02w_bench_prune$pr.err
python - $pr <<'PY'
import json, sys
try:
    b = json.load(open(f"gpurun_out/r02w_bench_prune{sys.argv[1]}.json"))
    pf = b.get("partitioned_fusion", {})
    print("OPB_SELECT_PRUNE=" + sys.argv[1], "value", round(b["value"], 1), "e2e", round(b["e2e"]["value"], 1), "select ms", b["details"]["step_breakdown_ms"]["cube_selection"],
          "cubes/frame", b["details"]["cubes_per_frame"], "| config4 N=1:", pf.get("frames_per_s"), pf.get("rank0", {}).get("select_ms"), pf.get("cubes_total"), pf.get("mesh_vertices_total"))
except Exception as e:
    print("bench parse failed", e)
PY
done
