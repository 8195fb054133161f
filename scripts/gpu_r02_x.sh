#!/bin/bash
# Round-2 GPU pass X (1 GPU): two-level candidate pruning, pass = 4 super-blocks: volume tests, A/B bench
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_volume_gpu.py tests/test_fusion_gpu.py tests/test_headline_gpu.py tests/test_dropin_cpp.py -m gpu -q > gpurun_out/r02x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02x_pytest.log )
tail -4 gpurun_out/r02x_pytest.log | cut -c1-300
for pr in 0 1; do
OPB_SELECT_PRUNE=$pr timeout 600 python bench.py --steps 20 --warmup 5 --no-odometry --no-cpu-baseline > gpurun_out/r02x_bench_prune$pr.json 2> gpurun_out/r02x_bench_prune$pr.err
python - $pr <<'PY'
import json, sys
try:
    b = json.load(open(f"gpurun_out/r02x_bench_prune{sys.argv[1]}.json"))
    pf = b.get("partitioned_fusion", {})
    print("OPB_SELECT_PRUNE=" + sys.argv[1], "value", round(b["value"], 1), "e2e", round(b["e2e"]["value"], 1), "select ms", b["details"]["step_breakdown_ms"]["cube_selection"],
          "cubes/frame", b["details"]["cubes_per_frame"], "| config4 N=1:", pf.get("frames_per_s"), pf.get("rank0", {}).get("select_ms"), pf.get("cubes_total"), pf.get("mesh_vertices_total"))
except Exception as e:
    print("bench parse failed", e)
PY
done
