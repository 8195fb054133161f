#!/bin/bash
# A/B of the seeded grid walk (OPB_ICP_SEED): ICP tests with the seed on, then the bench step with it off and on
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py tests/test_headline_gpu.py -m gpu -q -x > gpurun_out/r02seed_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02seed_pytest.log | cut -c1-300
for seed in 0 1; do
  OPB_ICP_SEED=$seed timeout 600 python bench.py --steps 100 --warmup 10 --no-odometry --no-cpu-baseline > gpurun_out/r02seed_bench_$seed.json 2> gpurun_out/r02seed_bench_$seed.err
  python - <<PY
import json
b = json.load(open("gpurun_out/r02seed_bench_$seed.json"))
print("seed $seed value", round(b["value"], 1), "ms", round(b["ms_per_step"], 4), "e2e", round(b["e2e"]["value"], 1), b["details"]["step_breakdown_ms"], "parity", b.get("parity_check", {}).get("ok"))
PY
done
