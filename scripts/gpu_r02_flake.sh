#!/bin/bash
# repeats the shared-device split-ICP test (it timed out once in ~8 runs before the NULL-stream copies left the call path)
mkdir -p gpurun_out
pass=0; fail=0
for i in $(seq 1 12); do
  if timeout 300 python -m pytest tests/test_fusion_gpu.py -m gpu -q -k "split_icp_workspaces" > gpurun_out/r02flake_$i.log 2>&1; then pass=$((pass+1)); else fail=$((fail+1)); tail -5 gpurun_out/r02flake_$i.log | cut -c1-300; fi
done
echo "split-ICP shared-device test: $pass passed, $fail failed of 12 runs"
