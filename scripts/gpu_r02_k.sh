#!/bin/bash
# Round-2 GPU pass K (2 GPUs): fusion / odometry / icp tests, config-4 slab sweep at N=2
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_fusion_gpu.py tests/test_odometry_gpu.py -m gpu -q -s > gpurun_out/r02k_pytest_focus.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest_focus.log )
grep -n "default loop form\|passed\|failed\|FAILED\|Error" gpurun_out/r02k_pytest_focus.log | cut -c1-300 | tail -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 scripts/sweep_config4.py 8 2 \
    > gpurun_out/r02k_sweep_n2.log 2> gpurun_out/r02k_sweep_n2.err; echo "sweep exit $?"
grep "^{" gpurun_out/r02k_sweep_n2.log; grep "OpbError" gpurun_out/r02k_sweep_n2.err | head -4
