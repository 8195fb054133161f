#!/bin/bash
# Round-2 GPU pass J (8 GPUs): focus tests incl. the 2-GPU worker, config-4 slab sweep at N=8, bench N=8
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_fusion_gpu.py tests/test_odometry_gpu.py tests/test_icp_gpu.py tests/test_reference_mains.py -m gpu -q -s > gpurun_out/r02j_pytest_focus.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest_focus.log )
grep -n "main:\|main,\|default loop form\|exact 8-way\|passed\|failed\|FAILED" gpurun_out/r02j_pytest_focus.log | tail -14
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/sweep_config4.py 8 4 2 \
    > gpurun_out/r02j_sweep_n8.log 2> gpurun_out/r02j_sweep_n8.err; echo "sweep exit $?"
grep "^{" gpurun_out/r02j_sweep_n8.log; tail -c 600 gpurun_out/r02j_sweep_n8.err
