#!/bin/bash
# Round-2 GPU pass P (8 GPUs): config-4 slab sweep at N=8
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/sweep_config4.py 8 4 2 1 \
    > gpurun_out/r02p_sweep_n8.log 2> gpurun_out/r02p_sweep_n8.err; echo "sweep exit $?"
grep "^{" gpurun_out/r02p_sweep_n8.log; grep "OpbError" gpurun_out/r02p_sweep_n8.err | head -3
