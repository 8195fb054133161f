#!/bin/bash
# one `ncu --set full` capture of the three k-d tree kernels that carry the time
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:kd_normals_kernel -c 1 -f -o gpurun_out/kd_normals_kernel \
    python scripts/gpu_kdtree_once.py 1 > gpurun_out/ncu_kd_normals.log 2>&1; echo "ncu kd_normals exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:kd_radius_kernel -c 1 -f -o gpurun_out/kd_radius_kernel \
    python scripts/gpu_kdtree_once.py 1 > gpurun_out/ncu_kd_radius.log 2>&1; echo "ncu kd_radius exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:kd_split_kernel -c 1 -f -o gpurun_out/kd_split_kernel \
    python scripts/gpu_kdtree_once.py 1 > gpurun_out/ncu_kd_split.log 2>&1; echo "ncu kd_split exit $?"
