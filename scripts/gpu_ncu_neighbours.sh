#!/bin/bash
# one `ncu --set full` capture per kernel either side of the hot path
mkdir -p gpurun_out
for k in bilateral_kernel resample_fill_kernel merge_kernel cluster_reduce_kernel normals_vertex_kernel mc_emit_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -f -o gpurun_out/nb_$k \
      python tests/probes/gpu_probe_neighbours.py > gpurun_out/ncu_nb_$k.log 2>&1; echo "ncu $k exit $?"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:estimate_normals_kernel -s 1 -c 1 -f -o gpurun_out/nb_estimate_normals_kernel \
    python tests/probes/gpu_probe_normals.py > gpurun_out/ncu_nb_estimate_normals.log 2>&1; echo "ncu estimate_normals exit $?"
