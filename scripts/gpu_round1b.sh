#!/bin/bash
# Round-1 second GPU pass: bench (both arms), neighbour-kernel timings, ncu launch list of the bench command, full captures.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log )
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 600 python tests/probes/gpu_probe_neighbours.py > gpurun_out/neighbours.json 2> gpurun_out/neighbours.err; echo "neighbours exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
for spec in integrate_pipelined_kernel:5 icp_loop_kernel:8 odo_loop_kernel:3; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/full_$k \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-odometry > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu $k exit $?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'bilateral_kernel|resample_fill_kernel|cluster_reduce_kernel|merge_kernel' -s 6 -c 6 -f -o gpurun_out/full_neighbours \
    python tests/probes/gpu_probe_neighbours.py > gpurun_out/ncu_full_neighbours.log 2>&1; echo "ncu neighbours exit $?"
cat gpurun_out/bench_ours.json; cat gpurun_out/neighbours.json
