#!/bin/bash
# Round-1 GPU pass: parity tests, bench (both arms), ncu launch list of the bench command, full captures of the top kernels.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
for k in integrate_kernel icp_search_kernel icp_accumulate_kernel odo_iteration_kernel odo_candidates_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o gpurun_out/full_$k \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu $k exit $?"
done
cat gpurun_out/bench_ours.json
