#!/bin/bash
# Round-2 GPU pass A: parity tests (incl. the new headline-size and unconverged-ICP cases), ICP per-pass phase times,
# full ncu captures of the two persistent solver loops.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest_gpu.log )
tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 300 python scripts/gpu_icp_phases.py gpurun_out/r02a_icp_phases.json > gpurun_out/r02a_icp_phases.log 2>&1; echo "phases exit $?"
tail -8 gpurun_out/r02a_icp_phases.log
timeout 300 python scripts/gpu_odo_once.py > gpurun_out/r02a_odo_once.log 2>&1; echo "odo exit $?"; tail -3 gpurun_out/r02a_odo_once.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:odo_loop_kernel -c 1 -f -o gpurun_out/r02a_full_odo_loop_kernel \
    python scripts/gpu_odo_once.py > gpurun_out/r02a_ncu_odo.log 2>&1; echo "ncu odo exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:icp_loop_kernel -c 1 -f -o gpurun_out/r02a_full_icp_loop_kernel \
    python scripts/gpu_icp_once.py > gpurun_out/r02a_ncu_icp.log 2>&1; echo "ncu icp exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r02a_bench.json
