#!/bin/bash
# Round-2 GPU pass N (1 GPU): 8-chain DMMA fold -- ICP / odometry tests, phases of both loops
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_icp_gpu.py tests/test_headline_gpu.py tests/test_odometry_gpu.py -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_pytest.log )
tail -4 gpurun_out/r02n_pytest.log | cut -c1-300
timeout 300 python scripts/gpu_icp_phases.py gpurun_out/r02n_icp_phases.json > gpurun_out/r02n_icp_phases.log 2>&1; echo "phases exit $?"
head -6 gpurun_out/r02n_icp_phases.log; tail -2 gpurun_out/r02n_icp_phases.log
timeout 300 python scripts/gpu_odo_once.py 2>&1 | head -5 | tee gpurun_out/r02n_odo.log
