#!/bin/bash
# Round-2 GPU pass I (1 GPU): all GPU tests (with the reference mains), smoke()
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q -s -k "mains or packed or fusion or odometry" > gpurun_out/r02i_pytest_focus.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest_focus.log )
grep -n "main:\|main,\|packed16 vs\|default loop form\|passed\|failed" gpurun_out/r02i_pytest_focus.log | tail -12
( timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest_gpu.log )
tail -8 gpurun_out/r02i_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r02i_smoke.log
