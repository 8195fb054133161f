#!/bin/bash
# Round-2 GPU pass S (1 GPU): barrier-free odometry iteration (lazy candidates + pointer jumping): tests, phases, smoke
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_odometry_gpu.py tests/test_dropin_cpp.py tests/test_reference_mains.py -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02s_pytest.log )
tail -4 gpurun_out/r02s_pytest.log | cut -c1-300
for form in 1 2; do echo "== form $form"; OPB_ODO_PERSISTENT=$form timeout 300 python scripts/gpu_odo_once.py 2>&1 | head -4; done | tee gpurun_out/r02s_odo.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
