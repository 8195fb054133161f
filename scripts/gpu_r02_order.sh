#!/bin/bash
# Round-2 GPU pass (1 GPU): reference-order mesh emission -- mains, drop-in, volume / mesh tests
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_reference_mains.py tests/test_dropin_cpp.py tests/test_volume_gpu.py tests/test_meshpost_gpu.py tests/test_resample_gpu.py tests/test_fusion_gpu.py -m gpu -q -s > gpurun_out/r02ord_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ord_pytest.log )
grep -n "main:\|main,\|bytes differ\|passed\|failed\|FAILED\|Error" gpurun_out/r02ord_pytest.log | cut -c1-400 | tail -12
