#!/bin/bash
# Round-2 GPU pass AC (2 GPUs): the multi-GPU tests (2-rank worker: partitioned fusion, split ICP, row-band frames, peer halo), bench N=2
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_fusion_gpu.py -m gpu -q -s > gpurun_out/r02_2gpu_pytest_fusion.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_2gpu_pytest_fusion.log )
grep -n "config 4 slice\|config 5 slice\|ICP whole\|ICP split\|passed\|failed" gpurun_out/r02_2gpu_pytest_fusion.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r02_2gpu_bench_n2.json 2> gpurun_out/r02_2gpu_bench_n2.err; echo "bench n2 exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r02_2gpu_bench_ref_n2.json 2> gpurun_out/r02_2gpu_bench_ref_n2.err; echo "bench ref n2 exit $?"
python - <<'PY'
import json
try:
    b = json.loads([l for l in open("gpurun_out/r02_2gpu_bench_n2.json") if l.startswith("{")][-1])
    print("value", b["value"], "ms", b["ms_per_step"], "scaling", b["scaling"], "e2e", b["e2e"]["value"], "parity", b["partitioned_fusion"]["partition_parity"])
    r = json.loads([l for l in open("gpurun_out/r02_2gpu_bench_ref_n2.json") if l.startswith("{")][-1])
    print("reference arm", r["value"], r["steps"], r["cpu_baseline"]["cores"], "same config", r["config"] == b["config"], r["cpu_baseline"]["sample"][:120])
except Exception as e:
    print("bench parse failed", e)
PY
