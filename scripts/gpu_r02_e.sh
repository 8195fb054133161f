#!/bin/bash
# Round-2 GPU pass E (2 GPUs): the multi-GPU tests, then bench.py --gpus 2 (config-4 partitioned fusion headline)
mkdir -p gpurun_out
nvidia-smi -L
( timeout 900 python -m pytest tests/test_fusion_gpu.py -m gpu -q > gpurun_out/r02e_pytest_fusion.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest_fusion.log )
tail -6 gpurun_out/r02e_pytest_fusion.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; echo "bench n2 exit $?"
tail -c 1500 gpurun_out/r02e_bench_n2.err
python - <<'PY'
import json
try:
    b = json.loads([l for l in open("gpurun_out/r02e_bench_n2.json") if l.startswith("{")][-1])
    print("value", b["value"], "ms", b["ms_per_step"], "scaling", b["scaling"], "e2e", b["e2e"]["value"])
    pf = b["partitioned_fusion"]
    print({k: v for k, v in pf.items() if k not in ("what", "e2e_note")})
    print("replicas", b["replicas"]["value"], b["replicas"]["e2e"])
    print("config5", b.get("dense_fusion_pipeline"))
except Exception as e:
    print("bench parse failed", e)
PY
