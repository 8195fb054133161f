#!/bin/bash
# Round-2 GPU pass O (8 GPUs): bench.py --gpus 8 and --gpus 4 with the row-band upload and the peer-memory halo exchange
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/r02_8gpu_bench_n$N.json 2> gpurun_out/r02_8gpu_bench_n$N.err; echo "bench n$N exit $?"
tail -c 800 gpurun_out/r02_8gpu_bench_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    b = json.loads([l for l in open(f"gpurun_out/r02_8gpu_bench_n{N}.json") if l.startswith("{")][-1])
    print("value", b["value"], "ms", b["ms_per_step"], "scaling", b["scaling"], "e2e", b["e2e"]["value"])
    pf = b["partitioned_fusion"]
    print({k: v for k, v in pf.items() if k not in ("what", "e2e_note", "single_gpu_same_workload", "roofline_per_gpu")})
    print("replicas", b["replicas"]["value"], b["replicas"]["e2e"])
    print("config5", {k: v for k, v in b.get("dense_fusion_pipeline", {}).items() if k != "what"})
except Exception as e:
    print("bench parse failed", e)
PY
done
