#!/bin/bash
# Round-2 GPU pass Q (2 GPUs): odometry with pointer-jumping chains (tests, phases), bench N=2 with slabs of 4
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_odometry_gpu.py tests/test_dropin_cpp.py -m gpu -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_pytest.log )
tail -4 gpurun_out/r02q_pytest.log | cut -c1-300
for form in 1 2; do echo "== form $form"; OPB_ODO_PERSISTENT=$form timeout 300 python scripts/gpu_odo_once.py 2>&1 | head -4; done | tee gpurun_out/r02q_odo.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err; echo "bench n2 exit $?"
python - <<'PY'
import json
try:
    b = json.loads([l for l in open("gpurun_out/r02q_bench_n2.json") if l.startswith("{")][-1])
    print("value", b["value"], "ms", b["ms_per_step"], "scaling", b["scaling"], "e2e", b["e2e"]["value"])
    pf = b["partitioned_fusion"]
    print({k: v for k, v in pf.items() if k not in ("what", "e2e_note", "single_gpu_same_workload", "roofline_per_gpu")})
    print("replicas", b["replicas"]["value"], b["replicas"]["e2e"])
except Exception as e:
    print("bench parse failed", e)
PY
