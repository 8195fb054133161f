#!/bin/bash
# Round-2 GPU pass G (2 GPUs): all GPU tests (incl. the 2-GPU worker and the peer-box halo exchange), odometry loop forms, bench N=1 and N=2
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g_pytest_gpu.log )
tail -15 gpurun_out/r02g_pytest_gpu.log
for form in 1 2 0; do
  echo "== OPB_ODO_PERSISTENT=$form"
  OPB_ODO_PERSISTENT=$form timeout 300 python scripts/gpu_odo_once.py 2>&1 | head -5
done > gpurun_out/r02g_odo_forms.log 2>&1
cat gpurun_out/r02g_odo_forms.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02g_bench.err
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02g_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"])
    print("odometry", b.get("dense_odometry"))
    print("parity", b.get("parity_check", {}).get("ok"))
    print("packed16", b.get("packed16_voxels"))
except Exception as e:
    print("bench parse failed", e)
PY
# A/B: the bulk-copy (cp.async.bulk + mbarrier) voxel update against the default pipelined kernel: parity first, then the kernel time
( OPB_INTEGRATE_BULK=1 timeout 600 python -m pytest tests/test_volume_gpu.py -m gpu -q > gpurun_out/r02g_pytest_bulk.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g_pytest_bulk.log )
tail -3 gpurun_out/r02g_pytest_bulk.log
for b in 0 1; do
OPB_INTEGRATE_BULK=$b timeout 600 python bench.py --steps 50 --warmup 5 --no-odometry --no-partitioned --no-cpu-baseline > gpurun_out/r02g_bench_bulk$b.json 2> gpurun_out/r02g_bench_bulk$b.err
python - $b <<'PY'
import json, sys
try:
    b = json.load(open(f"gpurun_out/r02g_bench_bulk{sys.argv[1]}.json"))
    print("OPB_INTEGRATE_BULK=" + sys.argv[1], "voxel_update ms", b["details"]["step_breakdown_ms"]["voxel_update"], "frac", b["roofline"]["frac"], "value", b["value"])
except Exception as e:
    print("bulk bench parse failed", e)
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; echo "bench n2 exit $?"
tail -c 800 gpurun_out/r02g_bench_n2.err
python - <<'PY'
import json
try:
    b = json.loads([l for l in open("gpurun_out/r02g_bench_n2.json") if l.startswith("{")][-1])
    print("value", b["value"], "ms", b["ms_per_step"], "scaling", b["scaling"], "e2e", b["e2e"]["value"])
    pf = b["partitioned_fusion"]
    print({k: v for k, v in pf.items() if k not in ("what", "e2e_note", "single_gpu_same_workload", "roofline_per_gpu")})
except Exception as e:
    print("bench parse failed", e)
PY
