#!/bin/bash
# Round-2 GPU pass H (1 GPU): all GPU tests, odometry loop forms, bench, launch list of the bench + full ncu captures
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest_gpu.log )
tail -15 gpurun_out/r02h_pytest_gpu.log
( OPB_INTEGRATE_BULK=1 timeout 600 python -m pytest tests/test_volume_gpu.py -m gpu -q > gpurun_out/r02h_pytest_bulk.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest_bulk.log )
tail -3 gpurun_out/r02h_pytest_bulk.log
for form in 1 2; do
  echo "== OPB_ODO_PERSISTENT=$form"
  OPB_ODO_PERSISTENT=$form timeout 300 python scripts/gpu_odo_once.py 2>&1 | head -5
done > gpurun_out/r02h_odo_forms.log 2>&1
cat gpurun_out/r02h_odo_forms.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02h_bench.err
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/r02h_bench.json"))
    print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"])
    print(b["details"]["step_breakdown_ms"], "roofline", b["roofline"]["frac"])
    o = b.get("dense_odometry", {})
    print("odometry", o.get("value"), o.get("device_ms_per_frame"), o.get("e2e"))
    print("parity", b.get("parity_check", {}).get("ok"))
    print("packed16", b.get("packed16_voxels"))
    print("config4 N=1", {k: v for k, v in b.get("partitioned_fusion", {}).items() if k in ("frames_per_s", "e2e_frames_per_s", "error")})
except Exception as e:
    print("bench parse failed", e)
PY
# launch list of the same bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02h_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-odometry --no-partitioned --no-cpu-baseline > gpurun_out/r02h_ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:odo_loop2_kernel -c 1 -f -o gpurun_out/r02h_full_odo_loop2_kernel \
    python scripts/gpu_odo_once.py > gpurun_out/r02h_ncu_odo.log 2>&1; echo "ncu odo exit $?"
OPB_INTEGRATE_BULK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:integrate_bulk_kernel -s 12 -c 1 -f -o gpurun_out/r02h_full_integrate_bulk_kernel \
    python bench.py --steps 3 --warmup 3 --no-odometry --no-partitioned --no-cpu-baseline > gpurun_out/r02h_ncu_bulk.log 2>&1; echo "ncu bulk exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:integrate_pipelined_kernel -s 12 -c 1 -f -o gpurun_out/r02h_full_integrate_pipelined_kernel \
    python bench.py --steps 3 --warmup 3 --no-odometry --no-partitioned --no-cpu-baseline > gpurun_out/r02h_ncu_pipe.log 2>&1; echo "ncu pipelined exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:integrate_packed_kernel -s 12 -c 1 -f -o gpurun_out/r02h_full_integrate_packed_kernel \
    python bench.py --steps 3 --warmup 3 --no-odometry --no-cpu-baseline > gpurun_out/r02h_ncu_packed.log 2>&1; echo "ncu packed exit $?"
ls -la gpurun_out/*.ncu-rep | tail -6
