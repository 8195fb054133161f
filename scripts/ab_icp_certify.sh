#!/bin/bash
# A/B of the nearest-neighbour certificates in the ICP loop: bench step breakdown with them on / off and for several guards
for cfg in "OPB_ICP_FUSED_GRID=0" "OPB_ICP_FUSED_GRID=1" "OPB_ICP_FUSED_GRID=0" "OPB_ICP_FUSED_GRID=1"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-odometry | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('fps %.1f  e2e %.1f  breakdown %s' % (d['value'], d['e2e']['value'], d['config']['step_breakdown_ms']))"
done
