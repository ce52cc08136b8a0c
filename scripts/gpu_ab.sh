#!/bin/bash
# A/B timing helper: runs bench.py (no CPU baseline) once per environment setting given as args,
# e.g.  bash scripts/gpu_ab.sh "" "DPC_DRC_BWD_STAGE=1"
mkdir -p gpurun_out
for envs in "$@"; do
  echo "== env: '$envs'"
  env $envs python bench.py --steps 90 --warmup 6 --no-cpu-baseline 2> gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f  us/step %.1f  e2e %.0f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value']))
print({k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
"
  tail -2 gpurun_out/ab_err.log
done
