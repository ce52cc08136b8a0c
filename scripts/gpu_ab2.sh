#!/bin/bash
# A/B of bench.py flag sets:  bash scripts/gpu_ab2.sh "--workload B" "--workload B --global-grid"
mkdir -p gpurun_out
for flags in "$@"; do
  echo "== bench.py $flags"
  python bench.py --steps 60 --warmup 6 --no-cpu-baseline $flags 2> gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f  us/step %.1f  e2e %.0f  step-roofline %.3f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['roofline_step']['frac']))
print({k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
"
  tail -2 gpurun_out/ab_err.log
done
