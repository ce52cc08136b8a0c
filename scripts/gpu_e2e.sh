#!/bin/bash
# e2e A/B: "<lanes>x<steps per captured graph>" ...
for cfg in "$@"; do
  l=${cfg%x*}; g=${cfg#*x}
  DPC_E2E_LANES=$l DPC_E2E_GRAPH_STEPS=$g python bench.py --steps ${STEPS:-240} --warmup 6 --no-cpu-baseline 2> gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lanes $l x $g steps: value %.0f  e2e %.0f (%.1f us)  e2e_rep %.0f (%.1f us)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']*1e3, d['e2e_replica_aware']['value'], d['e2e_replica_aware']['ms_per_step']*1e3))
"
  tail -2 gpurun_out/ab_err.log
done
