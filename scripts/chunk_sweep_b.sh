#!/bin/bash
# value of workload B (128^3) over the chunk size of the two-stream split (L2 residency of the
# producer -> consumer grids: 8 MiB per projection)
for c in ${CHUNKS:-0 32 16 8 4}; do
  echo -n "DPC_CHUNK=$c: "
  DPC_CHUNK=$c python bench.py --workload ${1:-B} --main-only --no-fused --no-parity-check --no-cpu-baseline --steps 30 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f  %.1f us/step  e2e %.0f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value']))"
done
