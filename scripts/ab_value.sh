#!/bin/bash
# A/B of the device-resident step (`value`) and the fused step between the default library and
# every variant under lib/variants/:  bash scripts/ab_value.sh
run() {
  python bench.py --main-only --no-fused --no-parity-check --no-cpu-baseline --steps 120 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f  %.1f us/step  e2e %.0f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value']))"
  python scripts/fused_time.py A | tail -1
}
for rep in 1 2; do
  echo "== default"; run
  for f in pytorch-unsup-pc_b200/lib/variants/*.so; do
    echo "== $f"; DPC_B200_LIB=$PWD/$f run
  done
done
