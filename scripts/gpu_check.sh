#!/bin/bash
# One GPU call: parity tests, smoke, a short bench.  Output under gpurun_out/.
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 60 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value %.0f  us/step %.1f  e2e %.0f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value']))
print({k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'step', round(d['roofline_step']['frac'],3), 'cpu', d['cpu_baseline'])
"
tail -3 gpurun_out/bench_err.log
