for c in 64 32 16 8 4; do
  echo "DPC_CHUNK=$c (graph)"; DPC_CHUNK=$c python bench.py --steps 300 --warmup 20 --no-cpu-baseline --graph 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(' value %.0f  us/step %.1f'%(d['value'], d['ms_per_step']*1e3))"
done
