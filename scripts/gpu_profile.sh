#!/bin/bash
# One GPU call: the ncu evidence for profiles/ (B200_PROFILING.md recipe).
#   1. the bench command exits 0 without ncu, then its launch list (gpu__time_duration per launch)
#   2. one `--set full` capture of a whole-batch step (DPC_CHUNK=P: one launch per kernel)
# usage: bash scripts/gpu_profile.sh <tag>     -> gpurun_out/launches_<tag>.csv, prof_<tag>.ncu-rep
tag=${1:-rxx}
mkdir -p gpurun_out
BENCH="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --main-only --no-parity-check --no-fused"
$BENCH > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log || { echo "bench failed"; tail -5 gpurun_out/bench_${tag}_err.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$tag.csv $BENCH > gpurun_out/ncu_$tag.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_$tag.csv)"
DPC_CHUNK=100000 python scripts/profile_step.py --steps 2 > gpurun_out/plain_$tag.log 2>&1 || { echo "profile_step failed"; exit 1; }
DPC_CHUNK=100000 timeout 600 ncu --set full --clock-control none --import-source on -f \
    -o gpurun_out/prof_$tag python scripts/profile_step.py --steps 2 > gpurun_out/ncu_full_$tag.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_$tag.ncu-rep
# summaries are made here: the report itself may exceed what gpurun copies back (64 MiB in all)
python scripts/ncu_summary.py gpurun_out/prof_$tag.ncu-rep > gpurun_out/ncu_full_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
python scripts/ncu_sass_mix.py gpurun_out/prof_$tag.ncu-rep > gpurun_out/sass_mix_$tag.txt 2>/dev/null
rm -f gpurun_out/prof_$tag.ncu-rep
ls -la gpurun_out/ | tail -8
