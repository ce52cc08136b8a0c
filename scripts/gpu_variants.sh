#!/bin/bash
# stage timings of the default library and every variant under lib/variants/ (A/B of kernel builds)
wl=${1:-A}
python scripts/stage_time.py $wl 2>&1 | tail -1
for f in pytorch-unsup-pc_b200/lib/variants/*.so; do
  DPC_B200_LIB=$PWD/$f python scripts/stage_time.py $wl 2>&1 | tail -1
done
