#!/bin/bash
# launch list of ONE training step (POEM-medium, 8 views, batch $1) -> gpurun_out/train_launches_b$1.csv
B=${1:-32}
POEM_TRAIN_WARM=0 POEM_TRAIN_ITERS=1 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
  --log-file gpurun_out/train_launches_b$B.csv python scripts/bench_train.py medium 8 $B > gpurun_out/train_under_ncu_b$B.log 2>&1
echo "ncu rc=$? lines=$(wc -l < gpurun_out/train_launches_b$B.csv)"
