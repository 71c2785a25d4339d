#!/bin/bash
# Round-end evidence run on one B200: all GPU tests, the full bench line, the reference arm, the small-batch table,
# the ncu launch list and ncu --set full captures of the three largest kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
bash scripts/gpu_check.sh 2>&1 | grep -E "==|passed|failed"
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "reference rc=$?"
python scripts/bench_small_batch.py > gpurun_out/r2_small_batch.json 2>/dev/null
# training path: the step line, eager-vs-graph table, live per-primitive profile, launch list
python bench.py --train-only > gpurun_out/r2_bench_train.json 2> gpurun_out/r2_bench_train.err; echo "train bench rc=$?"
bash scripts/train_gpu_checks.sh > gpurun_out/r2_train_checks.log 2>&1; echo "train checks rc=$?"
POEM_TRAIN_PROF=1 python scripts/bench_train.py medium 8 32 > gpurun_out/r2_train_profile_b32.txt 2>&1
bash scripts/ncu_train.sh 32 > /dev/null 2>&1; bash scripts/ncu_train.sh 4 > /dev/null 2>&1; echo "train launch lists rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --lean --steps 2 --warmup 1 --no-graph --min-timed-s 0.01 > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches.csv)"
CMD="python bench.py --lean --steps 1 --warmup 1 --no-graph --min-timed-s 0.01"
for k in va_fused mha_fwd sample_merge; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/r2_$k $CMD > gpurun_out/ncu_r2_$k.log 2>&1
  echo "ncu $k rc=$?"
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
