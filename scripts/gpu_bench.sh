#!/bin/bash
# bench line + ncu launch list (per-launch device times) for the default workload
mkdir -p gpurun_out
python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $? =="; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "${NCU_LIST:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
  echo "== ncu list exit $? =="; wc -l gpurun_out/launches.csv
fi
