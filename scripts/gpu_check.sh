#!/bin/bash
# Runs the GPU test groups in separate processes (a hung kernel only loses its own group); logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  timeout -s KILL $to python -m pytest "$@" -q -s -m gpu -p no:cacheprovider > gpurun_out/test_$name.log 2>&1
  echo "== $name: exit $? =="; tail -n 25 gpurun_out/test_$name.log | cut -c1-300
}
run linear 240 tests/test_kernels_gpu.py -k "linear"
run ln_knn 200 tests/test_kernels_gpu.py -k "layernorm or knn"
run sample 200 tests/test_kernels_gpu.py -k "project_sample"
run mha 240 tests/test_kernels_gpu.py -k "mha"
run vecattn 300 tests/test_kernels_gpu.py -k "vector_attention"
run parity 900 tests/test_parity_gpu.py
run hrnet 400 tests/test_hrnet_gpu.py
run model 400 tests/test_model_gpu.py
run metrics 200 tests/test_metrics_gpu.py
run parametric 300 tests/test_parametric_gpu.py
run train 600 tests/test_train_gpu.py
