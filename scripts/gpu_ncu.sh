#!/bin/bash
# ncu --set full captures of the hot kernels (one launch each), reports into gpurun_out/
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph"
cap() { # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "== ncu $1 exit $? =="; ls -la gpurun_out/$1.ncu-rep 2>/dev/null
}
for k in ${NCU_KERNELS:-va_fused mha}; do
  case $k in
    va_fused) cap va_fused va_fused 8;;
    mha) cap mha mha_fwd 6;;
    merge0a) cap gemm_merge0a gemm_op16 1;;
    ptproj) cap gemm_ptproj gemm_op16 5;;
    knn) cap knn knn32 5;;
    sample) cap sample project_sample 2;;
  esac
done
