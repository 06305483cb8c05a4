#!/bin/bash
# `ncu --set full` captures of the kernels the C2 bench does not exercise (run under gpurun, one GPU).
# usage: tools/ncu_kernels.sh <round-tag>     -> gpurun_out/ncu_<tag>_{gemm,cdl,coef,zmode}*
TAG=${1:-r01}
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/ncu_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1 || tail -3 gpurun_out/ncu_${TAG}_$name.log
  ncu -i gpurun_out/ncu_${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_${name}_raw.csv
  ncu -i gpurun_out/ncu_${TAG}_$name.ncu-rep --page details > gpurun_out/ncu_${TAG}_${name}_details.txt
  ncu -i gpurun_out/ncu_${TAG}_$name.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_${name}_src.csv 2>/dev/null
  rm -f gpurun_out/ncu_${TAG}_$name.ncu-rep
}
cap gemm spatial_gemm 2 env B=64 python tools/bench_spatial_gemm.py
cap cdl_propagate cdl_poly 1 python tools/config_report.py --configs C3 --steps 1
cap cdl_moment cdl_moment 1 python tools/config_report.py --configs C3 --steps 1
cap cdl_rays cdl_ray 1 python tools/config_report.py --configs C3 --steps 1
cap coef sos_poly_coef 2 python bench.py --steps 2 --warmup 3 --links 2048 --no-cpu-baseline --no-e2e
cap zmode tdl_tma 20 python tools/config_report.py --configs C4 --steps 1
ls -la gpurun_out | grep ncu_${TAG}
