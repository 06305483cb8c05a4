#!/bin/bash
TAG=r01b
cap() { local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/ncu_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1 || tail -3 gpurun_out/ncu_${TAG}_$name.log
  ncu -i gpurun_out/ncu_${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_${name}_raw.csv
  ncu -i gpurun_out/ncu_${TAG}_$name.ncu-rep --page details > gpurun_out/ncu_${TAG}_${name}_details.txt
  ncu -i gpurun_out/ncu_${TAG}_$name.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_${name}_src.csv 2>/dev/null
  rm -f gpurun_out/ncu_${TAG}_$name.ncu-rep; }
cap gemm spatial_gemm 2 env B=64 python tools/bench_spatial_gemm.py
cap zmode tdl_tma 2 python tools/config_report.py --configs C4 --steps 1
cap c5 tdl_tma 2 python tools/config_report.py --configs C5 --steps 1
