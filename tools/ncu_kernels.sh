#!/bin/bash
# `ncu --set full` captures of the dominant kernel of every BASELINE config + the launch list of a bench step.
# usage: tools/ncu_kernels.sh <round-tag>     -> gpurun_out/ncu_<tag>_*   (run under gpurun, one GPU)
TAG=${1:-r02}
mkdir -p gpurun_out
tools/ncu_one.sh tdl_tma ${TAG}_c2 2 --config C2
tools/ncu_one.sh sos_poly_coef ${TAG}_c2_coef 2 --config C2
tools/ncu_one.sh "tdl_tma|fused" ${TAG}_c4 2 --config C4
tools/ncu_one.sh spatial_gemm ${TAG}_c4_gemm 2 --config C4
tools/ncu_one.sh tdl_tma ${TAG}_c5 2 --config C5
tools/ncu_one.sh "cdl_poly|cdl_tc" ${TAG}_c3 1 --config C3
tools/ncu_one.sh tdl_ ${TAG}_c1 2 --config C1
for c in C1 C2 C3 C4 C5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}_$c.csv \
    python bench.py --config $c --only --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
ls -la gpurun_out | grep ${TAG}
