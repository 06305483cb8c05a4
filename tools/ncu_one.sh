#!/bin/bash
# One `ncu --set full` capture of a kernel of the C2 bench step (run under gpurun; report lands in gpurun_out/).
# usage: tools/ncu_one.sh <kernel-regex> <tag> [extra bench args]
set -e
K=${1:-tdl_window}
TAG=${2:-window}
shift 2 || true
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/ncu_$TAG \
  python bench.py --steps 2 --warmup 3 --links 2048 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$TAG.log 2>&1 || tail -5 gpurun_out/ncu_$TAG.log
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page details > gpurun_out/ncu_${TAG}_details.txt
tail -2 gpurun_out/ncu_$TAG.log
