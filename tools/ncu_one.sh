#!/bin/bash
# One `ncu --set full` capture of a kernel of a bench step (run under gpurun; summaries land in gpurun_out/).
# usage: tools/ncu_one.sh <kernel-regex> <tag> <launches-to-skip> [bench args, e.g. --config C4]
set -e
K=${1:-tdl_tma}
TAG=${2:-c2}
SKIP=${3:-2}
shift 3 || true
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/ncu_$TAG \
  python bench.py --steps 2 --warmup 3 --only --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$TAG.log 2>&1 || tail -5 gpurun_out/ncu_$TAG.log
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page details > gpurun_out/ncu_${TAG}_details.txt
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_src.csv 2>/dev/null || true
rm -f gpurun_out/ncu_$TAG.ncu-rep
tail -2 gpurun_out/ncu_$TAG.log
