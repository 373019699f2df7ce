#!/bin/bash
# Run ON THE GPU BOX (under gpurun).  Writes ncu launch list + full captures into gpurun_out/.
# usage: tools/profile.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/quick_bench.py 350 1700 8 60 0"
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/launches_$TAG.log 2>&1
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 2 -f -o $OUT/prof_${TAG}_$K $CMD > $OUT/prof_${TAG}_$K.log 2>&1
done
ls -la $OUT
