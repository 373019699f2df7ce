#!/bin/bash
# usage: tools/profile2.sh <tag> <quick_bench args...> -- kernel regexes
set -u
TAG=$1; shift
ARGS=""
while [ "$1" != "--" ]; do ARGS="$ARGS $1"; shift; done; shift
OUT=gpurun_out; mkdir -p $OUT
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o $OUT/prof_${TAG}_$K python tools/quick_bench.py $ARGS > $OUT/prof_${TAG}_$K.log 2>&1
done
ls -la $OUT | tail -5
