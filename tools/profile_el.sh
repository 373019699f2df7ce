#!/bin/bash
# Run ON THE GPU BOX (under gpurun): ncu launch list + one full capture per kernel regex of the elastic probe.
# usage: tools/profile_el.sh <tag> "<quick_bench_el args>" <skip> kernel-regex...
set -u
TAG=$1; ARGS=$2; SKIP=$3; shift 3
OUT=gpurun_out; mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file $OUT/launches_$TAG.csv python tools/quick_bench_el.py $ARGS > $OUT/launches_$TAG.log 2>&1
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $OUT/prof_${TAG}_$K python tools/quick_bench_el.py $ARGS > $OUT/prof_${TAG}_$K.log 2>&1
done
ls -la $OUT | tail -8
