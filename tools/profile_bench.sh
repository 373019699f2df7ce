#!/bin/bash
# Run ON THE GPU BOX (under gpurun): ncu launch list and full captures of the dominant kernels for the
# contract bench command itself.  usage: tools/profile_bench.sh <tag>
set -u
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/launches_$TAG.log 2>&1
for K in ac_adj_fused ac_fwd_fused; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 500 -c 1 -f -o $OUT/prof_${TAG}_$K $CMD > $OUT/prof_${TAG}_$K.log 2>&1
done
ls -la $OUT | tail -6
