#!/bin/bash
# Run ON THE GPU BOX (under gpurun): ncu launch list and full captures of the elastic kernels for the bench command
# (C3 workload, shortened to 400 steps so that the ncu replay stays within minutes).  usage: tools/profile_bench_el.sh <tag> [kernels...]
set -u
TAG=$1; shift
KERNELS=${@:-"elf_f elf_k1 elf_k2"}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload C3 --nt 400 --shots 15 --steps 1 --warmup 3"
if [ ! -f $OUT/launches_$TAG.csv ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/launches_$TAG.log 2>&1
fi
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 300 -c 1 -f -o $OUT/prof_${TAG}_$K $CMD > $OUT/prof_${TAG}_$K.log 2>&1
done
ls -la $OUT | tail -6
