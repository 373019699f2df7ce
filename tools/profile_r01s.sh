#!/bin/bash
# Run ON THE GPU BOX (under gpurun): ncu launch list + full captures of the two fused elastic kernels for the C3 bench
# command shortened to 400 steps / 15 shots (so that the ncu replay stays within minutes).  usage: tools/profile_r01s.sh <tag>
set -u
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload C3 --nt 400 --shots 15 --batch 15 --steps 1 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:elf_b -s 300 -c 1 -f -o $OUT/prof_${TAG}_elf_b $CMD > $OUT/prof_${TAG}_elf_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:elf_f -s 2500 -c 1 -f -o $OUT/prof_${TAG}_elf_f $CMD > $OUT/prof_${TAG}_elf_f.log 2>&1
ls -la $OUT | grep $TAG
