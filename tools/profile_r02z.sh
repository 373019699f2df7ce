#!/bin/bash
# Run ON THE GPU BOX: source-level capture of the split-PML reverse kernel elf_b on the C3 grid (30 shots per launch, as at full length)
set -u
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --set full --import-source on --clock-control none -f"
C3="python bench.py --workload C3 --nt 300 --shots 30 --batch 30 --steps 1 --warmup 3 --no-cpu-baseline"
$NCU -k regex:elf_b -s 100 -c 1 -o $OUT/prof_r02z_C3_elf_b $C3 > $OUT/prof_r02z_1.log 2>&1
for R in $OUT/prof_r02z_*.ncu-rep; do
  ncu -i $R --page raw --csv > ${R%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $R --page source --csv --print-source cuda,sass > ${R%.ncu-rep}.source.csv 2>/dev/null
  rm -f $R
done
ls -la $OUT | grep r02z_C3
