#!/bin/bash
# Run ON THE GPU BOX: source-level captures (per-line stall samples) of the elastic fused forward kernels on the C3 grid.
set -u
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --set full --import-source on --clock-control none -f"
C3="python bench.py --workload C3 --nt 400 --shots 15 --batch 15 --steps 1 --warmup 3 --no-cpu-baseline"
$NCU -k regex:ela_f -s 700 -c 1 -o $OUT/prof_r02x_C3abl_ela_f $C3 --abc gerjan > $OUT/prof_r02x_1.log 2>&1
$NCU -k regex:elf_f -s 700 -c 1 -o $OUT/prof_r02x_C3_elf_f $C3 > $OUT/prof_r02x_2.log 2>&1
$NCU -k regex:ela_b -s 100 -c 1 -o $OUT/prof_r02x_C3abl_ela_b $C3 --abc gerjan > $OUT/prof_r02x_3.log 2>&1
for R in $OUT/prof_r02x_*.ncu-rep; do
  ncu -i $R --page raw --csv > ${R%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $R --page source --csv --print-source cuda,sass > ${R%.ncu-rep}.source.csv 2>/dev/null
  rm -f $R
done
ls -la $OUT | grep r02x
