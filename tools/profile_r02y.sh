#!/bin/bash
# Run ON THE GPU BOX: source-level captures (per-line stall samples) of the acoustic fused kernels on the C2 grid.
set -u
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --set full --import-source on --clock-control none -f"
C2="python bench.py --workload C2 --nt 400 --shots 10 --batch 10 --steps 1 --warmup 3 --no-cpu-baseline"
$NCU -k regex:ac_fwd_fused -s 700 -c 1 -o $OUT/prof_r02y_C2_ac_fwd $C2 > $OUT/prof_r02y_1.log 2>&1
$NCU -k regex:ac_adj_fused -s 100 -c 1 -o $OUT/prof_r02y_C2_ac_adj $C2 > $OUT/prof_r02y_2.log 2>&1
for R in $OUT/prof_r02y_*.ncu-rep; do
  ncu -i $R --page raw --csv > ${R%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $R --page source --csv --print-source cuda,sass > ${R%.ncu-rep}.source.csv 2>/dev/null
  rm -f $R
done
ls -la $OUT | grep r02y
