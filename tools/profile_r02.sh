#!/bin/bash
# Run ON THE GPU BOX (under gpurun): round-2 profiles.
#  1. ncu launch list of a shortened default bench command (kernel SHARES of the step)
#  2. `ncu --set full` captures (DRAM bytes per launch, occupancy, stalls) of the kernels new or re-measured this round
#  3. compute-sanitizer memcheck / racecheck of golden-sized cases of every fused pipeline
set -u
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --set full --clock-control none -f"
# -- 1. launch list (C2, 1 gradient step after 3 warm-up steps; skip the observed-data modelling and the warm-up)
ncu --metrics gpu__time_duration.sum --clock-control none -s 72100 -c 2500 --csv --log-file $OUT/launches_r02p.csv \
    python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/launches_r02p.log 2>&1
# -- 2. full captures
C2="python bench.py --workload C2 --nt 400 --shots 10 --batch 10 --steps 1 --warmup 3 --no-cpu-baseline"
$NCU -k regex:ac_fwd_fused -s 700 -c 1 -o $OUT/prof_r02p_C2_ac_fwd_fused_recording $C2 > $OUT/prof_r02p_1.log 2>&1
$NCU -k regex:ac_adj_fused -s 100 -c 1 -o $OUT/prof_r02p_C2_ac_adj_fused $C2 > $OUT/prof_r02p_2.log 2>&1
$NCU -k regex:ac_fwd_fused -s 700 -c 1 -o $OUT/prof_r02p_C2rho_ac_fwd_fused_save2 $C2 --rho-grad > $OUT/prof_r02p_3.log 2>&1
$NCU -k regex:ac_adj_fused -s 100 -c 1 -o $OUT/prof_r02p_C2rho_ac_adj_fused_g2 $C2 --rho-grad > $OUT/prof_r02p_4.log 2>&1
C3="python bench.py --workload C3 --nt 400 --shots 15 --batch 15 --steps 1 --warmup 3 --no-cpu-baseline"
$NCU -k regex:ela_f -s 700 -c 1 -o $OUT/prof_r02p_C3abl_ela_f $C3 --abc gerjan > $OUT/prof_r02p_5.log 2>&1
$NCU -k regex:ela_b -s 100 -c 1 -o $OUT/prof_r02p_C3abl_ela_b $C3 --abc gerjan > $OUT/prof_r02p_6.log 2>&1
C4="python bench.py --workload C4 --nt 400 --shots 15 --batch 15 --steps 1 --warmup 3 --no-cpu-baseline"
$NCU -k regex:elf_f -s 700 -c 1 -o $OUT/prof_r02p_C4_elf_f $C4 > $OUT/prof_r02p_7.log 2>&1
$NCU -k regex:elf_b -s 100 -c 1 -o $OUT/prof_r02p_C4_elf_b $C4 > $OUT/prof_r02p_8.log 2>&1
C1="python bench.py --workload C1 --shots 40 --steps 1 --warmup 3 --no-cpu-baseline"
$NCU -k regex:acp_fwd -s 2 -c 1 -o $OUT/prof_r02p_C1_acp_fwd $C1 > $OUT/prof_r02p_9.log 2>&1
$NCU -k regex:acp_adj -s 1 -c 1 -o $OUT/prof_r02p_C1_acp_adj $C1 > $OUT/prof_r02p_10.log 2>&1
# the reports are too big to travel (64 MiB limit on gpurun_out): keep their raw metric pages as CSV, drop the reports
for R in $OUT/prof_r02p_*.ncu-rep; do
  ncu -i $R --page raw --csv > ${R%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $R --page details --csv > ${R%.ncu-rep}.details.csv 2>/dev/null
  rm -f $R
done
# -- 3. sanitizers
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > $OUT/sanitizer_r02p_$tool.log 2>&1
  echo "$tool rc=$?" >> $OUT/sanitizer_r02p_$tool.log
done
ls -la $OUT | grep r02p
tail -n 4 $OUT/sanitizer_r02p_memcheck.log; tail -n 6 $OUT/sanitizer_r02p_racecheck.log
