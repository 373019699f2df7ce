#!/bin/bash
# Run ON THE GPU BOX (under gpurun): full ncu captures of the two fused acoustic kernels on the C5 grid (8 shots per launch,
# 250-step slice, checkpointed every 125 steps): the pure DRAM-streaming case (1.7 GB of state per launch).
set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload C5 --nt 250 --steps 1 --warmup 3"
for K in ac_adj_fused ac_fwd_fused; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 400 -c 1 -f -o $OUT/prof_r01z_$K $CMD > $OUT/prof_r01z_$K.log 2>&1
done
ls -la $OUT | grep r01z
