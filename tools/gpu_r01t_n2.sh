#!/bin/bash
# Run ON A 2-GPU BOX (gpurun --gpus 2): the driver's own launch line for N=2, C2 and C3 workloads.
set -u
OUT=gpurun_out; mkdir -p $OUT
for W in C2 C3; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --workload $W > $OUT/bench_r01t_n2_$W.json 2> $OUT/bench_r01t_n2_$W.err; echo "$W n2 rc=$?"; tail -1 $OUT/bench_r01t_n2_$W.json | cut -c1-300
done
