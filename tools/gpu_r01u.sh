#!/bin/bash
# Run ON THE GPU BOX (under gpurun): whole GPU parity suite, smoke(), then the bench lines C2 (default), C3, C4.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r01u_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $OUT/r01u_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for W in C2 C3 C4; do
  timeout 900 python bench.py --workload $W > $OUT/bench_r01u_$W.json 2> $OUT/bench_r01u_$W.err; echo "$W rc=$?"; cut -c1-200 $OUT/bench_r01u_$W.json
done
