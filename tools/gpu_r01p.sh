#!/bin/bash
# Run ON THE GPU BOX (under gpurun): whole GPU parity suite, bench lines of all five workloads, ncu launch list + full
# captures of the two fused elastic kernels for the C3 bench command (shortened to 400 steps).
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/r01p_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r01p_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $OUT/r01p_pytest_gpu.log
for W in C2 C3 C1 C4 C5; do
  timeout 900 python bench.py --workload $W > $OUT/bench_r01p_$W.json 2> $OUT/bench_r01p_$W.err; echo "$W rc=$?"; cut -c1-330 $OUT/bench_r01p_$W.json
done
timeout 300 python bench.py --impl reference > $OUT/bench_r01p_ref.json 2> $OUT/bench_r01p_ref.err; echo "ref rc=$?"; cut -c1-200 $OUT/bench_r01p_ref.json
CMD="python bench.py --workload C3 --nt 400 --shots 15 --steps 1 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 600 --csv --log-file $OUT/launches_r01p.csv $CMD > $OUT/launches_r01p.log 2>&1
for K in elf_b elf_f; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2500 -c 1 -f -o $OUT/prof_r01p_$K $CMD > $OUT/prof_r01p_$K.log 2>&1
done
ls -la $OUT | tail -5
