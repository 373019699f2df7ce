#!/bin/bash
# Run ON THE GPU BOX (under gpurun): GPU parity suite, lean-adjoint A/B on the C3 grid, then the C2 / C3 bench lines.
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/r01k_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_elastic_gpu.py -m gpu -x -q -k "fused_large or pml" > $OUT/r01k_pytest_el.log 2>&1; echo "pytest elastic(pml) rc=$?"
tail -3 $OUT/r01k_pytest_el.log
AB="--workload C3 --nt 400 --shots 15 --steps 2 --warmup 3"
timeout 300 python bench.py $AB > $OUT/r01k_c3s_lean.json 2> $OUT/r01k_c3s_lean.err; echo "c3 lean rc=$?"
ADFWI_B200_EL_LEAN=0 timeout 300 python bench.py $AB > $OUT/r01k_c3s_full.json 2> $OUT/r01k_c3s_full.err; echo "c3 full rc=$?"
python - <<'PY'
import json
for t in ("lean", "full"):
    try:
        d = json.loads(open(f"gpurun_out/r01k_c3s_{t}.json").read().strip().splitlines()[-1])
        print(t, round(d["value"], 2), d["roofline"]["per_kernel_avg_ms"], round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(t, "failed", e)
PY
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r01k_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -3 $OUT/r01k_pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_r01k_C2.json 2> $OUT/bench_r01k_C2.err; echo "C2 rc=$?"; cut -c1-400 $OUT/bench_r01k_C2.json
timeout 900 python bench.py --workload C3 > $OUT/bench_r01k_C3.json 2> $OUT/bench_r01k_C3.err; echo "C3 rc=$?"; cut -c1-400 $OUT/bench_r01k_C3.json
