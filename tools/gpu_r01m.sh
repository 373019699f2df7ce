#!/bin/bash
# Run ON THE GPU BOX (under gpurun): parity of the fused reverse kernel elf_b, then fused-vs-split A/B on a C3 slice.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_elastic_gpu.py -m gpu -x -q -k "pml_o4 or fused_large or model_level" > $OUT/r01m_pytest_el.log 2>&1; rc=$?; echo "pytest elastic(fused adjoint) rc=$rc"
tail -5 $OUT/r01m_pytest_el.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|FAILED" $OUT/r01m_pytest_el.log | head -20; fi
AB="--workload C3 --nt 400 --shots 15 --steps 2 --warmup 3"
timeout 300 python bench.py $AB > $OUT/r01m_c3s_fused.json 2> $OUT/r01m_c3s_fused.err; echo "c3 fused rc=$?"
ADFWI_B200_EL_ADJ_SPLIT=1 timeout 300 python bench.py $AB > $OUT/r01m_c3s_split.json 2> $OUT/r01m_c3s_split.err; echo "c3 split rc=$?"
python - <<'PY'
import json
for t in ("fused", "split"):
    try:
        d = json.loads(open(f"gpurun_out/r01m_c3s_{t}.json").read().strip().splitlines()[-1])
        print(t, round(d["value"], 2), d["roofline"]["per_kernel_avg_ms"], round(d["roofline"]["frac"], 3), d["roofline"].get("frac_by_sweep"))
    except Exception as e:
        print(t, "failed", e)
PY
tail -3 $OUT/r01m_c3s_fused.err
