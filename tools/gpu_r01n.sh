#!/bin/bash
# Run ON THE GPU BOX (under gpurun): full elastic parity suite, then the C3 slice bench.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_elastic_gpu.py tests/test_properties_gpu.py -m gpu -x -q -k "elastic" > $OUT/r01n_pytest_el.log 2>&1; rc=$?; echo "pytest elastic rc=$rc"
tail -5 $OUT/r01n_pytest_el.log
AB="--workload C3 --nt 400 --shots 15 --steps 2 --warmup 3"
timeout 300 python bench.py $AB > $OUT/r01n_c3s.json 2> $OUT/r01n_c3s.err; echo "c3 slice rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r01n_c3s.json").read().strip().splitlines()[-1])
print(round(d["value"], 2), d["roofline"]["per_kernel_avg_ms"], round(d["roofline"]["frac"], 3), d["roofline"].get("frac_by_sweep"))
PY
