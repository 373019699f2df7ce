#!/bin/bash
# Run ON THE GPU BOX: batch-size sweep on C3 / C4 slices.
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift; timeout 600 python bench.py "$@" > $OUT/r01q_$tag.json 2> $OUT/r01q_$tag.err; python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r01q_{t}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(t, round(d["value"],2), {k:round(v,4) for k,v in r["per_kernel_avg_ms"].items()}, r.get("frac_by_sweep"))
except Exception as e: print(t,"failed",e)
PY
}
run c3_b15 --workload C3 --nt 400 --shots 30 --batch 15 --steps 2
run c3_b30 --workload C3 --nt 400 --shots 30 --batch 30 --steps 2
run c3_b10 --workload C3 --nt 400 --shots 30 --batch 10 --steps 2
run c4_b5 --workload C4 --nt 800 --shots 15 --batch 5 --steps 2
run c4_b15 --workload C4 --nt 800 --shots 15 --batch 15 --steps 2
