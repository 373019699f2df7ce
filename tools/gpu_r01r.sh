#!/bin/bash
# Run ON THE GPU BOX: shots-per-chunk sweep (tile-walk length) on C3 / C2 slices.
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift; timeout 600 python bench.py "$@" > $OUT/r01r_$tag.json 2> $OUT/r01r_$tag.err; python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r01r_{t}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(t, round(d["value"],2), {k:round(v,4) for k,v in r["per_kernel_avg_ms"].items()}, r.get("frac_by_sweep"), round(r["frac"],3))
except Exception as e: print(t,"failed",e)
PY
}
for c in 3 8 15; do ADFWI_B200_SHOTS_PER_CHUNK=$c run c3_ch$c --workload C3 --nt 400 --shots 15 --batch 15 --steps 2; done
for c in 0 5 10; do ADFWI_B200_SHOTS_PER_CHUNK=$c run c2_ch$c --workload C2 --steps 1; done
