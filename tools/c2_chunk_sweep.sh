#!/bin/bash
# acoustic C2 grid: shots per (tile, chunk) item -> per-item setup cost vs balance
for c in 0 1 2 3 4 5 10; do
  python bench.py --workload C2 --nt 400 --shots 10 --batch 10 --steps 3 --no-cpu-baseline --cfg shots_per_chunk=$c 2>/dev/null > /tmp/c2q.json
  python - "$c" <<'PY'
import json, sys
d = json.loads(open("/tmp/c2q.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("chunk", sys.argv[1], "value", round(d["value"], 1), {k: round(v, 3) for k, v in r["frac_by_sweep"].items()}, {k: round(v, 4) for k, v in r["per_kernel_avg_ms"].items()})
PY
done
