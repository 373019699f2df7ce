#!/bin/bash
# quick C2-grid checks: vp-only and vp+rho gradients (nt 400 x 10 shots)
for extra in "" "--rho-grad"; do
  python bench.py --workload C2 --nt 400 --shots 10 --batch 10 --steps 3 --no-cpu-baseline $extra 2>/dev/null > /tmp/c2q.json
  python - "$extra" <<'PY'
import json, sys
d = json.loads(open("/tmp/c2q.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("C2", sys.argv[1] or "vp-only", "value", round(d["value"], 1), r["frac_by_sweep"], {k: round(v, 4) for k, v in r["per_kernel_avg_ms"].items()})
PY
done
