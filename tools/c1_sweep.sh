#!/bin/bash
# C1 (small grid) sweep over the shots one CTA walks through per tile; prints one line per setting
for c in 0 1 2 3 4 5 8; do
  ADFWI_B200_SHOTS_PER_CHUNK=$c python bench.py --workload C1 --no-cpu-baseline --steps 3 2>/dev/null > /tmp/c1_$c.json
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
d = json.loads(open(f"/tmp/c1_{c}.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("chunk", c, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 2), r["per_kernel_avg_ms"], r["kernel_share_of_step"])
PY
done
