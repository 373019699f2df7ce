#!/bin/bash
# C1 with the cluster-persistent kernels: chosen plan, and forced cluster sizes
run() {
  python bench.py --workload C1 --no-cpu-baseline --steps 3 2>/tmp/c1.err > /tmp/c1.json
  grep "persistent plan" /tmp/c1.err | sort | uniq | head -8
  python - "$1" <<'PY'
import json, sys
d = json.loads(open("/tmp/c1.json").read().strip().splitlines()[-1]); r = d["roofline"]
print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 2), r["per_kernel_avg_ms"], r["kernel_share_of_step"])
PY
}
ADFWI_B200_DEBUG=1 run "auto"
for nc in 6 7 8; do ADFWI_B200_PERSIST_NC=$nc run "NC=$nc"; done
ADFWI_B200_PERSIST=0 run "per-step"
