#!/bin/bash
# O(2,6) split-PML on the C3 grid: fused reverse step elf_b<3> (one CTA per SM) against the pair elf_k1 + elf_k2
run() {
  python bench.py --workload C3 --order 6 --nt 400 --shots 15 --batch 15 --no-cpu-baseline --steps 2 2>/dev/null > /tmp/o26.json
  python - "$1" <<'PY'
import json, sys
d = json.loads(open("/tmp/o26.json").read().strip().splitlines()[-1]); r = d["roofline"]
print(sys.argv[1], "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), r["frac_by_sweep"], {k: round(v, 4) for k, v in r["per_kernel_avg_ms"].items()})
PY
}
run "fused elf_b<3>"
ADFWI_B200_EL_ADJ_SPLIT=1 run "pair elf_k1+elf_k2"
