#!/bin/bash
# the elastic configurations at their full length (nt 4000, checkpointed history) and the C5 slice, for the record
for w in C3 C4; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02z_${w}_full.json 2> gpurun_out/bench_r02z_${w}_full.err
  python - $w <<'PY'
import json, sys
w = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_r02z_{w}_full.json").read().strip().splitlines()[-1]); r = d["roofline"]
print(w, "full length: value %.1f e2e %.1f shots/s %.2f ms/step %.0f" % (d["value"], d["e2e"]["value"], d.get("shots_per_s", 0), d["ms_per_step"]),
      {k: round(v, 3) for k, v in r["frac_by_sweep"].items()}, "whole", round(r["whole_step_frac"], 3), "ckpt row", round(r["whole_step_frac_checkpointed_row"], 3), d["config"].get("history"))
PY
done
