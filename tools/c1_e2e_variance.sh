#!/bin/bash
# C1 (whole-sweep kernels): run-to-run spread of the resident and the end-to-end number, with the per-step host trace
for i in 1 2 3 4 5 6; do
  ADFWI_BENCH_TRACE=1 python bench.py --workload C1 --steps 8 --warmup 3 --no-cpu-baseline 2> /tmp/c1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f e2e %.1f ms/step %.1f e2e ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
  grep trace /tmp/c1.err
done
