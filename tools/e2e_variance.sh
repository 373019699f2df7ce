#!/bin/bash
# run-to-run spread of the end-to-end number of a short line, with the per-step host trace and the allocator's driver traffic
for i in 1 2 3 4 5 6; do
  ADFWI_BENCH_TRACE=1 python bench.py "$@" --warmup 3 --no-cpu-baseline 2> /tmp/e2e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f e2e %.1f ms/step %.1f e2e ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
  grep trace /tmp/e2e.err
done
