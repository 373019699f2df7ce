#!/bin/bash
# run-to-run variance of the short elastic lines: value, launches, per-kernel averages, memory at start
run() {
  nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader | tr '\n' ' '
  python bench.py --nt 400 --shots 15 --batch 15 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('  value %.1f e2e %.1f ms/step %.1f fwd %.3f adj %.3f whole %.3f launches %d avg_ms %s cfg %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['frac_by_sweep']['forward_recording'], r['frac_by_sweep']['adjoint'], r['whole_step_frac'], d['gpu_launches'], r.get('per_kernel_avg_ms'), {k: d['config'].get(k) for k in ('ckpt_interval','memory_scheme','history')}))"
}
for i in 1 2 3 4 5 6; do run --workload C4 --steps 3; done
for i in 1 2 3; do run --workload C3 --abc gerjan --steps 3; done
