#!/bin/bash
# the sponge line only (A/B helper)
python bench.py --workload C3 --abc gerjan --nt 400 --shots 15 --batch 15 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('  sponge value %.1f  fwd %.3f adj %.3f whole %.3f' % (d['value'], r['frac_by_sweep']['forward_recording'], r['frac_by_sweep']['adjoint'], r['whole_step_frac']))"
