#!/bin/bash
# C3 (split PML) and C4 lines only (A/B helper)
for w in C3 C4; do
python bench.py --workload $w --nt 400 --shots 15 --batch 15 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('  %s value %.1f  fwd %.3f adj %.3f whole %.3f' % (sys.argv[1], d['value'], r['frac_by_sweep']['forward_recording'], r['frac_by_sweep']['adjoint'], r['whole_step_frac']))" $w
done
