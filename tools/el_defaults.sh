#!/bin/bash
# the elastic secondary lines with the planner's own choices
run() {
  python bench.py --nt 400 --shots 15 --batch 15 --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('  value %.1f  fwd %.3f adj %.3f whole %.3f' % (d['value'], r['frac_by_sweep']['forward_recording'], r['frac_by_sweep']['adjoint'], r['whole_step_frac']))"
}
echo "C4"; run --workload C4
echo "C3 PML"; run --workload C3
echo "C3 sponge"; run --workload C3 --abc gerjan
echo "C3 O(2,6)"; run --workload C3 --order 6
