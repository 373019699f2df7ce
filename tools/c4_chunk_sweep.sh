#!/bin/bash
# C4 (VTI, split PML): costly-tiles-first item order on/off and shots per (tile, chunk) item; prints value + per-kernel fractions
run() {
  python bench.py --workload C4 --nt 400 --shots 15 --batch 15 --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('  value %.1f  fwd %.3f adj %.3f whole %.3f' % (d['value'], r['frac_by_sweep']['forward_recording'], r['frac_by_sweep']['adjoint'], r['whole_step_frac']))"
}
for ord in 0 1; do
  for ch in 0 1 2 3 5; do
    echo "order=$ord shots_per_chunk=$ch"
    ADFWI_B200_EL_ORDER=$ord run --cfg shots_per_chunk=$ch
  done
done
echo "C3 PML order 0/1"
for ord in 0 1; do
  ADFWI_B200_EL_ORDER=$ord python bench.py --workload C3 --nt 400 --shots 15 --batch 15 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('  value %.1f  fwd %.3f adj %.3f whole %.3f' % (d['value'], r['frac_by_sweep']['forward_recording'], r['frac_by_sweep']['adjoint'], r['whole_step_frac']))"
done
