#!/bin/bash
# shots per (tile, chunk) item of the reverse kernels: C4, C3 PML, C3 sponge
run() {
  python bench.py --nt 400 --shots 15 --batch 15 --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('  value %.1f  fwd %.3f adj %.3f whole %.3f' % (d['value'], r['frac_by_sweep']['forward_recording'], r['frac_by_sweep']['adjoint'], r['whole_step_frac']))"
}
for c in 0 3 4 5 8 15; do echo "C4 reverse chunk $c"; run --workload C4 --cfg shots_per_chunk_reverse=$c; done
for c in 0 5 8 15; do echo "C3 PML reverse chunk $c"; run --workload C3 --cfg shots_per_chunk_reverse=$c; done
for c in 0 5 8 15; do echo "C3 sponge reverse chunk $c"; run --workload C3 --abc gerjan --cfg shots_per_chunk_reverse=$c; done
for c in 2 3 4; do echo "C4 forward chunk $c (reverse default)"; run --workload C4 --cfg shots_per_chunk=$c --cfg shots_per_chunk_reverse=4; done
