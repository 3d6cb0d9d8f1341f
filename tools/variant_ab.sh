#!/bin/bash
# A/B of library variants on the RK4 bench workload (measurement aid): tools/variant_ab.sh tag1 tag2 ...
one() { env "$@" python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu $BARGS 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],4), 'resume', round(r['resume_pass_ms'],1))"; }
echo -n "base: "; one X=1
for v in "$@"; do echo -n "$v: "; one RAYS_B200_LIB=$PWD/rays_b200/lib/librays_b200_$v.so; done
echo -n "base again: "; one X=1
