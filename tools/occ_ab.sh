#!/bin/bash
# measurement aid: CTA size (registers per thread) of the synchronised damping kernel
one() { env "$@" python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-also --no-config5 $BARGS 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],4), 'resume', round(r['resume_pass_ms'],1))"; }
V=RAYS_B200_LIB=$PWD/rays_b200/lib/librays_b200_
BARGS="--workload axisym_deposition_fan"; echo -n "base: "; one X=1; for v in "$@"; do echo -n "$v: "; one ${V}$v.so; done
