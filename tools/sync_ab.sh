#!/bin/bash
# measurement aid: the per-step CTA barrier (RAYS_RK4_SYNC) on the other RK4 kernel families
one() { env "$@" python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-also --no-config5 $BARGS 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],4), 'resume', round(r['resume_pass_ms'],1))"; }
V=RAYS_B200_LIB=$PWD/rays_b200/lib/librays_b200_
echo "mirror:"; BARGS="--workload mirror_fan_1M"; one X=1; one ${V}m256.so
echo "headline:"; BARGS=""; one X=1; one ${V}s128.so; one ${V}s384.so
echo "cold:"; BARGS="--deriv cold"; one X=1; one ${V}s128.so; one ${V}s256.so
