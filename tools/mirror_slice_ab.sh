#!/bin/bash
# measurement aid: time slicing on the mirror fan, whose rays all run to nstep_max
one() { env "$@" python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-also --no-config5 --workload mirror_fan_1M 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],4), 'launches', r['launches_per_fan'], 'resume', round(r['resume_pass_ms'],1))"; }
for s in "" 0 250; do echo -n "slice '$s': "; if [ -n "$s" ]; then one RAYS_B200_SLICE=$s; else one X=1; fi; done
