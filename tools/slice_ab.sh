#!/bin/bash
# slice-length A/B (measurement aid): RAYS_B200_SLICE forces suspend/resume every n steps on every pass
one() {  # label, bench args..., env in $E
  env $E python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu "${@:2}" 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$1', '$E', 'ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],1), 'frac', round(r['frac'],4), 'launches', r['launches_per_fan'], 'resume_ms', round(r['resume_pass_ms'],1))"
}
for s in "" 167 125 64; do E="X=1"; [ -n "$s" ] && E="RAYS_B200_SLICE=$s"; one rk4_num_1M; done
for s in "" 125 64 32; do E="X=1"; [ -n "$s" ] && E="RAYS_B200_SLICE=$s"; one sg_cold_1M --ode SG_ODE --deriv cold; done
