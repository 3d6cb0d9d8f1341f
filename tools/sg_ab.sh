#!/bin/bash
# A/B runs of the SG kernel (measurement aid): library variants x ray_deriv; SG parity tests first
python -m pytest tests -m gpu -x -q -k "sg or SG or figures or spline or eqdsk" 2>&1 | tail -3
run() {  # tag, env...
  tag=$1; shift
  for der in cold numerical; do
    env "$@" python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --ode SG_ODE --deriv $der --rays 262144 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('$tag deriv=$der', 'ray-steps/s %.3e'%d['value'], 'ms', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), 'ctas', d['roofline']['ctas_per_sm'], 'steps', d['config']['ray_steps_per_fan'])"
  done
}
run base X=1
for v in "$@"; do run $v RAYS_B200_LIB=$PWD/rays_b200/lib/librays_b200_$v.so; done
