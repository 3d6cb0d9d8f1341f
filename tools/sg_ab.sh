#!/bin/bash
# A/B of the SG kernel's slot alignment (measurement aid; results are identical either way)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for al in 0 1; do
  for der in cold numerical; do
    RAYS_B200_SG_ALIGN=$al python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --ode SG_ODE --deriv $der --rays 262144 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('align=$al deriv=$der', 'ray-steps/s %.3e'%d['value'], 'ms', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), 'rays', d['config']['rays_per_gpu'], 'steps', d['config']['ray_steps_per_fan'])"
  done
done
