#!/bin/bash
# quick before/after numbers (measurement aid): RK4 bench workload + SG cold/numerical on a 262k-ray fan + SG cold 1M
one() { python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$*', '| ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],1), 'frac', round(r['frac'],4), r['kernel'])"; }
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
one
one --ode SG_ODE --deriv cold --rays 262144
one --ode SG_ODE --deriv numerical --rays 262144
one --ode SG_ODE --deriv cold
