#!/bin/bash
# measurement aid: device-resident numbers of every bench workload (no e2e, no CPU leg), 3 steps each
one() { python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-also --no-config5 "$@" 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$*', '| ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],4), r['kernel'])"; }
one
one --deriv cold
one --workload mirror_fan_1M
one --workload axisym_deposition_fan
one --ode SG_ODE --deriv cold
one --ode SG_ODE --deriv numerical --rays 262144
