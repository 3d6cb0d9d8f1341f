#!/bin/bash
# measurement aid: the SG-related GPU tests + SG bench numbers (cold 262k / cold 1M) of the current build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sg or SG or slot or copy_out or smoke" 2>&1 | tail -4
one() { timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-also --no-config5 "$@" 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$*', '| ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],1), 'frac', round(r['frac'],4), r['kernel'], 'grid', r['grid'], 'ctas/sm', r['ctas_per_sm'])"; }
one --ode SG_ODE --deriv cold --rays 262144
one --ode SG_ODE --deriv cold
one --ode SG_ODE --deriv numerical --rays 262144
