#!/bin/bash
# A/B of the SG slot machine's scheduling policies (measurement aid): RAYS_B200_SG_MIXED = 0 | 1 | 2 on the cold 1M fan and the numerical 262k fan
one() { env "$1" timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-also --no-config5 --ode SG_ODE ${@:2} 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('$*', '| ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],1))"; }
python -m pytest tests -m gpu -x -q -k "slot_machine" 2>&1 | tail -2
for p in 0 1 2; do one RAYS_B200_SG_MIXED=$p --deriv cold; done
for p in 0 1 2; do one RAYS_B200_SG_MIXED=$p --deriv numerical --rays 262144; done
