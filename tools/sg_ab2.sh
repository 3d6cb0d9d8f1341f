#!/bin/bash
# A/B of SG slot-machine variants (measurement aid) on the 1M-ray Solov'ev fan, deriv_cold, tol 1e-6: tools/sg_ab2.sh tag...
one() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-also --no-config5 --ode SG_ODE --deriv cold 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$tag', '| ray-steps/s %.4e'%d['value'], 'ms', round(d['ms_per_step'],1), 'grid', r['grid'], 'ctas/sm', r['ctas_per_sm'])"; }
one base X=1
for v in "$@"; do
  case $v in
    slots=*) one $v RAYS_B200_SG_SLOTS=${v#slots=} ;;
    *) one $v RAYS_B200_LIB=$PWD/rays_b200/lib/librays_b200_$v.so ;;
  esac
done
