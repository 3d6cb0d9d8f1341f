#!/bin/bash
# measurement aid: GPU tests, the headline kernel, and the copier kernel against the in-kernel copy-out on the 524k-ray fan
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-also --no-config5 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('headline ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],4))"
BARGS="--rays 524288" tools/copier_ab.sh RAYS_B200_COPIER=1 RAYS_B200_COPIER=0 RAYS_B200_COPIER=1
