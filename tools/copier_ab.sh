#!/bin/bash
# measurement aid: page-locked end-to-end path, copier kernel (default) against the in-kernel streaming copy-out (RAYS_B200_COPIER=0)
one() { env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-also --no-config5 --e2e-pinned-only $BARGS 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('device ms', round(d['ms_per_step'],2), '| e2e ms', round(e['ms_per_step'],2), 'GB/s', round(e['d2h_gbs_achieved_all_ranks'],2), 'dma', round(e['d2h_dma_gbs_all_ranks'],2), 'parity', (d.get('parity') or {}).get('bitwise'))"; }
for v in "$@"; do echo -n "$v: "; one $v; done
