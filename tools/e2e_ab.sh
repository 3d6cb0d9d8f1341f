#!/bin/bash
# A/B of library variants on the page-locked end-to-end path (measurement aid): tools/e2e_ab.sh tag1 tag2 ...   (two rounds, interleaved:
# the host side of a shared box is noisy, see the DMA figure measured in the same run)
one() { env "$@" python bench.py --steps 3 --warmup 2 --no-cpu --no-also --no-config5 --e2e-pinned-only $BARGS 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('device ms', round(d['ms_per_step'],2), '| e2e ms', round(e['ms_per_step'],2), 'GB/s', round(e['d2h_gbs_achieved_all_ranks'],2), 'dma', round(e['d2h_dma_gbs_all_ranks'],2), 'ratio', round(e['d2h_gbs_achieved_all_ranks']/e['d2h_dma_gbs_all_ranks'],3))"; }
for rep in 1 2; do for v in "$@"; do echo -n "$v: "; one RAYS_B200_LIB=$PWD/rays_b200/lib/librays_b200_$v.so; done; done
