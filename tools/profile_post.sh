#!/bin/bash
# post-process one ncu report ON the GPU box (gpurun_out/ travels back only below 64 MiB, the reports do not fit):
#   tools/profile_post.sh <tag> <kernel> <EQ> <NS> <DERIV> <DAMP> <GRADS> <n-kernels> [keep]
# writes gpurun_out/r2_<tag>_ncu_summary.txt (tools/ncu_summary.py per captured launch) and r2_<tag>_sass.csv.gz (the ncu
# source page: per-instruction counters and stall samples) and removes the report unless `keep` is given.  The per-function
# table is made afterwards where the objects are: tools/profile_by_function.sh <tag> ... (same arguments), with the SAME build.
tag=$1; kern=$2; eq=$3; ns=$4; der=$5; dmp=$6; grd=$7; cnt=$8; keep=$9
rep=gpurun_out/r2_prof_$tag.ncu-rep
[ -f $rep ] || { echo "no report for $tag"; exit 0; }
enc() { if [ "$1" -lt 0 ]; then echo "Lin${1#-}E"; else echo "Li${1}E"; fi; }
mangled="_ZN8rays_dev${#kern}${kern}INS_6TraitsI$(enc $eq)$(enc $ns)$(enc $der)$(enc $dmp)$(enc $grd)EEEEvNS_9TraceArgsE"
: > gpurun_out/r2_${tag}_ncu_summary.txt
for ((k = 0; k < cnt; k++)); do
  echo "# captured launch $k" >> gpurun_out/r2_${tag}_ncu_summary.txt
  python tools/ncu_summary.py $rep $k >> gpurun_out/r2_${tag}_ncu_summary.txt 2>&1
done
tmp=$(mktemp -d)
ncu -i $rep --page source --csv --print-source sass --launch-count 1 > $tmp/sass.csv 2> $tmp/sass.err || ncu -i $rep --page source --csv --print-source sass > $tmp/sass.csv 2>> $tmp/sass.err
gzip -c $tmp/sass.csv > gpurun_out/r2_${tag}_sass.csv.gz
rm -rf $tmp
[ "$keep" = keep ] || rm -f $rep
