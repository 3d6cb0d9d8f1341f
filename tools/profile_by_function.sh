#!/bin/bash
# per-function view of a capture (run HERE, with the build that was profiled still in build/obj):
#   tools/profile_by_function.sh <tag> <kernel> <EQ> <NS> <DERIV> <DAMP> <GRADS> <ODE 1|2>   -> profiles/r2_<tag>_by_function.txt
tag=$1; kern=$2; eq=$3; ns=$4; der=$5; dmp=$6; grd=$7; ode=$8
enc() { if [ "$1" -lt 0 ]; then echo "Lin${1#-}E"; else echo "Li${1}E"; fi; }
mangled="_ZN8rays_dev${#kern}${kern}INS_6TraitsI$(enc $eq)$(enc $ns)$(enc $der)$(enc $dmp)$(enc $grd)EEEEvNS_9TraceArgsE"
tmp=$(mktemp -d)
(cd $tmp && cuobjdump -xelf all /root/repo/build/obj/tu_${eq}_${ode}.o > /dev/null 2>&1)
nvdisasm -gi $tmp/*.cubin > $tmp/dis_all.txt 2> /dev/null
awk -v k=".text.$mangled:" 'BEGIN{on=0} { if ($0==k) {on=1; print; next} if (on && ($0 ~ /^\t\.section/ || $0 ~ /^\.text\./)) exit; if (on) print }' $tmp/dis_all.txt > $tmp/dis.txt
zcat gpurun_out/r2_${tag}_sass.csv.gz > $tmp/sass.csv
python tools/ncu_by_function.py $tmp/dis.txt $tmp/sass.csv "$mangled" | tee profiles/r2_${tag}_by_function.txt
cp $tmp/dis.txt /tmp/r2_${tag}_dis.txt; cp $tmp/sass.csv /tmp/r2_${tag}_sass.csv
rm -rf $tmp
