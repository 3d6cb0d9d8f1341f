#!/bin/bash
# one kernel of tools/profile_all.sh (same recipe: plain run first, then ncu --set full of the same command, post-processed on the box):
#   tools/profile_one.sh damp | mirror | num | cold | sg2
mkdir -p gpurun_out
COMMON="--steps 1 --warmup 1 --no-e2e --no-cpu --no-also --no-config5"
prof() {  # tag kernel-regex count bench-args...
  tag=$1; kre=$2; cnt=$3; shift 3
  python bench.py $COMMON "$@" > gpurun_out/r2_plain_$tag.log 2>&1 || { echo "plain run failed: $tag"; tail -3 gpurun_out/r2_plain_$tag.log; return; }
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$kre -c $cnt -o gpurun_out/r2_prof_$tag python bench.py $COMMON "$@" > gpurun_out/r2_ncu_$tag.log 2>&1
  tail -n 1 gpurun_out/r2_ncu_$tag.log
}
post() { bash tools/profile_post.sh "$@"; }
case "$1" in
  num) prof trace_rk4_num trace_rk4_kernel 2; post trace_rk4_num trace_rk4_kernel 2 2 2 0 0 2 ;;
  cold) prof trace_rk4_cold trace_rk4_kernel 2 --deriv cold; post trace_rk4_cold trace_rk4_kernel 2 2 1 0 0 2 ;;
  mirror) prof trace_rk4_mirror trace_rk4_kernel 1 --workload mirror_fan_1M; post trace_rk4_mirror trace_rk4_kernel 4 2 1 0 1 1 ;;
  damp) prof trace_rk4_damp trace_rk4_kernel 1 --workload axisym_deposition_fan --config5-grid 1024; post trace_rk4_damp trace_rk4_kernel 3 2 1 1 0 1 ;;
  sg2) prof trace_sg2 trace_sg2_kernel 1 --ode SG_ODE --deriv cold --rays 262144; post trace_sg2 trace_sg2_kernel 2 2 1 0 0 1 ;;
esac
