#!/bin/bash
# ncu captures behind profiles/r2_*: every trace kernel of DESIGN.md section 3 on its own workload (run under gpurun on one B200).
# Each command runs plain first and must exit 0; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
COMMON="--steps 1 --warmup 1 --no-e2e --no-cpu --no-also --no-config5"
prof() {  # tag kernel-regex count bench-args...
  tag=$1; kre=$2; cnt=$3; shift 3
  python bench.py $COMMON "$@" > gpurun_out/r2_plain_$tag.log 2>&1 || { echo "plain run failed: $tag"; tail -3 gpurun_out/r2_plain_$tag.log; return; }
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$kre -c $cnt -o gpurun_out/r2_prof_$tag python bench.py $COMMON "$@" > gpurun_out/r2_ncu_$tag.log 2>&1
  tail -n 1 gpurun_out/r2_ncu_$tag.log
}
post() { bash tools/profile_post.sh "$@"; }
prof trace_rk4_num trace_rk4_kernel 2                         # headline: first pass + resume pass over the suspended rays
post trace_rk4_num trace_rk4_kernel 2 2 2 0 0 2
prof trace_rk4_cold trace_rk4_kernel 2 --deriv cold
post trace_rk4_cold trace_rk4_kernel 2 2 1 0 0 2
prof trace_rk4_mirror trace_rk4_kernel 1 --workload mirror_fan_1M
post trace_rk4_mirror trace_rk4_kernel 4 2 1 0 1 1
prof trace_rk4_damp trace_rk4_kernel 1 --workload axisym_deposition_fan --config5-grid 1024
post trace_rk4_damp trace_rk4_kernel 3 2 1 1 0 1
prof fp64_peak fp64_peak_kernel 2 --rays 16384
python tools/ncu_summary.py gpurun_out/r2_prof_fp64_peak.ncu-rep 1 > gpurun_out/r2_fp64_peak_ncu_summary.txt 2>&1; rm -f gpurun_out/r2_prof_fp64_peak.ncu-rep
prof trace_sg2 trace_sg2_kernel 1 --ode SG_ODE --deriv cold --rays 262144
post trace_sg2 trace_sg2_kernel 2 2 1 0 0 1 keep
# launch list of the default bench step
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
tail -c 600 gpurun_out/r2_plain_bench.log
