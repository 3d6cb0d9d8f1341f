#!/bin/bash
# ncu recipe used for profiles/ (run under gpurun on one B200; see /opt/skills/guides/B200_PROFILING.md):
#   launch list of the bench step, then one --set full capture of the trace kernel.  The plain command runs
#   first and must exit 0; numbers printed under ncu are never bench values.
set -x
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_rk4_kernel -c 1 -o gpurun_out/prof_trace $CMD > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu2.log
