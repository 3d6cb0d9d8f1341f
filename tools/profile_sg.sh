#!/bin/bash
# ncu recipe for the Shampine-Gordon trace kernel (run under gpurun on one B200): plain run first, then one
# --set full capture of the first trace_sg_kernel launch.  Numbers printed under ncu are never bench values.
set -x
TAG=${1:-sg}
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --ode SG_ODE --deriv cold --rays 262144"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_sg_kernel -c 1 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -n 2 gpurun_out/plain_${TAG}.log
tail -n 3 gpurun_out/ncu_${TAG}.log
