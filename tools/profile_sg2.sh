#!/bin/bash
# ncu capture of the Shampine-Gordon slot machine alone (262k-ray Solov'ev fan, deriv_cold, tol 1e-6): plain run, then --set full
mkdir -p gpurun_out
COMMON="--steps 1 --warmup 1 --no-e2e --no-cpu --no-also --no-config5 --ode SG_ODE --deriv cold --rays 262144"
python bench.py $COMMON > gpurun_out/r2_plain_trace_sg2.log 2>&1 || { tail -3 gpurun_out/r2_plain_trace_sg2.log; exit 1; }
tail -c 400 gpurun_out/r2_plain_trace_sg2.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_sg2_kernel -c 1 -f -o gpurun_out/r2_prof_trace_sg2 python bench.py $COMMON > gpurun_out/r2_ncu_trace_sg2.log 2>&1
bash tools/profile_post.sh trace_sg2 trace_sg2_kernel 2 2 1 0 0 1
