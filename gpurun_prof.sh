set -x
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --ode SG_ODE --deriv cold --rays 65536"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_sg_kernel -c 1 -o gpurun_out/prof_sg_v1 $CMD > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu2.log
