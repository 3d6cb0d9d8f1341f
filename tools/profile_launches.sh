#!/bin/bash
# launch list of a bench step without the `also` workloads (headline device-resident + the three end-to-end modes + config 5):
# plain run first, then the same command under ncu --metrics gpu__time_duration.sum.  ncu serialises kernels, so the copier
# kernel of the page-locked end-to-end path hands over to its post-trace delivery there (two copy_out_kernel launches per call).
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-also"
timeout 100 $CMD > gpurun_out/r2b_plain_bench_noalso.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/r2b_plain_bench_noalso.log; exit 1; }
timeout 160 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2b_launches_bench_noalso.csv $CMD > gpurun_out/r2b_ncu_launches.log 2>&1
tail -c 300 gpurun_out/r2b_ncu_launches.log; wc -l gpurun_out/r2b_launches_bench_noalso.csv
