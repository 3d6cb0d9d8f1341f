#!/bin/bash
# measurement aid: all GPU tests + the default bench (every BASELINE config) of the current build
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_gputests.log 2>&1; tail -4 gpurun_out/r2_gputests.log
(time python bench.py) > gpurun_out/r2_bench_default.log 2>&1
python tools/show_bench.py gpurun_out/r2_bench_default.log | grep -E "^value|workload|  value|ms_per_fan|ms_per_step|frac:|bitwise|max_rel|ok:|kernel:" 
