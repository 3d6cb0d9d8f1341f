#!/bin/bash
# round-end check: GPU tests, the default bench line, and the ncu capture of the config-5 kernel
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time timeout 600 python bench.py > gpurun_out/r2b_bench_default.log 2>&1 ) 2>&1 | grep real
tail -c 1500 gpurun_out/r2b_bench_default.log
tools/profile_one.sh damp
