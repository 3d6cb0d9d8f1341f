#!/bin/bash
# round-end check: GPU tests; the copy-out tests once more with kernel launches serialised (the copier kernel must hand over to the
# post-trace delivery instead of waiting for a trace kernel that cannot start); the default bench line
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CUDA_LAUNCH_BLOCKING=1 timeout 200 python -m pytest tests -m gpu -x -q -k "copy_out or smoke" 2>&1 | tail -3
( time timeout 600 python bench.py > gpurun_out/r2b_bench_default.log 2>&1 ) 2>&1 | grep real
python tools/show_bench.py gpurun_out/r2b_bench_default.log 2>/dev/null | head -30 || tail -c 600 gpurun_out/r2b_bench_default.log
