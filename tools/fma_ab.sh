#!/bin/bash
# cost of the parity contract (VERDICT r1 missing #7): the default bench step with the FMA-contracted build
# (make -C rays_b200/csrc FMAD=true OBJ=../../build/obj_fma LIB=../lib/librays_b200_fma.so) next to the parity build.
# The JSON lines carry value, roofline and the parity block (vs the FMA-free oracle) of the headline fan and of every `also` workload.
mkdir -p gpurun_out
RAYS_B200_LIB=$PWD/rays_b200/lib/librays_b200_fma.so python bench.py --steps 2 --warmup 1 --no-config5 > gpurun_out/r2_bench_fma.log 2>&1
python tools/show_bench.py gpurun_out/r2_bench_fma.log | grep -E "workload|value|bitwise|max_rel|max_abs|ok:|frac:|ms_per" | head -80
