#!/bin/bash
# measurement aid: a library variant that differs from the main build only in the Solov'ev Shampine-Gordon translation unit
#   tools/build_sg_variant.sh <tag> <extra nvcc flags...>   ->  rays_b200/lib/librays_b200_<tag>.so   (use with RAYS_B200_LIB)
tag=$1; shift
O=build/obj_$tag; mkdir -p $O
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC "$@" -DRAYS_TU_EQ=2 -DRAYS_TU_ODE=2 -c rays_b200/csrc/trace_tu.cu -o $O/tu_2_2.o || exit 1
objs=$(ls build/obj/*.o | grep -v tu_2_2.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o rays_b200/lib/librays_b200_$tag.so $objs $O/tu_2_2.o -Xlinker -z -Xlinker defs -lpthread -ldl
