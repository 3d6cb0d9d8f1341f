#!/bin/bash
# measurement aid: a library variant that differs from the main build only in ONE trace translation unit
#   tools/build_tu_variant.sh <tag> <eq 1..4> <ode 1|2> <extra nvcc flags...>   ->  rays_b200/lib/librays_b200_<tag>.so   (use with RAYS_B200_LIB)
tag=$1; eq=$2; ode=$3; shift 3
O=build/obj_$tag; mkdir -p $O
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC "$@" -DRAYS_TU_EQ=$eq -DRAYS_TU_ODE=$ode -c rays_b200/csrc/trace_tu.cu -o $O/tu_${eq}_${ode}.o || exit 1
objs=$(ls build/obj/*.o | grep -v tu_${eq}_${ode}.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o rays_b200/lib/librays_b200_$tag.so $objs $O/tu_${eq}_${ode}.o -Xlinker -z -Xlinker defs -lpthread -ldl
