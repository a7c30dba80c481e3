#!/bin/bash
# builds a variant of libswb200.so with extra -D flags for ONE translation unit (A/B kernel experiments):
#   tools/build_variant.sh <name> <file.cu> [-DFOO=1 ...]   ->  gpurun_out/variants/libswb200_<name>.so   (use with SWB_LIB=...)
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/simpleworks_b200/_build/variants; mkdir -p $out
obj=$out/${src%.cu}_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fopenmp --expt-relaxed-constexpr -ccbin /usr/bin/g++ -Xptxas=-v "$@" -c $root/simpleworks_b200/csrc/$src -o $obj 2>&1 | grep -E "registers|spill" | head -${SWB_VERBOSE_LINES:-6}
objs=$(ls $root/simpleworks_b200/_build/*.o | grep -v "/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -o $out/libswb200_$name.so $objs $obj -lcudart_static -ldl -lrt -lpthread -lgomp
echo $out/libswb200_$name.so
