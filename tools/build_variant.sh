#!/bin/bash
# tools/build_variant.sh <name> <source.cu> <nvcc -D flags...>: an experimental libclipppo_b200 with ONE object rebuilt under extra
# defines, written to gpurun_exp/lib_<name>.so (travels with the gpurun snapshot; selected with CLIPPPO_LIB=<path>).
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p gpurun_exp/obj_$name
obj=gpurun_exp/obj_$name/${src%.cu}.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include "$@" -Xptxas=-v \
  -c clip-ppo_b200/csrc/$src -o $obj 2> gpurun_exp/obj_$name/ptxas.log
others=$(ls clip-ppo_b200/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o gpurun_exp/lib_$name.so $obj $others
echo gpurun_exp/lib_$name.so
