#!/bin/bash
# Build an A/B variant of the C-ABI library into variants/libemp_<name>.so with extra -D flags / sed edits.
# usage: scripts/build_variant.sh <name> '<sed expr on emp_logl.cuh>' [extra nvcc flags]
set -e
name=$1; expr=$2; shift 2
d=/tmp/emp_variant_$name; rm -rf $d; mkdir -p $d/astroemperor_b200/csrc $d/include
cp /root/repo/include/emperor_b200.h $d/include/
cp /root/repo/astroemperor_b200/csrc/*.cu /root/repo/astroemperor_b200/csrc/*.cuh /root/repo/astroemperor_b200/csrc/*.cpp $d/astroemperor_b200/csrc/
sed -i "$expr" $d/astroemperor_b200/csrc/emp_logl.cuh
mkdir -p /root/repo/variants
(cd $d/astroemperor_b200/csrc && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" -shared -o /root/repo/variants/libemp_$name.so emp_abi.cu emp_draws.cpp -lcudart 2>&1 | grep -A2 "logl_rv_kernel" | grep -E "spill|Used")
