#!/bin/bash
# Builds build/lib_default.so and one library per kernel experiment (tools/experiments/README.md) by recompiling
# only the two search translation units with the experiment's macro and relinking with the default objects.
# usage: tools/build_experiments.sh            (run from the repo root; needs nvcc, no GPU)
set -e
cd "$(dirname "$0")/.."
make -C diskann_b200/csrc -j8 > /dev/null
mkdir -p build /tmp/dab_exp
cp diskann_b200/libdiskann_b200.so build/lib_default.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall -cudart static"
cd diskann_b200/csrc
others=$(ls *.o | grep -v -e '^search_kernel_v2.o$' -e '^search_kernel.o$')
build_one() {  # name, macros...
  name=$1; shift
  $NVCC $FLAGS "$@" -c -o /tmp/dab_exp/${name}_v2.o search_kernel_v2.cu
  $NVCC $FLAGS "$@" -c -o /tmp/dab_exp/${name}_sk.o search_kernel.cu
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../build/lib_${name}.so $others /tmp/dab_exp/${name}_v2.o /tmp/dab_exp/${name}_sk.o
  echo "build/lib_${name}.so  ($*)"
}
build_one tag16 -DDAB_V2_TAG16_BUILD=1 &
build_one split -DDAB_V2_SPLIT_WAIT=1 &
build_one defer -DDAB_V2_DEFER_CAS=1 -DDAB_V2_MIN_CTAS=23 &
build_one wide -DDAB_V2_WIDE_LDS=1 &
build_one int -DDAB_V2_INT_BUILD=1 &
wait
echo "then: gpurun --timeout 900 -- 'for l in build/lib_*.so; do tools/quick_check.sh \$l; done; tools/sweep_libs.sh build/lib_*.so'"
