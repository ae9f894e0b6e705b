#!/bin/bash
# Build A/B variants of the edge kernel: each argument is name:"nvcc -D flags" for csrc/rhs_window.cu; the other objects
# of the library are reused.  Results: graph-neural-pde_b200/build/variants/lib_<name>.so (git-ignored, shipped to the
# GPU box by gpurun).  Select one at run time with GRAND_LIB=/root/repo/graph-neural-pde_b200/build/variants/lib_<name>.so.
#   tools/build_variants.sh shfl:-DGRAND_SEGMAX_SHUFFLE=1 tb2:-DGRAND_EPI_TERMS_BATCH=2 "both:-DGRAND_SEGMAX_SHUFFLE=1 -DGRAND_EPI_TERMS_BATCH=2"
set -e
cd "$(dirname "$0")/../graph-neural-pde_b200"
python build_native.py > /dev/null
mkdir -p build/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -c csrc/rhs_window.cu -o build/variants/rw_$name.o
    objs=$(ls build/*.o | grep -v rhs_window.o)
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $objs build/variants/rw_$name.o
    rm -f build/variants/rw_$name.o
    echo "built $name ($flags)"
  ) &
done
wait
