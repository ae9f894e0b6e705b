#!/bin/bash
# Build A/B variants of one kernel source: each argument is name:"nvcc -D flags" for csrc/$SRC (default rhs_stream.cu); the
# other objects of the library are reused.  Results: graph-neural-pde_b200/build/variants/lib_<name>.so (git-ignored,
# shipped to the GPU box by gpurun).  Time them in one process with tools/edge_ab.py, or select one at run time with
# GRAND_LIB=/root/repo/graph-neural-pde_b200/build/variants/lib_<name>.so.
#   tools/build_variants.sh t448:-DGRAND_STREAM_THREADS=448 "t384b4:-DGRAND_STREAM_THREADS=384 -DGRAND_STREAM_BATCH=4"
#   SRC=rhs_ring.cu tools/build_variants.sh ring12:-DGRAND_RING_WARPS=12
set -e
cd "$(dirname "$0")/../graph-neural-pde_b200"
SRC=${SRC:-rhs_stream.cu}
BASE=${SRC%.cu}
python build_native.py > /dev/null
mkdir -p build/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -c csrc/$SRC -o build/variants/v_$name.o
    objs=$(ls build/*.o | grep -v "/$BASE.o")
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $objs build/variants/v_$name.o
    rm -f build/variants/v_$name.o
    echo "built $name ($flags)"
  ) &
done
wait
