#!/bin/bash
# Experiment builds of libofb.so with -DOFB_DBG=<n> (ablation switches in the kernels).
# usage: tools/build_variant.sh <n> ...  -> opticalflowcontainer_b200/csrc/build/variants/libofb_dbg<n>.so
set -e
cd "$(dirname "$0")/../opticalflowcontainer_b200/csrc"
mkdir -p build/variants
for n in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -DOFB_DBG=$n \
       -c farneback.cu -o build/variants/farneback_dbg$n.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o build/variants/libofb_dbg$n.so \
       build/variants/farneback_dbg$n.o $(ls build/*.o | grep -v farneback.o)
  echo built build/variants/libofb_dbg$n.so
done
