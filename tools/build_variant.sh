#!/bin/bash
# Experiment builds of libofb.so: the shipped library has one code path; kernel experiments are compile-time flags.
# usage: tools/build_variant.sh <name> "<nvcc -D flags>" [<name> "<flags>" ...]
#   -> opticalflowcontainer_b200/csrc/build/variants/libofb_<name>.so   (run with OFB_LIB=<that path>)
# flags: -DOFB_EXP_TMEM=false  -DOFB_EXP_NBUF=2|3  -DOFB_EXP_FUSE_UPS=false  (farneback.cu)
set -e
cd "$(dirname "$0")/../opticalflowcontainer_b200/csrc"
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden $flags \
       -c farneback.cu -o build/variants/farneback_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o build/variants/libofb_$name.so \
       build/variants/farneback_$name.o $(ls build/*.o | grep -v "/farneback.o")
  echo built build/variants/libofb_$name.so
done
