#!/bin/bash
# A/B runs of tools/profile_run.py (stage timers, 1080p) over experiment builds of the library.
# usage: tools/gpu_ab.sh <batch> <lib.so|default> ...
mkdir -p gpurun_out
B=$1; shift
for lib in "$@"; do
  echo "=== $lib (batch $B)" | tee -a gpurun_out/ab.log
  if [ "$lib" = default ]; then
    timeout 180 python tools/profile_run.py 3 $B 2>&1 | tail -2 | tee -a gpurun_out/ab.log
  else
    OFB_LIB=$PWD/$lib timeout 180 python tools/profile_run.py 3 $B 2>&1 | tail -2 | tee -a gpurun_out/ab.log
  fi
done
