#!/bin/bash
# A/B runs of tools/profile_run.py (stage timers, batch 8, 1080p) under different kernel switches.
# usage: tools/gpu_ab.sh "<ENV=.. ENV=..>" "<ENV..>" ...   (each arg = one configuration)
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "=== $cfg" | tee -a gpurun_out/ab.log
  env $cfg timeout 120 python tools/profile_run.py 3 8 2>&1 | tail -2 | tee -a gpurun_out/ab.log
done
