#!/bin/bash
# A/B of experiment builds of the sparse path (tools/lk_profile.py host-to-host time + ncu durations of the tracker).
# usage: tools/lk_unroll_ab.sh default <lib.so> ...
mkdir -p gpurun_out
for lib in "$@"; do
  if [ "$lib" = default ]; then unset OFB_LIB; else export OFB_LIB=$PWD/$lib; fi
  echo "== $lib"; python tools/lk_profile.py 40 | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_lk_track -c 6 --csv python tools/lk_profile.py 4 2>/dev/null | grep k_lk_track | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
done
