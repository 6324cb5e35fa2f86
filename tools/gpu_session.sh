#!/bin/bash
mkdir -p gpurun_out
V=opticalflowcontainer_b200/csrc/build/variants
{
echo "##### smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "##### bitwise: default vs no fused upsample"; timeout 300 python tools/compare_variants.py opticalflowcontainer_b200/libofb.so $V/libofb_noups.so 2>&1 | tail -4
echo "##### hbm probe"; timeout 120 python tools/probes/hbm_rw_probe.py 2>&1 | tail -1
} > gpurun_out/s_smoke.log 2>&1
cat gpurun_out/s_smoke.log
rm -f gpurun_out/ab.log
tools/gpu_ab.sh 18 default $V/libofb_noups.so default > /dev/null 2>&1
cat gpurun_out/ab.log
# ncu: one level-0 launch of the iteration kernel (12 launches per call; launches 9..11 are level 0)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_iter_v --launch-skip 21 --launch-count 1 \
  -o gpurun_out/r2_iter_v -f python tools/profile_run.py 3 18 > gpurun_out/ncu_iter.log 2>&1
tail -3 gpurun_out/ncu_iter.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polyexp_march --launch-skip 7 --launch-count 1 \
  -o gpurun_out/r2_polyexp -f python tools/profile_run.py 3 18 > gpurun_out/ncu_px.log 2>&1
tail -3 gpurun_out/ncu_px.log
