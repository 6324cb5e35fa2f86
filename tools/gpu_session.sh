#!/bin/bash
# One GPU session: smoke -> variant A/B -> GPU tests -> bench.  Everything under timeouts; logs in gpurun_out/.
mkdir -p gpurun_out
V=opticalflowcontainer_b200/csrc/build/variants
{
echo "##### smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "##### bitwise: default vs smem ring"; timeout 300 python tools/compare_variants.py opticalflowcontainer_b200/libofb.so $V/libofb_smemring.so 2>&1 | tail -4
echo "##### bitwise: default vs no fused upsample"; timeout 300 python tools/compare_variants.py opticalflowcontainer_b200/libofb.so $V/libofb_noups.so 2>&1 | tail -4
} > gpurun_out/s1_smoke.log 2>&1
cat gpurun_out/s1_smoke.log
rm -f gpurun_out/ab.log
tools/gpu_ab.sh 18 default $V/libofb_tmem2.so $V/libofb_tmem3.so $V/libofb_smemring.so $V/libofb_noups.so $V/libofb_smem_noups.so default > /dev/null 2>&1
cat gpurun_out/ab.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; tail -5 gpurun_out/s1_pytest.log
timeout 600 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; tail -c 1500 gpurun_out/s1_bench.json
