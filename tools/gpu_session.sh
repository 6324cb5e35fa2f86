#!/bin/bash
mkdir -p gpurun_out
V=opticalflowcontainer_b200/csrc/build/variants
{
echo "##### smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "##### bitwise: default vs no fused upsample"; timeout 300 python tools/compare_variants.py opticalflowcontainer_b200/libofb.so $V/libofb_noups.so 2>&1 | tail -4
} > gpurun_out/s_smoke.log 2>&1
cat gpurun_out/s_smoke.log
rm -f gpurun_out/ab.log
tools/gpu_ab.sh 18 default $V/libofb_noups.so $V/libofb_l2pf2.so $V/libofb_l2pf4.so $V/libofb_l2pf4noups.so default > /dev/null 2>&1
cat gpurun_out/ab.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/s_pytest.log 2>&1; tail -25 gpurun_out/s_pytest.log
timeout 900 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; tail -c 3000 gpurun_out/s_bench.json; tail -5 gpurun_out/s_bench.err
