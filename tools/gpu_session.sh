#!/bin/bash
# One GPU session: smoke, bitwise A/B against experiment builds, stage timings, GPU tests, bench.
mkdir -p gpurun_out
V=opticalflowcontainer_b200/csrc/build/variants
{
echo "##### smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for o in $V/libofb_pf0.so; do
  [ -f $o ] && { echo "##### bitwise: default vs $o"; timeout 300 python tools/compare_variants.py opticalflowcontainer_b200/libofb.so $o 2>&1 | tail -4; }
done
} > gpurun_out/s_smoke.log 2>&1
cat gpurun_out/s_smoke.log
rm -f gpurun_out/ab.log
tools/gpu_ab.sh 18 default $(ls $V/*.so 2>/dev/null) default > /dev/null 2>&1
cat gpurun_out/ab.log
if [ "$1" != "quick" ]; then
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/s_pytest.log 2>&1; tail -15 gpurun_out/s_pytest.log
timeout 900 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; tail -c 1500 gpurun_out/s_bench.json; tail -5 gpurun_out/s_bench.err
fi
