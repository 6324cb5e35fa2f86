#!/bin/bash
# The Farneback GPU tests against a debug build of libofb.so with in-kernel bounds asserts (-DOFB_DBG=1: ring slots,
# staging buffers, output rows / columns of k_iter_v).  Build first (no GPU needed):
#   tools/build_variant.sh dbg "-DOFB_DBG=1"   (plus iter_fixed_a.cu / iter_fixed_b.cu with the same flag for the other window sizes)
# then on the GPU box:  bash tools/debug_asserts.sh   -> gpurun_out/debug_asserts.log
# A failing check aborts the kernel ("device-side assert triggered", with file and line on stderr).
mkdir -p gpurun_out
L=opticalflowcontainer_b200/csrc/build/variants/libofb_dbg.so
[ -f $L ] || { echo "build the debug variant first"; exit 1; }
OFB_LIB=$PWD/$L timeout 1200 python -m pytest tests/test_farneback_gpu.py tests/test_tiled_gpu.py -m gpu -q -x > gpurun_out/debug_asserts.log 2>&1
tail -4 gpurun_out/debug_asserts.log
echo "lines mentioning an assert: $(grep -c "assert" gpurun_out/debug_asserts.log)"
