"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every
symbol include/ofb.h declares; the ctypes structs match the C layouts; errors are loud."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ofb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ofb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libofb.so does not export %s" % n


def test_binding_covers_header(built_lib):
    from opticalflowcontainer_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    lib = _lib.load()
    assert lib.ofb_version() == 100
    assert lib.ofb_status_string(3).decode().startswith("no usable CUDA device")


def test_struct_layouts():
    from opticalflowcontainer_b200._lib import FarnebackParams, LKParams, GfttParams
    assert C.sizeof(FarnebackParams) == 40
    assert FarnebackParams.poly_sigma.offset == 24 and FarnebackParams.flags.offset == 32
    assert C.sizeof(LKParams) == 40
    assert C.sizeof(GfttParams) == 40 and GfttParams.use_harris_detector.offset == 28 and GfttParams.harris_k.offset == 32


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the engine must raise, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import opticalflowcontainer_b200 as ofb
    with pytest.raises(ofb.OfbError) as ei:
        ofb.FlowEngine(64, 64)
    assert ei.value.status == 3
    import numpy as np
    with pytest.raises(ofb.OfbError):
        ofb.calcOpticalFlowFarneback(np.zeros((64, 64), np.uint8), np.zeros((64, 64), np.uint8), None, 0.5, 3, 15, 3,
                                     5, 1.2, 0)


def test_create_argument_errors(built_lib):
    from opticalflowcontainer_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.ofb_create(0, 0, 10, 1, C.byref(h)) == 1
    assert b"max_width" in lib.ofb_last_error(None)
    assert lib.ofb_destroy(None) == 0


def test_product_does_not_import_oracle():
    """The product package must never import or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "opticalflowcontainer_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|cv2)\b", s, flags=re.M), f
