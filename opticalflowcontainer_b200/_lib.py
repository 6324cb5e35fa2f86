"""ctypes binding of libofb.so (include/ofb.h).  Fails loudly: there is no CPU fallback.

This is the reference-side binding a maintainer would add (INTEGRATION.md shows it stand-alone);
the reference's own precedent for a native op is the pybind module
``ros2_ws/src/liteflownet3/correlation_package/correlation_cuda.cc:169-172``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFB_LIB") or os.path.join(_HERE, "libofb.so")   # OFB_LIB: experiment builds only


class OfbError(RuntimeError):
    """Raised for every non-zero ofb_status (cv2 would raise cv2.error); carries ``.status``."""

    def __init__(self, status: int, msg: str):
        super().__init__("libofb: %s (status %d)" % (msg, status))
        self.status = status


class JunctionParams(C.Structure):
    _fields_ = [("grid_area", C.c_int), ("grid_area_threshold", C.c_float), ("eps", C.c_int), ("dampen", C.c_int),
                ("dampen_min", C.c_double), ("dampen_max", C.c_double)]


class ClaheParams(C.Structure):
    _fields_ = [("adaptive", C.c_int), ("clip_limit", C.c_double), ("clip_min", C.c_double), ("clip_max", C.c_double),
                ("c_min", C.c_double), ("c_max", C.c_double), ("tiles_x", C.c_int), ("tiles_y", C.c_int)]


class FarnebackParams(C.Structure):
    _fields_ = [("pyr_scale", C.c_double), ("levels", C.c_int), ("winsize", C.c_int), ("iterations", C.c_int),
                ("poly_n", C.c_int), ("poly_sigma", C.c_double), ("flags", C.c_int)]


class LKParams(C.Structure):
    _fields_ = [("win_w", C.c_int), ("win_h", C.c_int), ("max_level", C.c_int), ("max_count", C.c_int),
                ("epsilon", C.c_double), ("flags", C.c_int), ("min_eig_threshold", C.c_double)]


class GfttParams(C.Structure):
    _fields_ = [("max_corners", C.c_int), ("quality_level", C.c_double), ("min_distance", C.c_double),
                ("block_size", C.c_int), ("use_harris_detector", C.c_int), ("harris_k", C.c_double)]


# every symbol include/ofb.h declares: name -> (restype, argtypes)
_u8p = C.c_void_p
_SIGNATURES = {
    "ofb_version": (C.c_int, []),
    "ofb_status_string": (C.c_char_p, [C.c_int]),
    "ofb_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ofb_destroy": (C.c_int, [C.c_void_p]),
    "ofb_last_error": (C.c_char_p, [C.c_void_p]),
    "ofb_stream": (C.c_void_p, [C.c_void_p]),
    "ofb_synchronize": (C.c_int, [C.c_void_p]),
    "ofb_farneback": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t,
                                C.POINTER(FarnebackParams)]),
    "ofb_farneback_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int,
                                      C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_size_t,
                                      C.POINTER(FarnebackParams)]),
    "ofb_farneback_batch_async": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int,
                                            C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_size_t,
                                            C.POINTER(FarnebackParams)]),
    "ofb_wait": (C.c_int, [C.c_void_p]),
    "ofb_farneback_batch_stats": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int,
                                            C.c_int, C.c_size_t, C.POINTER(FarnebackParams), C.c_void_p,
                                            C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "ofb_farneback_batch_stats_async": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                  C.c_int, C.c_int, C.c_size_t, C.POINTER(FarnebackParams), C.c_void_p,
                                                  C.c_void_p, C.c_void_p]),
    "ofb_flow_u_stats_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ofb_flow_postfilter": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_void_p), C.c_size_t, C.c_int]),
    "ofb_flow_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ofb_flow_to_bgr": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "ofb_flow_to_bgr_speed": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_size_t]),
    "ofb_flow_download": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_size_t]),
    "ofb_farneback_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                       C.c_size_t, C.c_void_p, C.POINTER(FarnebackParams)]),
    "ofb_farneback_sequence_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                                C.c_size_t, C.c_void_p, C.POINTER(FarnebackParams)]),
    "ofb_farneback_stream": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_size_t,
                                       C.POINTER(C.c_void_p), C.c_size_t, C.POINTER(FarnebackParams), C.POINTER(C.c_int)]),
    "ofb_farneback_stream_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                              C.c_void_p, C.POINTER(FarnebackParams), C.POINTER(C.c_int)]),
    "ofb_stream_reset": (C.c_int, [C.c_void_p]),
    "ofb_resize_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                C.c_size_t]),
    "ofb_ingest_gray": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_int,
                                  C.c_int, C.c_size_t]),
    "ofb_clahe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p,
                            C.c_size_t]),
    "ofb_adapt_prefilter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_double)]),
    "ofb_bilateral_u8c3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_double,
                                     C.c_void_p, C.c_size_t]),
    "ofb_jpeg_info": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ofb_jpeg_entropy_decode": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ofb_jpeg_set_host_entropy": (C.c_int, [C.c_void_p, C.c_int]),
    "ofb_jpeg_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "ofb_jpeg_decode_luma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "ofb_ingest_jpeg_gray": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]),
    "ofb_find_junctions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(JunctionParams),
                                     C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "ofb_junction_threshold": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                         C.POINTER(JunctionParams), C.c_void_p, C.c_size_t]),
    "ofb_cluster_junctions": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "ofb_launch_count": (C.c_uint64, [C.c_void_p]),
    "ofb_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "ofb_timing_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "ofb_timing_read_samples": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]),
    "ofb_tiled_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ofb_tiled_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ofb_tiled_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ofb_tiled_import_local": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ofb_farneback_tiled_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                             C.c_void_p, C.POINTER(FarnebackParams), C.POINTER(C.c_int),
                                             C.POINTER(C.c_int)]),
    "ofb_tiled_barrier": (C.c_int, [C.c_void_p]),
    "ofb_tiled_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "ofb_farneback_tiled_emulated": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                               C.c_size_t, C.c_void_p, C.POINTER(FarnebackParams)]),
    "ofb_flow_u_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "ofb_cvt_gray": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]),
    "ofb_cvt_gray_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p,
                                      C.c_size_t]),
    "ofb_good_features": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_size_t, C.POINTER(GfttParams),
                                    C.c_void_p, C.POINTER(C.c_int)]),
    "ofb_good_features_masked": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_size_t, _u8p, C.c_size_t,
                                           C.POINTER(GfttParams), C.c_void_p, C.POINTER(C.c_int)]),
    "ofb_lk_stream": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_size_t, C.POINTER(GfttParams), C.POINTER(LKParams),
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p,
                                C.POINTER(C.c_int)]),
    "ofb_lk_stream_reset": (C.c_int, [C.c_void_p]),
    "ofb_corner_min_eigenval": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p]),
    "ofb_lk_pyramid": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "ofb_pyrlk": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.POINTER(LKParams)]),
}

TILED_EXPORT_BYTES = 320   # OFB_TILED_EXPORT_BYTES

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load libofb.so (built in-tree by ``opticalflowcontainer_b200.build``).  Raises if it is
    missing — the product path never falls back to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libofb.so not found at %s — build it with `python -m opticalflowcontainer_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, handle=None):
    if status != 0:
        lib = load()
        msg = lib.ofb_last_error(handle)
        msg = msg.decode() if msg else ""
        if not msg:
            msg = lib.ofb_status_string(status).decode()
        raise OfbError(status, msg)
