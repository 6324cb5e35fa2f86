"""Spatially tiled Farneback: ONE frame pair split into row strips over the GPUs of a node
(BASELINE.json config 5).  Host-side set-up of ``ofb_tiled_*`` (include/ofb.h).

Two ways to bring the ranks together:

* one process per GPU (``torchrun``): :func:`setup_distributed` all-gathers the CUDA IPC blobs of the
  handles through ``torch.distributed`` (any backend; it is 320 bytes per rank) and imports them;
* all handles in one process (single-process multi-GPU, or the one-GPU emulation the tests use):
  :func:`setup_local`.

The data path has no collective: the kernels read the neighbours' rows through NVLink peer pointers
and a flag barrier in peer memory orders the stages.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

from . import _lib
from ._lib import OfbError, FarnebackParams
from .engine import FlowEngine


def row_range(height: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [begin, end) of a level of ``height`` rows that ``rank`` owns (ceil split, as the engine does)."""
    rpr = (height + world - 1) // world
    b = min(rank * rpr, height)
    return b, min(b + rpr, height)


def setup_local(engines: Sequence[FlowEngine]) -> None:
    """All ranks live in this process: rank i = engines[i]."""
    lib = _lib.load()
    world = len(engines)
    for r, e in enumerate(engines):
        _lib.check(lib.ofb_tiled_init(e._h, r, world), e._h)
    arr = (C.c_void_p * world)(*[e._h for e in engines])
    for e in engines:
        _lib.check(lib.ofb_tiled_import_local(e._h, arr), e._h)


def setup_distributed(engine: FlowEngine, rank: int, world: int) -> None:
    """One process per GPU: exchange the CUDA IPC handles through torch.distributed and import them."""
    import torch
    import torch.distributed as dist

    lib = _lib.load()
    _lib.check(lib.ofb_tiled_init(engine._h, rank, world), engine._h)
    blob = (C.c_ubyte * _lib.TILED_EXPORT_BYTES)()
    _lib.check(lib.ofb_tiled_export(engine._h, blob), engine._h)
    mine = bytes(blob)
    if world == 1:
        everyone = [mine]
    else:
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
    allb = b"".join(everyone)
    buf = (C.c_ubyte * len(allb)).from_buffer_copy(allb)
    _lib.check(lib.ofb_tiled_import(engine._h, buf), engine._h)
    if world > 1:
        dist.barrier()   # every rank has opened every handle before anyone launches
    del torch


def farneback_tiled_device(engine: FlowEngine, d_prev: int, d_next: int, width: int, height: int, pitch: int,
                           d_flow: int, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5,
                           poly_sigma=1.2, flags=0) -> Tuple[int, int]:
    """Enqueue this rank's part of one tiled pair; returns the rows [begin, end) of d_flow it writes."""
    lib = _lib.load()
    p = FarnebackParams(float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n), float(poly_sigma),
                        int(flags))
    b, e = C.c_int(0), C.c_int(0)
    _lib.check(lib.ofb_farneback_tiled_device(engine._h, d_prev, d_next, width, height, pitch, d_flow, C.byref(p),
                                              C.byref(b), C.byref(e)), engine._h)
    return b.value, e.value


def tiled_barrier(engine: FlowEngine) -> None:
    """Enqueue one cross-GPU flag barrier on the engine's stream (all ranks must call it)."""
    lib = _lib.load()
    _lib.check(lib.ofb_tiled_barrier(engine._h), engine._h)


def tiled_status(engine: FlowEngine) -> bool:
    """Synchronise; True if a cross-GPU barrier timed out since the last call."""
    lib = _lib.load()
    t = C.c_int(0)
    _lib.check(lib.ofb_tiled_status(engine._h, C.byref(t)), engine._h)
    return bool(t.value)


def farneback_tiled_emulated(engines: Sequence[FlowEngine], d_prev: int, d_next: int, width: int, height: int,
                             pitch: int, d_flow: int, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5,
                             poly_sigma=1.2, flags=0) -> None:
    """All ranks on ONE device in this process (tests): fills the whole of d_flow, synchronously."""
    lib = _lib.load()
    p = FarnebackParams(float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n), float(poly_sigma),
                        int(flags))
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    st = lib.ofb_farneback_tiled_emulated(arr, len(engines), d_prev, d_next, width, height, pitch, d_flow, C.byref(p))
    if st:
        for e in engines:
            msg = lib.ofb_last_error(e._h)
            if msg:
                raise OfbError(st, msg.decode())
        _lib.check(st, engines[0]._h)
