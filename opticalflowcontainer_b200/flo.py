"""Middlebury ``.flo`` files — the format the reference writes its flow fields in
(``ros2_ws/src/pwc_net/pwc_net/pytorch_pwc_master/run.py:324-329``: bytes ``PIEH``, int32 width, int32 height,
then float32 ``[height][width][2]`` (u, v) rows).  cv2's ``[H,W,2]`` layout — what
``calcOpticalFlowFarneback`` and this package return — is exactly the payload, so a field is written as is."""
from __future__ import annotations

import numpy as np

_MAGIC = b"PIEH"   # 80, 73, 69, 72 — the float 202021.25 in little-endian


def write_flo(path: str, flow_hw2: np.ndarray) -> None:
    flow = np.ascontiguousarray(flow_hw2, dtype="<f4")
    if flow.ndim != 3 or flow.shape[2] != 2:
        raise ValueError("flow must be [H,W,2]")
    with open(path, "wb") as f:
        f.write(_MAGIC)
        np.array([flow.shape[1], flow.shape[0]], "<i4").tofile(f)
        flow.tofile(f)


def read_flo(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        if f.read(4) != _MAGIC:
            raise ValueError("%s is not a .flo file (bad magic)" % path)
        w, h = np.fromfile(f, "<i4", 2)
        if w <= 0 or h <= 0:
            raise ValueError("bad .flo size %dx%d" % (w, h))
        data = np.fromfile(f, "<f4", int(w) * int(h) * 2)
    if data.size != int(w) * int(h) * 2:
        raise ValueError("%s is truncated" % path)
    return data.reshape(int(h), int(w), 2)
