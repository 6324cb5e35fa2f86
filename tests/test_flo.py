"""`.flo` writer/reader: byte layout as the reference writes it (pytorch_pwc_master/run.py:324-329)."""
import io
import os

import numpy as np
import pytest

from opticalflowcontainer_b200 import flo


def test_flo_layout_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    flow = rng.normal(size=(7, 11, 2)).astype(np.float32)
    p = os.path.join(tmp_path, "a.flo")
    flo.write_flo(p, flow)
    # the reference's writer, restated: three tofile() calls on a [2,H,W] tensor transposed to [H,W,2]
    ten = np.transpose(flow, (2, 0, 1))
    buf = io.BytesIO()
    buf.write(np.array([80, 73, 69, 72], np.uint8).tobytes())
    buf.write(np.array([ten.shape[2], ten.shape[1]], np.int32).tobytes())
    buf.write(np.array(ten.transpose(1, 2, 0), np.float32).tobytes())
    assert open(p, "rb").read() == buf.getvalue()
    assert np.array_equal(flo.read_flo(p), flow)


def test_flo_rejects_garbage(tmp_path):
    p = os.path.join(tmp_path, "b.flo")
    open(p, "wb").write(b"nope")
    with pytest.raises(ValueError):
        flo.read_flo(p)
    with pytest.raises(ValueError):
        flo.write_flo(p, np.zeros((3, 3), np.float32))
