"""Spatially tiled mode (config 5): the strip-wise result must agree with the whole-frame result of the
same engine (same kernels; only the block alignment of the float vertical sums differs with the
segmentation, so agreement is ~1e-6 px, not bit-exact) and meet the cv2 gate.  One GPU: the ranks are
emulated in one process (stage by stage, no barrier kernel).  Two or more GPUs: real peer-memory
run in one process with the flag barrier."""
import numpy as np
import pytest

from oracle import cv2_oracle as C
from oracle import synth

pytestmark = pytest.mark.gpu


def _run_whole(eng, a, b, **kw):
    return eng.farneback(a, b, None, **kw)


@pytest.mark.parametrize("world,shape,kw", [
    (2, (256, 320), {}),
    (4, (480, 640), {}),
    (8, (481, 637), {}),
    (3, (300, 400), dict(winsize=9, iterations=2)),
    (2, (200, 300), dict(levels=1, poly_n=7, poly_sigma=1.5)),
])
def test_tiled_emulated_matches_whole_frame(built_lib, world, shape, kw):
    import torch
    import opticalflowcontainer_b200 as ofb
    from opticalflowcontainer_b200 import tiled

    h, w = shape
    a, b = synth.synth_pair(h, w, 7, (5.3, -3.7))
    whole = ofb.FlowEngine(w, h, 1, 0)
    ref = _run_whole(whole, a, b, **kw)
    whole.close()
    engs = [ofb.FlowEngine(w, h, 1, 0) for _ in range(world)]
    try:
        tiled.setup_local(engs)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        out = torch.full((h, w, 2), float("nan"), dtype=torch.float32, device="cuda")
        tiled.farneback_tiled_emulated(engs, da.data_ptr(), db.data_ptr(), w, h, w, out.data_ptr(), **kw)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
    finally:
        for e in engs:
            e.close()
    assert np.isfinite(got).all()
    # same kernels, different segment boundaries -> float summation order differs slightly
    assert float(np.abs(got - ref).max()) <= 1e-3, float(np.abs(got - ref).max())
    mean, mx = C.epe(C.farneback(a, b, **kw), got)
    assert mean <= 1e-3 and mx <= 1e-2, (mean, mx)


def test_tiled_rejects_unsupported(built_lib):
    import torch
    import opticalflowcontainer_b200 as ofb
    from opticalflowcontainer_b200 import tiled

    engs = [ofb.FlowEngine(64, 64, 1, 0) for _ in range(2)]
    try:
        tiled.setup_local(engs)
        z = torch.zeros((64, 64), dtype=torch.uint8, device="cuda")
        o = torch.zeros((64, 64, 2), dtype=torch.float32, device="cuda")
        with pytest.raises(ofb.OfbError):
            tiled.farneback_tiled_emulated(engs, z.data_ptr(), z.data_ptr(), 64, 64, 64, o.data_ptr(), flags=256)
    finally:
        for e in engs:
            e.close()


def test_tiled_two_gpus_peer_memory(built_lib):
    """Real run: one handle per GPU in this process, NVLink peer loads + flag barrier."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import opticalflowcontainer_b200 as ofb
    from opticalflowcontainer_b200 import tiled

    world = min(torch.cuda.device_count(), 4)
    h, w = 1080, 1920
    a, b = synth.synth_pair(h, w, 3, (6.2, 3.4))
    whole = ofb.FlowEngine(w, h, 1, 0)
    ref = whole.farneback(a, b)
    whole.close()
    engs = [ofb.FlowEngine(w, h, 1, r) for r in range(world)]
    try:
        tiled.setup_local(engs)
        ins, outs, rows = [], [], [None] * world
        for r in range(world):
            with torch.cuda.device(r):
                ins.append((torch.from_numpy(a).cuda(r), torch.from_numpy(b).cuda(r)))
                outs.append(torch.zeros((h, w, 2), dtype=torch.float32, device="cuda:%d" % r))
        for r in range(world):
            torch.cuda.synchronize(r)
        # the launches are asynchronous: one host thread can enqueue all ranks; the barrier kernels meet on the GPUs
        for rep in range(2):
            for r in range(world):
                rows[r] = tiled.farneback_tiled_device(engs[r], ins[r][0].data_ptr(), ins[r][1].data_ptr(), w, h, w,
                                                       outs[r].data_ptr())
            for r in range(world):
                assert not tiled.tiled_status(engs[r])
        got = np.empty((h, w, 2), np.float32)
        for r in range(world):
            yb, ye = rows[r]
            assert (yb, ye) == tiled.row_range(h, world, r)
            got[yb:ye] = outs[r][yb:ye].cpu().numpy()
    finally:
        for e in engs:
            e.close()
    assert float(np.abs(got - ref).max()) <= 1e-3, float(np.abs(got - ref).max())
