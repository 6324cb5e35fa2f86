"""One process, one handle per GPU (setup_local, real peer memory + flag barrier): per-rank stage times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import opticalflowcontainer_b200 as ofb
from opticalflowcontainer_b200 import tiled
from oracle import synth
world = torch.cuda.device_count()
W, H = 7680, 4320
t = synth.cheap_texture(H, W, 400)
a = t; b = synth.subpixel_shift(t, 9.5, -4.25)
engs = [ofb.FlowEngine(W, H, 1, r) for r in range(world)]
tiled.setup_local(engs)
ins, outs = [], []
for r in range(world):
    ins.append((torch.from_numpy(a).cuda(r), torch.from_numpy(b).cuda(r)))
    outs.append(torch.zeros((H, W, 2), dtype=torch.float32, device="cuda:%d" % r))
for rep in range(4):
    if rep == 3:
        for e in engs: e.timing_enable(True)
    for r in range(world):
        tiled.farneback_tiled_device(engs[r], ins[r][0].data_ptr(), ins[r][1].data_ptr(), W, H, W, outs[r].data_ptr())
    for r in range(world):
        assert not tiled.tiled_status(engs[r])
for r, e in enumerate(engs):
    st = e.timing_read()
    print("rank", r, {k: round(v[0], 3) for k, v in st.items()}, [round(x, 3) for x in e.timing_samples("iteration")][-3:])
