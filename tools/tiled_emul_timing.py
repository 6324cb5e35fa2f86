"""One GPU: emulate `world` ranks of the tiled mode and print each rank's per-stage event times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import opticalflowcontainer_b200 as ofb
from opticalflowcontainer_b200 import tiled
from oracle import synth
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W, H = (7680, 4320) if len(sys.argv) < 3 else (int(sys.argv[2]), int(sys.argv[3]))
t = synth.cheap_texture(H, W, 400)
a = torch.from_numpy(t).cuda(); b = torch.from_numpy(synth.subpixel_shift(t, 9.5, -4.25)).cuda()
out = torch.zeros((H, W, 2), dtype=torch.float32, device="cuda")
engs = [ofb.FlowEngine(W, H, 1, 0) for _ in range(world)]
tiled.setup_local(engs)
for rep in range(3):
    if rep == 2:
        for e in engs: e.timing_enable(True)
    tiled.farneback_tiled_emulated(engs, a.data_ptr(), b.data_ptr(), W, H, W, out.data_ptr())
for r, e in enumerate(engs):
    st = e.timing_read()
    print("rank", r, {k: round(v[0], 3) for k, v in st.items()}, [round(x, 3) for x in e.timing_samples("iteration")])
