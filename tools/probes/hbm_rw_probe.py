"""HBM read-only / write-only / copy bandwidth of this GPU (torch ops over a 4 GiB buffer, CUDA events, best of 10).
The roofline denominators in MEASURED_PEAKS.json come from a copy (read + write); a write-heavy kernel such as PolyExp
(1 B in, 20 B out per pixel) is bounded by what a WRITE stream sustains.  Usage: python tools/probes/hbm_rw_probe.py"""
import json, torch
n = 1 << 30
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def best(fn, reps=10):
    fn(); torch.cuda.synchronize()
    b = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        b = min(b, e0.elapsed_time(e1))
    return b
gb = n * 4 / 1e9
out = {"write_only_gbs": gb / (best(lambda: x.fill_(1.0)) * 1e-3),
       "read_only_gbs": gb / (best(lambda: x.sum()) * 1e-3),
       "copy_gbs_read_plus_write": 2 * gb / (best(lambda: y.copy_(x)) * 1e-3),
       # 1 part read : 4 parts written (u8 -> f32-like expansion), the PolyExp-like mix is closer to write-only
       "buffer_gib": 4}
print(json.dumps(out))
