"""GPU box diagnostic: where does a device-resident step spend its time (host launch path vs kernels)?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hermespy_b200 import _lib
from hermespy_b200.batch import sample_fading_links
from hermespy_b200.kernels import FadingBatch, fading_propagate

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T, n = 15344, 4
ch = bench.make_channel(42)
blk = sample_fading_links(ch, B, n, n, 30.72e6)
fb = FadingBatch.from_numpy(device="cuda:0", **blk)
x = torch.view_as_complex(torch.randn((B, n, T, 2), device="cuda", dtype=torch.float32))
y = torch.empty((B, n, T + blk["max_delay"]), dtype=torch.complex64, device="cuda")
for _ in range(5):
    fading_propagate(x, fb, out=y)
torch.cuda.synchronize()
for label, prof in (("plain", False), ("profiled", True)):
    if prof:
        _lib.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(50):
        fading_propagate(x, fb, out=y)
    t_launch = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    rep = _lib.profile_end() if prof else None
    print(label, "host launch ms/step", 1e3 * t_launch / 50, "wall ms/step", 1e3 * t_all / 50, "event ms/step", e0.elapsed_time(e1) / 50)
    if rep:
        print({k: v["ms"] / max(1, v["launches"]) for k, v in rep.items() if v["launches"]})
