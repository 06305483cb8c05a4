"""GPU box: sweep of the host-buffer pipeline's chunk size (hb_fading_propagate_host, config C2, 512 links) against the
raw concurrent H2D + D2H copy rate of the box -- shows the end-to-end number of bench.py is PCIe-bound."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from hermespy_b200.batch import sample_fading_links
from hermespy_b200.kernels import fading_propagate_host
B, T, n = 512, 15344, 4
blk = sample_fading_links(bench.make_channel(42), B, n, n, 30.72e6)
D = blk["max_delay"]
xh = torch.empty((B, n, T), dtype=torch.complex128).pin_memory(); xh.normal_()
yh = torch.empty((B, n, T + D), dtype=torch.complex128).pin_memory()
for chunk in (0, 8, 16, 24, 48, 96, 171, 256):
    for _ in range(2): fading_propagate_host(xh.numpy(), out=yh.numpy(), precision="f32", chunk_links=chunk, **blk)
    t0 = time.perf_counter()
    for _ in range(5): fading_propagate_host(xh.numpy(), out=yh.numpy(), precision="f32", chunk_links=chunk, **blk)
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk {chunk:4d}: {dt*1e3:7.2f} ms  {B*T/dt/1e9:.3f} G samples/s  {(xh.numel()+yh.numel())*16/dt/1e9:.1f} GB/s both ways")
# complex64 host buffers
x32 = torch.empty((B, n, T), dtype=torch.complex64).pin_memory(); x32.normal_()
y32 = torch.empty((B, n, T + D), dtype=torch.complex64).pin_memory()
for chunk in (0, 48, 171):
    for _ in range(2): fading_propagate_host(x32.numpy(), out=y32.numpy(), precision="f32", chunk_links=chunk, **blk)
    t0 = time.perf_counter()
    for _ in range(5): fading_propagate_host(x32.numpy(), out=y32.numpy(), precision="f32", chunk_links=chunk, **blk)
    dt = (time.perf_counter() - t0) / 5
    print(f"c64 chunk {chunk:4d}: {dt*1e3:7.2f} ms  {B*T/dt/1e9:.3f} G samples/s  {(x32.numel()+y32.numel())*8/dt/1e9:.1f} GB/s both ways")
# raw copy ceiling
xd = torch.empty_like(xh, device="cuda"); xd2 = torch.empty_like(xh, device="cuda")
yc = torch.empty_like(xh).pin_memory()  # contiguous pinned target (a strided slice of yh would not be a DMA copy)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2): yc.copy_(xd2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"raw concurrent H2D+D2H of {xh.numel()*16/1e6:.0f} MB each: {dt*1e3:.2f} ms -> {2*xh.numel()*16/dt/1e9:.1f} GB/s both ways")
