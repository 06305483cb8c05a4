"""Time hb_spatial_gemm_3xtf32 at the C4 shape (64 x 64 antennas, T = 16 384 + delay) with CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hermespy_b200.kernels import spatial_gemm

B, n, T = int(os.environ.get("B", 128)), 64, 16384 + 50
S = torch.randn(B, n, n, dtype=torch.complex128, device="cuda")
z = torch.randn(B, n, T, dtype=torch.complex64, device="cuda")
y = torch.empty(B, n, T, dtype=torch.complex64, device="cuda")
for _ in range(3):
    spatial_gemm(S, z, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    spatial_gemm(S, z, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
byt = 2 * B * n * T * 8
print(f"spatial_gemm B={B}: {ms:.3f} ms, {byt / ms / 1e6:.0f} GB/s algorithmic, {B * T / ms / 1e6:.2f} G samples/s, "
      f"{3 * 8 * n * n * B * T / ms / 1e9:.0f} TF32 TFLOP/s issued")
ref = torch.matmul(S.to(torch.complex64), z)
print("rel err vs torch c64 matmul:", float((y - ref).norm() / ref.norm()))
