"""Host->device copy bandwidth from pinned memory on this box, alone and while the conv kernels run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
n = 256 << 20
host = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(2):
    dst.copy_(host, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    dst.copy_(host, non_blocking=True)
b.record()
torch.cuda.synchronize()
print(f"H2D pinned, idle GPU: {5 * n / a.elapsed_time(b) / 1e6:.1f} GB/s")
side = torch.cuda.Stream()
x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
torch.cuda.synchronize()
with torch.cuda.stream(side):
    a.record(side)
    for _ in range(5):
        dst.copy_(host, non_blocking=True)
    b.record(side)
for _ in range(60):
    y = x @ x
torch.cuda.synchronize()
print(f"H2D pinned, GPU busy with GEMMs: {5 * n / a.elapsed_time(b) / 1e6:.1f} GB/s")
small = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
d2 = torch.empty_like(small, device=dev)
a.record()
for _ in range(5):
    d2.copy_(small, non_blocking=True)
b.record()
torch.cuda.synchronize()
print(f"H2D pinned fp32 256 MiB tensor: {5 * small.numel() * 4 / a.elapsed_time(b) / 1e6:.1f} GB/s")
