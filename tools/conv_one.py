"""Times one convolution shape (dev tool; also the target of single-kernel ncu captures).

    python tools/conv_one.py N D H W C1 C2 COUT K STRIDE DIL [residual_channels] [algo]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dram_b200  # noqa: E402,F401
from dram_b200 import ops  # noqa: E402

n, d, h, w, c1, c2, cout, k, s, dl = (int(v) for v in sys.argv[1:11])
res_c = int(sys.argv[11]) if len(sys.argv) > 11 else 0
algo = sys.argv[12] if len(sys.argv) > 12 else "auto"
dev, DT = torch.device("cuda:0"), torch.float16
x1 = torch.randn((n, d, h, w, c1), device=dev).to(DT)
x2 = torch.randn((n, d, h, w, c2), device=dev).to(DT) if c2 else None
wgt = (torch.randn(cout, k ** 3 * (c1 + c2), device=dev) * 0.02).to(DT)
bias = torch.zeros(cout, device=dev)
heads = None
if os.environ.get("HEADS") and cout == 32:  # us3: the two fused 1x1x1 heads + sigmoid, activation not stored
    heads = (torch.randn(2, 32, device=dev) * 0.1, torch.zeros(2, device=dev), (1, 1), True)
plan = ops.Conv3dPlan(x1, wgt, bias, x2=x2, kernel=k, stride=s, dilation=dl, algo=algo, heads=heads,
                      store_out=heads is None)
res = None
if res_c:
    res = torch.randn(plan.out_shape[:4] + (res_c,), device=dev).to(DT)
    plan = ops.Conv3dPlan(x1, wgt, bias, x2=x2, kernel=k, stride=s, dilation=dl, residual=res, algo=algo)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    plan.run()
ts = []
for _ in range(7):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    plan.run()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
vox_in, vox_out = n * d * h * w, plan.out_shape[0] * plan.out_shape[1] * plan.out_shape[2] * plan.out_shape[3]
nbytes = 2 * (vox_in * (c1 + c2) + vox_out * (cout + res_c)) + wgt.numel() * 2
print(f"conv {c1}+{c2}->{cout} k{k} s{s} d{dl} res{res_c} on {n}x{d}x{h}x{w} [{plan.algo}]: {ms:.3f} ms  "
      f"{plan.flops / ms / 1e9:.1f} TFLOP/s  {nbytes / ms / 1e6:.1f} GB/s of compulsory bytes")
