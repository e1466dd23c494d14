"""Times the backward kernels (K9 wgrad, dgrad on K1) on every conv shape of resnet34segreg at a cubic size
(SURVEY Appendix A.1).  Dev tool: per-layer TFLOP/s = algorithmic FLOPs (2*M*N*K of the forward conv) / CUDA-event
time, and the totals of one backward pass over the convolutions.
    python tools/backward_bench.py [size] [batch] [wgrad|dgrad|both] [name-substrings]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dram_b200  # noqa: E402,F401
from dram_b200 import backward  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
WHAT = sys.argv[3] if len(sys.argv) > 3 else "both"
ONLY = sys.argv[4].split(",") if len(sys.argv) > 4 else None
DT = torch.bfloat16
dev = torch.device("cuda:0")

# name, count, spatial divisor of the input, sources (channels), cout, kernel, stride, dil, needs dgrad
LAYERS = [
    ("layer1 3^3 64->64", 6, 4, (64,), 64, 3, 1, 1),
    ("layer2.0.conv1 s2", 1, 4, (64,), 128, 3, 2, 1),
    ("layer2 128->128", 7, 8, (128,), 128, 3, 1, 1),
    ("layer3.0.conv1 d2", 1, 8, (128,), 256, 3, 1, 2),
    ("layer3 256->256 d2", 11, 8, (256,), 256, 3, 1, 2),
    ("layer4.0.conv1 d4", 1, 8, (256,), 512, 3, 1, 4),
    ("layer4 512->512 d4", 5, 8, (512,), 512, 3, 1, 4),
    ("us1.0 576->64", 1, 4, (512, 64), 64, 3, 1, 1),
    ("us1.1 64->64", 1, 4, (64,), 64, 3, 1, 1),
    ("us2.0 128->64", 1, 2, (64, 64), 64, 3, 1, 1),
    ("us2.1 64->64", 1, 2, (64,), 64, 3, 1, 1),
    ("us3 64->32", 1, 2, (64,), 32, 3, 1, 1),
]


def timed(fn, flush, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = {"wgrad": [0.0, 0.0], "dgrad": [0.0, 0.0]}
    for name, count, div, srcs, cout, k, s, dl in LAYERS:
        if ONLY and not any(o in name for o in ONLY):
            continue
        dims = (S // div,) * 3
        pad = dl * (k - 1) // 2
        od = tuple((v + 2 * pad - dl * (k - 1) - 1) // s + 1 for v in dims)
        cin = sum(srcs)
        xs = [torch.randn((B,) + dims + (c,), device=dev).to(DT) for c in srcs]
        dy = torch.randn((B,) + od + (cout,), device=dev).to(DT)
        w = torch.randn((cout, cin, k, k, k), device=dev) * 0.02
        flops = 2.0 * B * od[0] * od[1] * od[2] * cout * cin * k ** 3
        line = f"{name:22s} x{count:2d}"
        if WHAT in ("wgrad", "both"):
            dw = torch.zeros((cout, cin, k, k, k), device=dev)
            plans, off = [], 0
            for x in xs:
                plans.append(backward.Conv3dWgradPlan(x, dy, dw=dw, kernel=k, stride=s, dilation=dl, cin_total=cin,
                                                      cin_offset=off))
                off += x.shape[4]
            ms = timed(lambda: [p.run() for p in plans], flush)
            tot["wgrad"][0] += ms * count
            tot["wgrad"][1] += flops * count
            line += f"  wgrad {ms:7.3f} ms {flops / ms / 1e9:7.1f} TFLOP/s (items {plans[0].items}, slices {plans[0].kslices}, N {plans[0].block_n})"
            del plans, dw
        if WHAT in ("dgrad", "both"):
            plans, off = [], 0
            for x in xs:
                c = x.shape[4]
                plans.append(backward.Conv3dDgradPlan(dy, w, dims, kernel=k, stride=s, dilation=dl,
                                                      cin_range=(off, off + c)))
                off += c
            ms = timed(lambda: [p.run() for p in plans], flush)
            tot["dgrad"][0] += ms * count
            tot["dgrad"][1] += flops * count
            line += f"  dgrad {ms:7.3f} ms {flops / ms / 1e9:7.1f} TFLOP/s ({plans[0].plan.algo})"
            del plans
        print(line, flush=True)
        del xs, dy
        torch.cuda.empty_cache()
    for kname, (ms, fl) in tot.items():
        if ms > 0:
            print(f"TOTAL {kname}: {ms:.3f} ms per {B} volume(s) of {S}^3, {fl / 1e12:.3f} TFLOP -> {fl / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
