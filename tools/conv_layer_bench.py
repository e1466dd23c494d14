"""Times K1 on every conv shape of resnet34segreg at a given cubic size.

Dev tool: prints per-layer TFLOP/s (algorithmic FLOPs / CUDA-event time) and the FLOP-weighted total.
    python tools/conv_layer_bench.py [size] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dram_b200  # noqa: E402,F401
from dram_b200 import ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ALGO = sys.argv[3] if len(sys.argv) > 3 else "auto"
ONLY = sys.argv[4].split(",") if len(sys.argv) > 4 else None  # substrings of layer names to keep
DT = torch.float16
dev = torch.device("cuda:0")

# name, count, spatial divisor of input, c1, c2, cout, kernel, stride, dil, tile
LAYERS = [
    ("stem(unfolded)", 1, (1, 2, 2), 64, 0, 64, (7, 1, 1), (2, 1, 1), 1, (16, 8, 1)),
    ("layer1 3^3 64->64", 6, 4, 64, 0, 64, 3, 1, 1, None),
    ("layer2.0.conv1 s2", 1, 4, 64, 0, 128, 3, 2, 1, None),
    ("layer2 128->128", 7, 8, 128, 0, 128, 3, 1, 1, None),
    ("layer3.0.conv1 d2", 1, 8, 128, 0, 256, 3, 1, 2, None),
    ("layer3 256->256 d2", 11, 8, 256, 0, 256, 3, 1, 2, None),
    ("layer4.0.conv1 d4", 1, 8, 256, 0, 512, 3, 1, 4, None),
    ("layer4 512->512 d4", 5, 8, 512, 0, 512, 3, 1, 4, None),
    ("us1.0 576->64", 1, 4, 512, 64, 64, 3, 1, 1, None),
    ("us1.1 64->64", 1, 4, 64, 0, 64, 3, 1, 1, None),
    ("us2.0 128->64", 1, 2, 64, 64, 64, 3, 1, 1, None),
    ("us2.1 64->64", 1, 2, 64, 0, 64, 3, 1, 1, None),
    ("us3 64->32 +heads", 1, 2, 64, 0, 32, 3, 1, 1, None),
]


def main():
    tot_t = tot_f = 0.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, count, div, c1, c2, cout, k, s, dl, tile in LAYERS:
        if ONLY and not any(o in name for o in ONLY):
            continue
        dv = (div, div, div) if isinstance(div, int) else div
        dims = tuple(S // q for q in dv)
        k3 = (k, k, k) if isinstance(k, int) else k
        x1 = torch.randn((B,) + dims + (c1,), device=dev).to(DT)
        x2 = torch.randn((B,) + dims + (c2,), device=dev).to(DT) if c2 else None
        taps = k3[0] * k3[1] * k3[2]
        w = (torch.randn(cout, taps * (c1 + c2), device=dev) * 0.02).to(DT)
        b = torch.zeros(cout, device=dev)
        heads = None
        if cout == 32:
            heads = (torch.randn(2, 32, device=dev), torch.zeros(2, device=dev), (1, 1), True)
        pad = (3, 0, 0) if k3 == (7, 1, 1) else None
        plan = ops.Conv3dPlan(x1, w, b, x2=x2, kernel=k3, stride=s, dilation=dl, padding=pad, heads=heads,
                              store_out=heads is None, tile=tile, algo=ALGO if tile is None else "tiles")
        for _ in range(2):
            plan.run()
        times = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        t = sorted(times)[len(times) // 2]
        fl = plan.flops
        if k3 == (7, 1, 1):
            fl = fl * 343 // 448
        print(f"{name:22s} x{count:2d} tiles {plan.m_tiles:6d}x{plan.n_tiles} bn {plan.block_n:3d} "
              f"{t:8.3f} ms {fl / t / 1e9:8.1f} TFLOP/s  (alg {fl / 1e9:8.1f} GF) {plan.algo}", flush=True)
        tot_t += t * count
        tot_f += fl * count
        del plan, x1, x2, w
    print(f"TOTAL conv: {tot_t:.3f} ms per {B} volume(s) of {S}^3, {tot_f / 1e12:.3f} TFLOP -> "
          f"{tot_f / tot_t / 1e9:.1f} TFLOP/s, {B / tot_t * 1e3:.2f} vol/s (conv only)")


if __name__ == "__main__":
    main()
