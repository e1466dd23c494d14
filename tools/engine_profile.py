"""Per-launch timing of one engine step (CUDA events around every recorded launch).

    python tools/engine_profile.py [arch] [D,H,W] [batch]
Prints name, ms, TFLOP/s for the conv launches and a summary; dev tool for finding the slow layers of a config.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "med3ddram"
dims = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (256, 256, 256)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
module = bench.build_module(dev, arch)
hu, lungs, ess = bench.make_volumes(B, dims, dev, seed=0)
for _ in range(2):
    module.predict_step_from_hu(hu, lungs, ess)
eng = module.model.engine(B, dims, dev)
torch.cuda.synchronize()
reps = 5
tot = [0.0] * len(eng.steps)
for _ in range(reps):
    evs = []
    for s in eng.steps:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s.fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        tot[i] += a.elapsed_time(b) / reps
conv_ms = conv_fl = all_ms = 0.0
for s, ms in zip(eng.steps, tot):
    all_ms += ms
    if s.flops:
        conv_ms += ms
        conv_fl += s.flops
        print(f"{s.name:28s} {ms:8.3f} ms {s.flops / ms / 1e9:8.1f} TFLOP/s  ({s.flops / 1e9:9.1f} GF)")
    else:
        print(f"{s.name:28s} {ms:8.3f} ms")
print(f"engine step {all_ms:.3f} ms, conv {conv_ms:.3f} ms = {conv_fl / conv_ms / 1e9:.1f} TFLOP/s over {conv_fl / 1e12:.3f} TFLOP")
