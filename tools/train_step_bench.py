"""Times one training step (see training.py) of med3ddram (ResNet-34) — dev tool, NOT a bench.py number.
    python tools/train_step_bench.py [size] [batch] [arch] [loss: native|aten] [optimizer: native|torch]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dram_b200  # noqa: E402,F401
from dram_b200 import med3d, training  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ARCH = sys.argv[3] if len(sys.argv) > 3 else "resnet34segreg"
LOSS = sys.argv[4] if len(sys.argv) > 4 else "native"
OPT = sys.argv[5] if len(sys.argv) > 5 else "native"
dev = torch.device("cuda:0")


def main():
    torch.manual_seed(0)
    model = getattr(med3d, ARCH)().to(dev).train()
    step = training.TrainStep(model, lr=1e-5, loss=LOSS, optimizer=OPT)
    g = torch.Generator().manual_seed(1)
    lung = torch.zeros((B, S, S, S), dtype=torch.bool)
    lung[:, S // 8: -S // 8, S // 6: -S // 6, S // 8: -S // 8] = True
    batch = {"image": torch.randn((B, S, S, S), generator=g).to(dev), "lung_mask": lung.to(dev),
             "em_mask": (torch.rand((B, S, S, S), generator=g) < 0.1).to(dev) & lung.to(dev),
             "cls_label": torch.full((B,), 2).to(dev), "pse_label": torch.full((B,), 1).to(dev)}
    bands = torch.tensor([[0.05, 0.1]] * B).to(dev), torch.tensor([[0.01, 0.05]] * B).to(dev)
    w = torch.ones(B, device=dev)
    for i in range(3):
        t0 = time.time()
        loss = step.step(batch, bands[0], bands[1], w, w)
        torch.cuda.synchronize()
        print(f"warm-up {i}: loss {float(loss):.4f}  {1e3 * (time.time() - t0):.1f} ms wall", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        step.step(batch, bands[0], bands[1], w, w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{ARCH} {S}^3 batch {B} loss={LOSS} optimizer={OPT}: {ms:.2f} ms per training step (forward + loss + backward + Adam), "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step.step(batch, bands[0], bands[1], w, w)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
