"""TEST INFRASTRUCTURE (a checker script, not collected by pytest; it lives under tests/ because it runs the oracle).
Dense-map / score error of the B200 path against the CPU oracle for bf16 vs fp16 storage.
    python tests/precision_report.py [--cases small|all]
Prints max / p99.9 / mean abs error of the sigmoid dense maps (what becomes the dRAM) and the
relative error of the pooled scores, per architecture and storage type."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dram_b200  # noqa: E402,F401
from dram_b200 import med3d  # noqa: E402
from oracle import med3d_oracle as M  # noqa: E402
from oracle import synthetic  # noqa: E402

FACTORY = {"med3d": "resnet34segcls", "med3d18": "resnet18segcls", "med3d50": "resnet50segcls",
           "med3ddram": "resnet34segreg", "med3ddram18": "resnet18segreg", "med3ddram50": "resnet50segreg"}
CASES = [("med3ddram18", (32, 32, 32)), ("med3ddram", (32, 40, 48)), ("med3ddram50", (32, 32, 32)),
         ("med3ddram", (64, 64, 64)), ("med3ddram50", (48, 48, 48)), ("med3d18", (64, 64, 64))]
if "--cases" in sys.argv and sys.argv[sys.argv.index("--cases") + 1] == "all":
    CASES += [("med3ddram", (128, 128, 128)), ("med3ddram18", (96, 112, 144))]

dev = torch.device("cuda:0")
for arch, dims in CASES:
    sd = synthetic.make_state_dict(arch, seed=0, calib_dims=dims)
    x, lung, _ = synthetic.make_network_input(0, dims)
    x, lungs = x[None, None], lung[None, None].float()
    d_ref, s_ref = M.forward(sd, arch, x, lungs)
    for dt in (torch.bfloat16, torch.float16):
        model = getattr(med3d, FACTORY[arch])()
        model.load_state_dict(sd)
        model.act_dtype = dt
        model = model.to(dev).eval()
        dense, scores = model(x.to(dev), lungs.to(dev))
        parts = []
        for k in (0, 1):
            err = (dense[k].cpu() - d_ref[k]).abs().flatten()
            kk = max(1, int(err.numel() * 0.999))
            rel = ((scores[k].cpu() - s_ref[k]).abs() / s_ref[k].abs().clamp_min(1e-3)).max().item()
            parts.append(f"map{k}: max {err.max():.4f} p99.9 {err.kthvalue(kk).values:.4f} mean {err.mean():.5f} "
                         f"score rel {rel:.2e}")
        print(f"{arch:12s} {str(dims):16s} {str(dt)[6:]:9s} ref map std {d_ref[0].std():.3f} | " + " | ".join(parts),
              flush=True)
        del model
