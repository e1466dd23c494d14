"""Probe of the plane-ring kernel's UMMA descriptor addressing: runs one conv with the descriptor
base-offset field left 0 and set to (addr>>7)&7 and prints the error of each against F.conv3d."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dram_b200  # noqa: E402,F401
from dram_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for mode in ("0", "1"):
    os.environ["DRAM_B200_DESC_BASE_OFFSET"] = mode
    for dims, maxc in (((8, 16, 16), 0), ((12, 32, 24), 2)):
        x = torch.randn((1, 64) + dims, generator=g).half().float()
        w = (torch.randn(64, 64, 3, 3, 3, generator=g) * 0.03).half().float()
        b = torch.randn(64, generator=g) * 0.1
        ref = (F.conv3d(x, w, None, padding=1) + b.view(1, -1, 1, 1, 1)).relu()
        try:
            plan = ops.Conv3dPlan(ops.to_ndhwc_16(x.to(dev), torch.float16), ops.pack_conv_weight(w, dtype=torch.float16).to(dev),
                                  b.to(dev), algo="planes")
            got = ops.to_ncdhw_f32(plan.run(maxc)).cpu()
            torch.cuda.synchronize()
            err = (got - ref).abs()
            print(f"base_offset_mode={mode} dims={dims} max_ctas={maxc}: max err {err.max():.4f} mean {err.mean():.5f} "
                  f"ref absmax {ref.abs().max():.3f} frac>0.02 {(err > 0.02).float().mean():.4f}", flush=True)
            if err.max() > 0.05:
                bad = torch.nonzero(err > 0.05)
                print("   first bad idx (n,c,d,h,w):", bad[:5].tolist(), " bad w hist:",
                      torch.bincount(bad[:, 4], minlength=dims[2]).tolist())
        except Exception as e:  # noqa: BLE001
            print(f"base_offset_mode={mode} dims={dims}: FAILED {e}", flush=True)
            break
