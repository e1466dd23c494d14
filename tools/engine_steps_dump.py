"""Runs N eager engine passes (no CUDA graph, no saturation probe) of one configuration and writes the recorded
launch sequence (name, algorithmic and executed FLOPs per conv launch) to JSON — the launch order an `ncu -k
regex:conv3d_` capture of the same command sees.  Dev tool for tools/gpu_profile_pipe.sh.

    python tools/engine_steps_dump.py OUT.json [arch] [D,H,W] [batch] [passes]
"""
import json
import os
import sys

import torch

os.environ["DRAM_B200_GRAPH"] = "0"
os.environ["DRAM_B200_SAT_CHECK"] = "off"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

out = sys.argv[1]
arch = sys.argv[2] if len(sys.argv) > 2 else "med3ddram"
dims = tuple(int(v) for v in sys.argv[3].split(",")) if len(sys.argv) > 3 else (256, 256, 256)
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
passes = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda:0")
module = bench.build_module(dev, arch)
hu, lungs, ess = bench.make_volumes(B, dims, dev, seed=0)
for _ in range(passes):
    module.predict_step_from_hu(hu, lungs, ess)
torch.cuda.synchronize()
eng = module.model.engine(B, dims, dev)
steps = [{"name": s.name, "flops": s.flops, "executed_flops": s.executed_flops} for s in eng.steps]
with open(out, "w") as f:
    json.dump({"arch": arch, "dims": dims, "batch": B, "passes": passes, "steps": steps}, f)
print("wrote", out, len(steps), "steps per pass")
