"""Is a training step reproducible?  Two copies of the network from the same state, the same batch, the same (native)
route, `reps` times: every weight gradient must come out bit-identical.  Prints, in backward order, the parameters
whose gradients differ between the copies / repetitions.  Dev tool (round 2: isolates kernel nondeterminism from the
native-vs-autograd route comparison of tests/test_training_gpu.py).
    python tools/train_repro_check.py [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dram_b200  # noqa: E402,F401
from dram_b200 import med3d, training  # noqa: E402
from oracle import training_oracle as T  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
case = T.train_case()
fix = torch.load(os.path.join(ROOT, "tests", "golden", "train_step_med3ddram18.pt"), weights_only=False)
batch = {k: case[k].to(dev) for k in ("image", "lung_mask", "em_mask", "cls_label", "pse_label")}
args = (fix["cle_bands"].to(dev), fix["pse_bands"].to(dev), case["cle_weights"].to(dev), case["pse_weights"].to(dev))
grads = []
for r in range(reps):
    model = med3d.resnet18segreg()
    model.load_state_dict(case["sd"])
    model = model.to(dev).train()
    step = training.TrainStep(model, lr=0.0)   # lr 0: the weights stay put, only the gradients matter
    loss = float(step.step(batch, *args))
    torch.cuda.synchronize()
    grads.append({n: step.buckets.view(n).detach().clone() for n, _ in model.named_parameters()})
    print(f"rep {r}: loss {loss!r}")
    del step, model
names = list(grads[0])
bad = 0
for n in reversed(names):
    for r in range(1, reps):
        a, b = grads[0][n], grads[r][n]
        if not torch.equal(a, b):
            rel = float((a - b).abs().max()) / (float(a.abs().max()) + 1e-30)
            print(f"  {n}: rep {r} differs from rep 0 by {rel:.3g} of the largest entry")
            bad += 1
            break
print("REPRODUCIBLE" if bad == 0 else f"NOT reproducible: {bad} of {len(names)} gradients differ")
