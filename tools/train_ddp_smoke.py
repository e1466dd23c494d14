"""Data-parallel training steps (one process per GPU, torchrun): every rank trains on its own seeded batch, gradients
are averaged per bucket over NCCL (backward.GradBuckets), and the ranks must hold bit-identical weights after each
step.  Three passes from the same initial weights: per-rank BatchNorm statistics, SyncBatchNorm over NCCL, and
SyncBatchNorm over the peer-memory exchange kernel (K10x, backward.PeerExchange) — with two ranks the last two must
agree to the bit (a two-term sum commutes); with more ranks the summation orders differ in the last fp64 bit and the
weight digests are compared to 1e-4.  Dev/verification tool for the
training row (§8f f4):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_ddp_smoke.py [size]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dram_b200  # noqa: E402,F401
from dram_b200 import med3d, training  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def run(mode, rank, world, dev):
    torch.manual_seed(0)  # identical initial weights on every rank
    model = med3d.resnet18segreg().to(dev).train()
    step = training.TrainStep(model, lr=1e-4, bucket_bytes=16 << 20, sync_bn=mode)
    g = torch.Generator().manual_seed(100 + rank)  # a different batch per rank
    lung = torch.zeros((1, S, S, S), dtype=torch.bool)
    lung[:, S // 8: -S // 8, S // 6: -S // 6, S // 8: -S // 8] = True
    batch = {"image": torch.randn((1, S, S, S), generator=g).to(dev), "lung_mask": lung.to(dev),
             "em_mask": ((torch.rand((1, S, S, S), generator=g) < 0.1) & lung).to(dev),
             "cls_label": torch.tensor([2 + rank % 2]).to(dev), "pse_label": torch.tensor([1]).to(dev)}
    bands = torch.tensor([[0.05, 0.1]]).to(dev), torch.tensor([[0.01, 0.05]]).to(dev)
    w = torch.ones(1, device=dev)
    digests = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step.step(batch, bands[0], bands[1], w, w)
        e1.record()
        torch.cuda.synchronize()
        if step.peer is not None:
            step.peer.check()
        # without SyncBatchNorm the running statistics are per rank; the parameters still have to agree
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()] +
                         ([b.detach().float().reshape(-1) for b in model.buffers()] if mode else []))
        digest = torch.stack([flat.double().sum(), flat.double().abs().sum()])
        all_d = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(all_d, digest)
        same = all(bool(torch.equal(all_d[0], d)) for d in all_d)
        losses = [torch.zeros_like(loss) for _ in range(world)]
        dist.all_gather(losses, loss)
        digests.append(digest.cpu())
        if rank == 0:
            print(f"sync_bn={mode!s:5} step {it}: losses {[round(float(l), 4) for l in losses]}  weights"
                  f"{'+buffers' if mode else ''} identical across {world} ranks: {same}  "
                  f"{e0.elapsed_time(e1):.1f} ms  buckets {step.buckets.num_buckets}", flush=True)
        assert same, "ranks diverged"
    if step.peer is not None:
        step.peer.close()
    return digests


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    run(False, rank, world, dev)
    d_nccl = run("nccl", rank, world, dev)
    d_peer = run("peer", rank, world, dev)
    equal = all(bool(torch.equal(a, b)) for a, b in zip(d_nccl, d_peer))
    close = all(bool(torch.allclose(a, b, rtol=1e-4)) for a, b in zip(d_nccl, d_peer))
    if rank == 0:
        print(f"SyncBatchNorm over NCCL vs peer-memory kernel after 4 steps: bit-identical {equal}, within 1e-4 {close}", flush=True)
    assert close and (equal or world > 2), "peer exchange disagrees with NCCL"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
