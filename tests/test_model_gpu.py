"""Whole-network parity on the B200: dram_b200.med3d (CUDA kernels through the C-ABI) against
(a) golden vectors produced by the unmodified reference and (b) the CPU oracle on the same seeded
inputs/weights.  Tolerances are the north-star's: regression scores / logits within 1e-2 relative,
sigmoid dense maps (what becomes the dRAM) within 2e-2 max-abs, argmax classes identical."""
import glob
import os

import pytest
import torch

from oracle import med3d_oracle as M
from oracle import synthetic

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORWARD_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "forward_*.pt")))
FACTORY = {"med3d": "resnet34segcls", "med3d18": "resnet18segcls", "med3d50": "resnet50segcls",
           "med3ddram": "resnet34segreg", "med3ddram18": "resnet18segreg", "med3ddram50": "resnet50segreg"}


def build_model(arch, sd, device, dtype=None):
    from dram_b200 import med3d

    model = getattr(med3d, FACTORY[arch])()
    model.load_state_dict(sd)
    model.act_dtype = dtype  # None -> the library default (fp16 storage)
    return model.to(device).eval()


def report(got, ref):
    err = (got - ref).abs().flatten()
    k = max(1, int(err.numel() * 0.999))
    return (f"max {err.max().item():.4g} p99.9 {err.kthvalue(k).values.item():.4g} mean {err.mean().item():.4g} "
            f"(ref absmax {ref.abs().max().item():.4g}, ref std {ref.std().item():.4g})")


def check_outputs(arch, dense, scores, dense_ref, scores_ref):
    head = M.ARCHS[arch][2]
    for k, (got, ref) in enumerate(zip(dense, dense_ref)):
        got = got.float().cpu()
        assert got.shape == ref.shape
        msg = f"{arch} dense[{k}]: " + report(got, ref)
        print(msg)
        if head == "reg":
            assert (got - ref).abs().max().item() <= 2e-2, msg
        else:
            assert (got - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item()), msg
    for k, (got, ref) in enumerate(zip(scores, scores_ref)):
        got = got.float().cpu()
        assert got.shape == ref.shape, (got.shape, ref.shape)
        # regression scores: plain relative error.  Class logits: error relative to the magnitude of the
        # sample's logit vector (a logit that happens to be ~0 has no meaningful relative error of its own).
        den = ref.abs().clamp_min(1e-3) if head == "reg" else ref.abs().amax(-1, keepdim=True).clamp_min(0.25)
        rel = ((got - ref).abs() / den).max().item()
        print(f"{arch} score[{k}]: got {got.flatten().tolist()[:6]} ref {ref.flatten().tolist()[:6]} rel {rel:.3g}")
        assert rel <= 1e-2, f"{arch} score[{k}] rel err {rel}"
        if head == "cls":
            assert torch.equal(got.argmax(-1), ref.argmax(-1)), f"{arch}: argmax class differs"


@pytest.mark.parametrize("path", FORWARD_FIXTURES, ids=[os.path.basename(p)[:-3] for p in FORWARD_FIXTURES])
def test_forward_matches_reference_golden(cuda, lib, path):
    fix = torch.load(path)
    arch, dims, batch = fix["arch"], tuple(fix["dims"]), fix["batch"]
    sd = synthetic.make_state_dict(arch, seed=fix["weight_seed"], calib_dims=dims)
    model = build_model(arch, sd, cuda)
    xs, ls, _ = zip(*[synthetic.make_network_input(i, dims) for i in range(batch)])
    x = torch.stack(xs).unsqueeze(1).to(cuda)
    lungs = torch.stack(ls).unsqueeze(1).float().to(cuda) if fix["with_lungs"] else None
    dense, scores = model(x, lungs)
    torch.cuda.synchronize()
    check_outputs(arch, dense, scores, fix["dense_outs"], fix["scores"])


@pytest.mark.parametrize("arch,dims,batch", [("med3ddram", (64, 64, 64), 2), ("med3ddram18", (48, 56, 72), 1),
                                             ("med3d18", (64, 64, 64), 1), ("med3ddram50", (48, 48, 48), 1)])
def test_forward_matches_oracle(cuda, lib, arch, dims, batch):
    sd = synthetic.make_state_dict(arch, seed=2, calib_dims=dims)
    model = build_model(arch, sd, cuda)
    xs, ls, _ = zip(*[synthetic.make_network_input(10 + i, dims) for i in range(batch)])
    x = torch.stack(xs).unsqueeze(1)
    lungs = torch.stack(ls).unsqueeze(1).float()
    d_ref, s_ref = M.forward(sd, arch, x, lungs)
    dense, scores = model(x.to(cuda), lungs.to(cuda))
    check_outputs(arch, dense, scores, d_ref, s_ref)
    # uint8 mask gives the same pooled scores as the float mask
    if M.ARCHS[arch][2] == "reg":
        _, scores_u8 = model(x.to(cuda), lungs.to(torch.uint8).to(cuda))
        for a, b in zip(scores, scores_u8):
            assert torch.equal(a, b)


def test_weights_are_repacked_after_load_state_dict(cuda, lib):
    arch, dims = "med3ddram18", (32, 32, 32)
    sd_a = synthetic.make_state_dict(arch, seed=4, calib_dims=dims)
    sd_b = synthetic.make_state_dict(arch, seed=5, calib_dims=dims)
    model = build_model(arch, sd_a, cuda)
    x, lung, _ = synthetic.make_network_input(0, dims)
    x = x[None, None].to(cuda)
    _, s_a = model(x, None)
    model.load_state_dict(sd_b)
    d_b, s_b = model(x, None)
    d_ref, s_ref = M.forward(sd_b, arch, x.cpu(), None)
    assert not torch.equal(s_a[0], s_b[0])
    check_outputs(arch, d_b, s_b, d_ref, s_ref)


def test_forward_rejects_unsupported_use(cuda, lib):
    from dram_b200 import med3d

    model = med3d.resnet18segreg()
    with pytest.raises(RuntimeError):
        model.eval()(torch.zeros(1, 1, 32, 32, 32))           # CPU tensor: no fallback
    model = model.to(cuda)
    with pytest.raises(RuntimeError):
        model.train()(torch.zeros(1, 1, 32, 32, 32, device=cuda))  # training mode
    with pytest.raises(ValueError):
        # 36 -> 18 -> 9 -> 5: the up-sampled map (10) is larger than its skip tensor (9); the reference's torch.cat
        # fails for such sizes too
        model.eval()(torch.zeros(1, 1, 36, 32, 32, device=cuda))
