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


REFINIT_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "refinit_*.pt")))


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
        # the plain element-wise relative error, reported beside it (DESIGN.md section 6 states the deviation: a class
        # logit that happens to lie near zero has an unbounded element-wise relative error at any precision)
        plain = ((got - ref).abs() / ref.abs().clamp_min(1e-6)).max().item()
        print(f"{arch} score[{k}]: got {got.flatten().tolist()[:6]} ref {ref.flatten().tolist()[:6]} rel {rel:.3g} "
              f"(element-wise relative {plain:.3g})")
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


@pytest.mark.parametrize("path", REFINIT_FIXTURES, ids=[os.path.basename(p)[:-3] for p in REFINIT_FIXTURES])
def test_reference_init_weights_match_reference_golden(cuda, lib, path, monkeypatch):
    """north_star: "random-init weights otherwise" — the reference's own initialisation (med3d.py:334-339), rebuilt
    from the seed (same random stream as the reference, pinned on CPU by test_oracle_cpu), with the fp16 saturation
    probe on for every pass.  This initialisation gives class logits / pre-sigmoid values of +-15 ... +-160, so the
    sigmoid maps sit at 0/1 almost everywhere and a voxel near a zero crossing turns any 16-bit rounding into an O(1)
    map difference; the errors are therefore asserted on the logit scale (2e-2 of the largest |logit|, the same 2e-2
    as the dRAM bar) and the sigmoid-map statistics are printed."""
    from dram_b200 import utils

    monkeypatch.setenv("DRAM_B200_SAT_CHECK", "always")
    fix = torch.load(path)
    arch, dims = fix["arch"], tuple(fix["dims"])
    torch.manual_seed(fix["init_seed"])
    model = utils.get_model_by_name(arch)
    sd = model.state_dict()
    assert abs(synthetic.state_dict_checksum(sd) - fix["weight_checksum"]) <= 1e-9 * abs(fix["weight_checksum"])
    model = model.to(cuda).eval()
    x, lung, _ = synthetic.make_network_input(0, dims)
    dense, scores = model(x[None, None].to(cuda), lung[None, None].float().to(cuda))
    eng = model.engine(1, dims, cuda)
    assert eng.act_dtype == torch.float16 and eng.last_saturation_count == 0   # nothing was clamped to +-65504
    head = M.ARCHS[arch][2]
    for k, (got, ref, logit_ref) in enumerate(zip(dense, fix["dense_outs"], fix["logits"])):
        got = got.float().cpu()
        scale = logit_ref.abs().max().item()
        if head == "cls":
            err = (got - ref).abs()
        else:  # compare on the logit scale; fp32 sigmoids saturate near |logit| = 17, so both sides are clamped at
            # logit(1e-6) = -13.8 / logit(1 - 1e-6) = +13.8 first (beyond that the maps are 0 / 1 to six digits anyway)
            eps = 1e-6
            err = (torch.logit(got.double().clamp(eps, 1 - eps)) - torch.logit(ref.double().clamp(eps, 1 - eps))).abs().float()
        print(f"{arch} refinit dense[{k}]: logit absmax {scale:.4g}, logit err max {err.max().item():.4g} "
              f"mean {err.mean().item():.4g} (= {err.max().item() / scale:.3g} of the scale); sigmoid/out map " + report(got, ref))
        assert err.max().item() <= 2e-2 * scale, (arch, k, err.max().item(), scale)
    for k, (got, ref) in enumerate(zip(scores, fix["scores"])):
        got = got.float().cpu()
        if head == "cls":
            den = ref.abs().amax(-1, keepdim=True)
            assert torch.equal(got.argmax(-1), ref.argmax(-1)), f"{arch}: argmax class differs"
        else:
            den = ref.abs().clamp_min(1e-3)
        rel = ((got - ref).abs() / den).max().item()
        print(f"{arch} refinit score[{k}]: got {got.flatten().tolist()} ref {ref.flatten().tolist()} rel {rel:.3g}")
        assert rel <= 1e-2, (arch, k, rel)


def test_fp16_saturation_is_detected_not_silent(cuda, lib, monkeypatch):
    """fp16 storage clamps at +-65504 in the epilogue (cvt.rn.satfinite); the probe turns that into an error on the
    first forward after a weight change, bf16 storage runs the same weights, and the check is repeated after
    load_state_dict."""
    from dram_b200 import ops

    arch, dims = "med3ddram18", (32, 32, 32)
    sd = synthetic.make_state_dict(arch, seed=3, calib_dims=dims)
    hot = dict(sd)
    hot["layer2.0.bn1.weight"] = sd["layer2.0.bn1.weight"] * 3.0e5   # pushes layer2.0.conv1's output beyond 65504
    x, lung, _ = synthetic.make_network_input(0, dims)
    x = x[None, None].to(cuda)
    model = build_model(arch, hot, cuda, torch.float16)
    with pytest.raises(ops.ActivationOverflow):
        model(x, None)
    with pytest.raises(ops.ActivationOverflow):   # still unchecked-clean: the next call probes again
        model(x, None)
    model.act_dtype = torch.bfloat16               # the documented way out
    d_bf, _ = model(x, None)
    assert all(bool(torch.isfinite(d).all()) for d in d_bf)
    model.act_dtype = torch.float16
    model.load_state_dict(sd)                      # sane weights: the same engine re-packs, re-probes, passes
    d_ref, s_ref = M.forward(sd, arch, x.cpu(), None)
    dense, scores = model(x, None)
    check_outputs(arch, dense, scores, d_ref, s_ref)
    eng = model.engine(1, dims, cuda)
    assert eng.last_saturation_count == 0
    monkeypatch.setenv("DRAM_B200_SAT_CHECK", "off")
    model.load_state_dict(hot)
    model(x, None)                                 # probe off: clamps silently (round-1 behaviour), by explicit choice


def test_cuda_graph_replay_equals_eager_launches(cuda, lib, monkeypatch):
    """The engine replays its launch sequence as one CUDA graph after the first (probed, eager) pass: results are
    bit-identical to eager launches, follow a new input and a load_state_dict, and DRAM_B200_GRAPH=0 keeps eager."""
    arch, dims = "med3ddram18", (32, 40, 48)
    sd = synthetic.make_state_dict(arch, seed=8, calib_dims=dims)
    xs = [synthetic.make_network_input(30 + i, dims)[0][None, None].to(cuda) for i in range(2)]
    lung = synthetic.make_network_input(30, dims)[1][None, None].to(cuda)
    monkeypatch.setenv("DRAM_B200_GRAPH", "0")
    eager = build_model(arch, sd, cuda)
    want = [[t.clone() for t in sum(eager(x, lung), [])] for x in xs]
    assert eager.engine(1, dims, cuda)._graph is None
    monkeypatch.setenv("DRAM_B200_GRAPH", "1")
    model = build_model(arch, sd, cuda)
    for rep in range(3):
        for x, w in zip(xs, want):
            got = sum(model(x, lung), [])
            assert all(torch.equal(a, b) for a, b in zip(got, w)), rep
    assert model.engine(1, dims, cuda)._graph is not None
    sd2 = synthetic.make_state_dict(arch, seed=9, calib_dims=dims)
    model.load_state_dict(sd2)
    eager.load_state_dict(sd2)
    for x in xs:
        assert all(torch.equal(a, b) for a, b in zip(sum(model(x, lung), []), sum(eager(x, lung), [])))


def test_stem_reads_the_callers_image_in_place(cuda, lib):
    """`engine.image_stem` (used by predict_step): the stem convolution launched on the batch's own fp32 tensor gives
    the bits of the route that first copies the image into the engine's buffer — also from a view at an odd offset
    (the stem's generic load path) and with the rest of the network replayed as a graph."""
    arch, dims = "med3ddram18", (32, 40, 48)
    model = build_model(arch, synthetic.make_state_dict(arch, seed=12, calib_dims=dims), cuda)
    eng = model.engine(2, dims, cuda)
    flat = torch.zeros(2 * 32 * 40 * 48 + 3, device=cuda)
    for rep in range(2):
        img = torch.stack([synthetic.make_network_input(50 + 2 * rep + i, dims)[0] for i in range(2)]).to(cuda)
        eng.load_image(img)
        want = [t.clone() for t in eng.run_network()]
        eng.image.zero_()
        got = [t.clone() for t in eng.run_network(first=eng.image_stem(img))]
        assert all(torch.equal(a, b) for a, b in zip(got, want)), rep
        view = flat[3:].view(2, *dims)          # 12-byte offset: not 16-byte aligned
        view.copy_(img)
        got = [t.clone() for t in eng.run_network(first=eng.image_stem(view))]
        assert all(torch.equal(a, b) for a, b in zip(got, want)), rep
    with pytest.raises(ValueError):
        eng.image_stem(img.double())


@pytest.mark.parametrize("arch,dims", [("med3ddram18", (32, 40, 48)), ("med3ddram50", (32, 32, 32)), ("med3d", (32, 32, 32))])
def test_commuted_us1_matches_direct_route_and_oracle(cuda, lib, arch, dims, monkeypatch):
    """us1.0 runs commuted by default (K13: low-resolution channel mixing + separable gather); DRAM_B200_US1=direct keeps
    K4 + one convolution over [up(x4) | x1].  Both stay within the parity tolerances of the oracle, and within a few
    16-bit roundings of each other."""
    sd = synthetic.make_state_dict(arch, seed=11, calib_dims=dims)
    x, lung, _ = synthetic.make_network_input(50, dims)
    x, lungs = x[None, None], lung[None, None].float()
    d_ref, s_ref = M.forward(sd, arch, x, lungs)
    outs = {}
    for mode in ("direct", "commute"):
        monkeypatch.setenv("DRAM_B200_US1", mode)
        model = build_model(arch, sd, cuda)
        dense, scores = model(x.to(cuda), lungs.to(cuda))
        names = [s.name for s in model.engine(1, dims, cuda).steps]
        assert ("us1.0.gather_w" in names) == (mode == "commute") and ("us1.upsample" in names) == (mode == "direct")
        check_outputs(arch, dense, scores, d_ref, s_ref)
        outs[mode] = [d.cpu() for d in dense]
    for a, b in zip(outs["direct"], outs["commute"]):
        print(f"{arch} commute vs direct: max {(a - b).abs().max().item():.4g}")
        assert (a - b).abs().max().item() <= 2e-2 * max(1.0, b.abs().max().item())
