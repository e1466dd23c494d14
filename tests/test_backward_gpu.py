"""Training-path kernels: conv3d weight gradient (K9) and data gradient against autograd on CPU
(oracle/backward_oracle.py) on the same 16-bit-rounded operands.

Tolerances: dw is an fp32 accumulation of exact 16-bit products, so only the summation order differs:
|err| <= 1e-4 * max|ref| (+ 1e-5 * |ref|).  dx is rounded once to the storage type: the K1 tolerance
2^-7 * |ref| + 2^-8 * max|ref| (bf16).  Shapes follow the network's conv families on small,
ragged volumes: 64->64 (layer1 / decoder), 128->64 (two-chunk input), stride 2, dilation 2 and 4 with Cout 256 /
512 (two cout blocks), 1x1x1, Cout 32 (us3), the concatenated decoder input (two sources), accumulate mode.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, gen, scale, dtype):
    return (torch.randn(shape, generator=gen) * scale).to(dtype).float()


def _case(cuda, n, dims, cin, cout, k=3, stride=1, dil=1, dtype=torch.bfloat16, seed=0):
    from oracle import backward_oracle as B

    g = torch.Generator().manual_seed(seed)
    k3 = (k, k, k) if isinstance(k, int) else k
    x = _rand((n, cin) + tuple(dims), g, 1.0, dtype)
    w = _rand((cout, cin) + k3, g, (cin * k3[0] * k3[1] * k3[2]) ** -0.5, dtype)
    pad = tuple(dil * (kk - 1) // 2 for kk in k3)
    od = tuple((dims[i] + 2 * pad[i] - dil * (k3[i] - 1) - 1) // stride + 1 for i in range(3))
    dy = _rand((n, cout) + od, g, 1.0, dtype)
    dx_ref, dw_ref = B.conv3d_grads(x, w, dy, stride=stride, dilation=dil, padding=pad)
    return x, w, dy, dx_ref, dw_ref, k3, pad


def _close_dw(got, ref, what):
    err = (got - ref).abs()
    tol = ref.abs().max() * 1e-4 + ref.abs() * 1e-5
    assert bool((err <= tol).all()), f"{what}: max err {err.max().item():.4g} vs max|ref| {ref.abs().max().item():.4g}"


def _close_dx(got, ref, what, dtype):
    eps = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    err = (got - ref).abs()
    tol = ref.abs() * eps + ref.abs().max() * eps / 2
    assert bool((err <= tol).all()), f"{what}: max err {err.max().item():.4g} vs max|ref| {ref.abs().max().item():.4g}"


WGRAD_CASES = [
    # n, dims, cin, cout, k, stride, dil
    (1, (8, 8, 8), 64, 64, 3, 1, 1),
    (1, (1, 16, 8), 64, 64, 3, 1, 1),
    (2, (20, 33, 19), 64, 64, 3, 1, 1),
    (2, (5, 9, 11), 64, 64, 3, 1, 1),
    (1, (6, 10, 9), 128, 64, 3, 1, 1),
    (1, (9, 10, 12), 64, 128, 3, 2, 1),
    (1, (7, 9, 10), 128, 256, 3, 1, 2),
    (1, (9, 9, 12), 64, 512, 3, 1, 4),
    (1, (6, 7, 9), 256, 64, 1, 1, 1),
    (2, (4, 9, 17), 64, 32, 3, 1, 1),
    (1, (5, 6, 7), 192, 128, (1, 3, 3), 1, 1),
]


@pytest.mark.parametrize("n,dims,cin,cout,k,stride,dil", WGRAD_CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("algo", ["auto", "tiles"])
def test_wgrad_matches_autograd(cuda, n, dims, cin, cout, k, stride, dil, dtype, algo):
    from dram_b200 import backward, ops

    if dtype == torch.float16 and (cout, dil) not in ((64, 1), (256, 2)):
        pytest.skip("fp16 re-runs a subset")
    planes_shape = k == 3 and stride == 1 and dil == 1 and cout <= 64
    if algo == "tiles" and not planes_shape:
        pytest.skip("the streaming kernel is what auto picks for this shape")
    x, w, dy, _, dw_ref, k3, pad = _case(cuda, n, dims, cin, cout, k, stride, dil, dtype)
    plan = backward.Conv3dWgradPlan(ops.to_ndhwc_16(x.to(cuda), dtype), ops.to_ndhwc_16(dy.to(cuda), dtype),
                                    kernel=k3, stride=stride, dilation=dil, padding=pad, algo=algo)
    assert plan.algo == ("planes" if planes_shape and algo == "auto" else "stream")
    dw = plan.run().cpu()
    torch.cuda.synchronize()
    _close_dw(dw, dw_ref, f"wgrad[{plan.algo}] {cin}->{cout} k{k} s{stride} d{dil} (items {plan.items}, slices {plan.kslices})")
    # deterministic: a second run reproduces the bits; accumulate adds
    dw2 = plan.run().clone()
    assert torch.equal(dw2.cpu(), dw)
    dw3 = plan.run(accumulate=True).cpu()
    assert torch.allclose(dw3, 2 * dw, rtol=1e-6, atol=0)


def test_wgrad_slices_agree(cuda):
    """A different CTA count changes nothing (the K split is fixed by the plan, not by the launch)."""
    from dram_b200 import backward, ops

    x, w, dy, _, dw_ref, k3, pad = _case(cuda, 1, (16, 16, 16), 64, 64, seed=3)
    plan = backward.Conv3dWgradPlan(ops.to_ndhwc_16(x.to(cuda), torch.bfloat16), ops.to_ndhwc_16(dy.to(cuda), torch.bfloat16),
                                    kernel=k3, padding=pad, algo="tiles")
    a = plan.run().clone()
    b = plan.run(max_ctas=5).clone()
    assert plan.kslices > 1
    assert torch.equal(a, b)
    _close_dw(a.cpu(), dw_ref, "wgrad 16^3")


def test_wgrad_concat_sources(cuda):
    """The decoder's cat([up(x4), x1]) input (med3d.py:87): one dw, two plans with channel offsets."""
    from dram_b200 import backward, ops

    dt = torch.bfloat16
    x, w, dy, _, dw_ref, k3, pad = _case(cuda, 1, (6, 8, 10), 192, 64, seed=5)
    x1, x2 = x[:, :128].contiguous(), x[:, 128:].contiguous()
    dyd = ops.to_ndhwc_16(dy.to(cuda), dt)
    dw = torch.zeros((64, 192, 3, 3, 3), dtype=torch.float32, device=cuda)
    p1 = backward.Conv3dWgradPlan(ops.to_ndhwc_16(x1.to(cuda), dt), dyd, dw=dw, cin_total=192, cin_offset=0)
    p2 = backward.Conv3dWgradPlan(ops.to_ndhwc_16(x2.to(cuda), dt), dyd, dw=dw, cin_total=192, cin_offset=128)
    p1.run()
    p2.run()
    _close_dw(dw.cpu(), dw_ref, "wgrad concat")


def test_wgrad_rejects_bad_arguments(cuda):
    from dram_b200 import _capi, backward

    x = torch.zeros((1, 4, 4, 4, 60), dtype=torch.bfloat16, device=cuda)
    dy = torch.zeros((1, 4, 4, 4, 64), dtype=torch.bfloat16, device=cuda)
    with pytest.raises(_capi.DramError, match="multiple of 64"):
        backward.Conv3dWgradPlan(x, dy)
    with pytest.raises(ValueError, match="dy shape"):
        backward.Conv3dWgradPlan(torch.zeros((1, 4, 4, 4, 64), dtype=torch.bfloat16, device=cuda),
                                 torch.zeros((1, 3, 4, 4, 64), dtype=torch.bfloat16, device=cuda))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        backward.Conv3dWgradPlan(torch.zeros((1, 4, 4, 4, 64), dtype=torch.bfloat16), dy)


DGRAD_CASES = [
    (1, (8, 8, 8), 64, 64, 3, 1, 1),
    (2, (5, 9, 11), 64, 128, 3, 1, 1),
    (1, (7, 9, 10), 128, 256, 3, 1, 2),
    (1, (9, 9, 12), 256, 512, 3, 1, 4),
    (1, (9, 10, 12), 64, 128, 3, 2, 1),
    (1, (8, 10, 12), 64, 128, 3, 2, 1),
    (1, (6, 7, 9), 256, 64, 1, 1, 1),
    (1, (4, 9, 17), 64, 32, 3, 1, 1),
]


@pytest.mark.parametrize("n,dims,cin,cout,k,stride,dil", DGRAD_CASES)
def test_dgrad_matches_autograd(cuda, n, dims, cin, cout, k, stride, dil):
    from dram_b200 import backward, ops

    dtype = torch.bfloat16
    x, w, dy, dx_ref, _, k3, pad = _case(cuda, n, dims, cin, cout, k, stride, dil, dtype, seed=1)
    plan = backward.Conv3dDgradPlan(ops.to_ndhwc_16(dy.to(cuda), dtype), w, dims, kernel=k3, stride=stride,
                                    dilation=dil, padding=pad)
    dx = ops.to_ncdhw_f32(plan.run()).cpu()
    _close_dx(dx, dx_ref, f"dgrad {cin}->{cout} k{k} s{stride} d{dil}", dtype)


def test_dgrad_concat_source_slice(cuda):
    """dx of one source of the concatenated decoder input = transposed conv with that channel slice."""
    from dram_b200 import backward, ops

    dt = torch.bfloat16
    x, w, dy, dx_ref, _, k3, pad = _case(cuda, 1, (6, 8, 10), 192, 64, seed=7)
    plan = backward.Conv3dDgradPlan(ops.to_ndhwc_16(dy.to(cuda), dt), w, (6, 8, 10), cin_range=(128, 192))
    dx = ops.to_ncdhw_f32(plan.run()).cpu()
    _close_dx(dx, dx_ref[:, 128:], "dgrad concat slice", dt)


# ------------------------------------------------------------------------------------------
# K10: train-mode BatchNorm (+ReLU, +residual), forward and backward
# ------------------------------------------------------------------------------------------
BN_CASES = [
    # n, dims, c, with residual, relu
    (2, (5, 6, 7), 64, False, True),
    (1, (9, 8, 11), 64, True, True),
    (2, (4, 4, 4), 512, True, True),
    (1, (6, 10, 9), 32, False, True),
    (1, (7, 5, 6), 128, False, False),
    (1, (3, 4, 5), 2048, True, True),
    (3, (16, 16, 16), 256, False, True),
]


@pytest.mark.parametrize("n,dims,c,with_res,relu", BN_CASES)
def test_bn_train_forward_backward_match_autograd(cuda, n, dims, c, with_res, relu):
    from dram_b200 import backward, ops
    from oracle import backward_oracle as B

    dt = torch.bfloat16
    g = torch.Generator().manual_seed(c + n)
    x = _rand((n, c) + dims, g, 1.5, dt) + _rand((1, c, 1, 1, 1), g, 1.0, dt)
    x = x.to(dt).float()
    res = _rand((n, c) + dims, g, 1.0, dt) if with_res else None
    dy = _rand((n, c) + dims, g, 1.0, dt)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    y_ref, rm_ref, rv_ref, grads_ref = B.bn_train(x, gamma, beta, res, relu, dy)

    xd = ops.to_ndhwc_16(x.to(cuda), dt)
    rd = None if res is None else ops.to_ndhwc_16(res.to(cuda), dt).requires_grad_(True)
    gd, bd = gamma.to(cuda).requires_grad_(True), beta.to(cuda).requires_grad_(True)
    xd.requires_grad_(True)
    rm, rv = torch.zeros(c, device=cuda), torch.ones(c, device=cuda)
    y = backward.BatchNormTrainFn.apply(xd, gd, bd, rm, rv, rd, relu, 1e-5, 0.1, None)
    y.backward(ops.to_ndhwc_16(dy.to(cuda), dt))
    torch.cuda.synchronize()

    _close_dx(ops.to_ncdhw_f32(y.detach()).cpu(), y_ref, "bn y", dt)
    assert torch.allclose(rm.cpu(), rm_ref, rtol=1e-4, atol=1e-5) and torch.allclose(rv.cpu(), rv_ref, rtol=1e-4, atol=1e-5)
    dx_ref, dg_ref, db_ref = grads_ref[:3]
    # dx is rounded once to bf16; its small terms are differences of O(1) quantities, so bound by the tensor's scale
    err = (ops.to_ncdhw_f32(xd.grad).cpu() - dx_ref).abs()
    assert bool((err <= dx_ref.abs() * 2.0 ** -7 + dx_ref.abs().max() * 2.0 ** -8).all()), err.max().item()
    assert torch.allclose(gd.grad.cpu(), dg_ref, rtol=2e-3, atol=2e-3 * float(dg_ref.abs().max()))
    assert torch.allclose(bd.grad.cpu(), db_ref, rtol=2e-3, atol=2e-3 * float(db_ref.abs().max()))
    if with_res:
        _close_dx(ops.to_ncdhw_f32(rd.grad).cpu(), grads_ref[3], "bn dres", dt)


def test_bn_train_is_deterministic_and_rejects_bad_shapes(cuda):
    from dram_b200 import _capi, backward

    x = (torch.randn((2, 8, 8, 8, 64), device=cuda)).to(torch.bfloat16)
    g, b = torch.ones(64, device=cuda), torch.zeros(64, device=cuda)
    outs = []
    for _ in range(2):
        rm, rv = torch.zeros(64, device=cuda), torch.ones(64, device=cuda)
        outs.append((backward.BatchNormTrainFn.apply(x, g, b, rm, rv, None, True, 1e-5, 0.1, None), rm, rv))
    assert all(torch.equal(a, b_) for a, b_ in zip(outs[0], outs[1]))
    bad = torch.zeros((1, 2, 2, 2, 24), dtype=torch.bfloat16, device=cuda)  # 256 % (24 / 8) != 0
    with pytest.raises(_capi.DramError, match="channels"):
        backward.BatchNormTrainFn.apply(bad, torch.ones(24, device=cuda), torch.zeros(24, device=cuda),
                                        torch.zeros(24, device=cuda), torch.ones(24, device=cuda), None, True, 1e-5, 0.1, None)


# ------------------------------------------------------------------------------------------
# K4T / K3T: backward of the x2 up-sampling and of the max-pool
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,dims,c", [(1, (4, 4, 4), 64), (2, (3, 5, 7), 64), (1, (8, 8, 8), 512), (1, (1, 2, 9), 8)])
def test_upsample2x_backward_matches_autograd(cuda, n, dims, c):
    from dram_b200 import backward, ops
    from oracle import backward_oracle as B

    dt = torch.bfloat16
    g = torch.Generator().manual_seed(11)
    x = _rand((n, c) + dims, g, 1.0, dt)
    dy = _rand((n, c) + tuple(2 * v for v in dims), g, 1.0, dt)
    y_ref, dx_ref = B.upsample2x_grad(x, dy)
    xd = ops.to_ndhwc_16(x.to(cuda), dt).requires_grad_(True)
    y = backward.Upsample2xFn.apply(xd)
    y.backward(ops.to_ndhwc_16(dy.to(cuda), dt))
    _close_dx(ops.to_ncdhw_f32(y.detach()).cpu(), y_ref, "upsample y", dt)
    _close_dx(ops.to_ncdhw_f32(xd.grad).cpu(), dx_ref, "upsample dx", dt)


@pytest.mark.parametrize("n,dims,c", [(1, (8, 8, 8), 64), (2, (5, 7, 9), 64), (1, (6, 6, 6), 8), (1, (1, 3, 2), 16)])
def test_maxpool3d_backward_matches_autograd_with_ties(cuda, n, dims, c):
    """Post-ReLU inputs: half of the values are exactly zero, so most windows hold ties; the gradient must land on the
    same element ATen picks (first maximum in scan order) — compared exactly up to the bf16 rounding of sums."""
    from dram_b200 import backward, ops
    from oracle import backward_oracle as B

    dt = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    x = _rand((n, c) + dims, g, 1.0, dt).relu()
    x = (x * 4).round() / 4  # coarse values: ties between positive entries too
    od = tuple((v - 1) // 2 + 1 for v in dims)
    dy = _rand((n, c) + od, g, 1.0, dt)
    y_ref, dx_ref = B.maxpool3d_grad(x, dy)
    xd = ops.to_ndhwc_16(x.to(cuda), dt).requires_grad_(True)
    y = backward.MaxPool3dFn.apply(xd)
    y.backward(ops.to_ndhwc_16(dy.to(cuda), dt))
    assert torch.equal(ops.to_ncdhw_f32(y.detach()).cpu(), y_ref)
    got = ops.to_ncdhw_f32(xd.grad).cpu()
    assert torch.equal((got != 0), (dx_ref != 0)), "gradient routed to different elements"
    _close_dx(got, dx_ref, "maxpool dx", dt)


def test_heads_sigmoid_forward_backward_match_autograd(cuda):
    """K5T against torch CPU autograd of sigmoid(conv1x1x1(x)) for the two regression heads (med3d.py:329-332, 382)."""
    import torch.nn.functional as F

    from dram_b200 import backward, ops

    dt = torch.bfloat16
    g = torch.Generator().manual_seed(9)
    x = _rand((2, 32, 5, 6, 7), g, 1.0, dt)
    ws = [(torch.randn((1, 32, 1, 1, 1), generator=g) * 0.3).requires_grad_(True) for _ in range(2)]
    bs = [(torch.randn(1, generator=g) * 0.5).requires_grad_(True) for _ in range(2)]
    gs = [torch.randn((2, 1, 5, 6, 7), generator=g) for _ in range(2)]
    xr = x.clone().requires_grad_(True)
    outs_ref = [torch.sigmoid(F.conv3d(xr, w, b)) for w, b in zip(ws, bs)]
    grads_ref = torch.autograd.grad(outs_ref, [xr] + ws + bs, gs)

    xd = ops.to_ndhwc_16(x.to(cuda), dt).requires_grad_(True)
    wd = [w.detach().to(cuda).requires_grad_(True) for w in ws]
    bd = [b.detach().to(cuda).requires_grad_(True) for b in bs]
    o0, o1 = backward.HeadsSigmoidFn.apply(xd, wd[0], bd[0], wd[1], bd[1])
    torch.autograd.backward([o0, o1], [gs[0].to(cuda), gs[1].to(cuda)])
    for got, ref in zip((o0, o1), outs_ref):
        assert torch.allclose(got.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-6)
    _close_dx(ops.to_ncdhw_f32(xd.grad).cpu(), grads_ref[0], "heads dx", dt)
    for got, ref in zip([w.grad for w in wd] + [b.grad for b in bd], grads_ref[1:]):
        assert torch.allclose(got.cpu(), ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max() + 1)), (got.cpu(), ref)


def test_pack_weight_into_equals_the_torch_packers(cuda):
    """The device re-packing kernel against ops.pack_conv_weight / backward.pack_dgrad_weight (bit-exact)."""
    from dram_b200 import backward, ops

    g = torch.Generator().manual_seed(2)
    for shape, rng in (((64, 128, 3, 3, 3), None), ((32, 64, 3, 3, 3), None), ((64, 192, 3, 3, 3), (128, 192)),
                       ((256, 64, 1, 1, 1), None)):
        w = torch.randn(shape, generator=g).to(cuda)
        for dt in (torch.bfloat16, torch.float16):
            if rng is None:
                ref = ops.pack_conv_weight(w, dtype=dt)
                assert torch.equal(backward.pack_weight_into(w, torch.empty_like(ref)), ref)
            ref = backward.pack_dgrad_weight(w, dtype=dt, cin_range=rng)
            assert torch.equal(backward.pack_weight_into(w, torch.empty_like(ref), transpose=True, cin_range=rng), ref)


# ------------------------------------------------------------------------------------------
# K11 (training loss + gradient) and K12 (Adam)
# ------------------------------------------------------------------------------------------
LOSS_CASES = [
    # batch, mask dims, map dims, (cle labels, pse labels)
    (2, (32, 32, 32), (16, 16, 16), ([3, 0], [1, 2])),      # the x2 case of the network
    (3, (20, 24, 28), (7, 9, 11), ([0, 2, 0], [0, 0, 1])),  # generic legacy-nearest indices; sample 0 has no label
    (1, (9, 10, 70), (9, 10, 70), ([5], [0])),              # same size, rows longer than two warp passes
    (2, (16, 16, 16), (8, 8, 8), ([0, 0], [0, 0])),         # no positive label at all: alpha clamps to 0.7
]


@pytest.mark.parametrize("case", LOSS_CASES)
def test_train_loss_kernel_matches_oracle(cuda, lib, case):
    """K11 against the CPU oracle of models.py:547-565 (oracle/training_oracle.total_loss, pinned to the reference by
    the golden training step) and autograd: loss and its four terms within 1e-5 relative, the lobe-masked means within
    1e-6, the gradient of both maps within 1e-5 of its largest entry (+1e-4 relative).  The maps include sums above 1
    (clamp(cle + pse, 0, 1) blocks the gradient) and values at the 1e-6 clamp of the cross entropy."""
    import torch.nn.functional as F

    from dram_b200.backward import TrainLossFn
    from oracle import training_oracle as T

    B, mdims, dims, (cl, pl) = case
    g = torch.Generator().manual_seed(B * 100 + dims[0])
    d0 = torch.rand((B, 1) + dims, generator=g) * 0.8
    d1 = torch.rand((B, 1) + dims, generator=g) * 0.8
    d0.view(-1)[:7], d1.view(-1)[:7] = 3e-7, 2e-7   # pt = 1 - 5e-7 > 1 - eps where t = 0, pt = 5e-7 < eps where t = 1
    lungs = torch.rand((B,) + mdims, generator=g) > 0.45
    ems = torch.rand((B,) + mdims, generator=g) > 0.7
    cl, pl = torch.tensor(cl), torch.tensor(pl)
    from dram_b200.models import CLE_RATIO_MAP, PSE_RATIO_MAP
    cb, pb = T.label_bands(cl, CLE_RATIO_MAP), T.label_bands(pl, PSE_RATIO_MAP)
    cw, pw = torch.rand(B, generator=g) + 0.5, torch.rand(B, generator=g) + 0.5

    a, b = d0.clone().requires_grad_(True), d1.clone().requires_grad_(True)
    lm = F.interpolate(lungs.unsqueeze(1).float(), dims, mode="nearest")
    regs_ref = [(x * lm).view(B, -1).sum(-1) / lm.view(B, -1).sum(-1) for x in (a, b)]
    loss_ref = T.total_loss([a, b], regs_ref, lungs.unsqueeze(1).float(), ems.unsqueeze(1).float(), cl, pl, cb, pb, cw, pw)
    (3.0 * loss_ref).backward()

    x0, x1 = d0.to(cuda).requires_grad_(True), d1.to(cuda).requires_grad_(True)
    loss, terms, regs = TrainLossFn.apply(x0, x1, lungs.to(cuda), ems.to(cuda), cl.to(cuda), pl.to(cuda), cb.to(cuda),
                                          pb.to(cuda), cw.to(cuda), pw.to(cuda))
    (3.0 * loss).backward()
    torch.cuda.synchronize()
    loss_v, ref_v = float(loss.detach()), float(loss_ref.detach())
    assert abs(loss_v - ref_v) <= 1e-5 * abs(ref_v), (loss_v, ref_v)
    assert abs(float(terms[0] + terms[1] + 2 * terms[2] + terms[3]) - loss_v) <= 1e-6 * abs(loss_v)
    for k in (0, 1):
        assert torch.allclose(regs[:, k].cpu(), regs_ref[k].detach(), rtol=1e-6, atol=1e-7)
        got, ref = (x0, x1)[k].grad.cpu(), (a, b)[k].grad
        tol = ref.abs().max() * 1e-5 + ref.abs() * 1e-4
        assert bool(((got - ref).abs() <= tol).all()), (k, (got - ref).abs().max().item(), ref.abs().max().item())
    # deterministic: a second evaluation gives the same bits
    loss2, _, _ = TrainLossFn.apply(x0.detach(), x1.detach(), lungs.to(cuda), ems.to(cuda), cl.to(cuda), pl.to(cuda),
                                    cb.to(cuda), pb.to(cuda), cw.to(cuda), pw.to(cuda))
    assert torch.equal(loss2, loss.detach())


def test_train_loss_rejects_bad_input(cuda, lib):
    from dram_b200.backward import TrainLossFn

    m = torch.rand(2, 1, 4, 4, 4, device=cuda)
    mask = torch.ones(2, 8, 8, 8, dtype=torch.bool, device=cuda)
    lab, band, w = torch.zeros(2, dtype=torch.int64), torch.zeros(2, 2), torch.ones(2)
    with pytest.raises(ValueError, match="cle_bands"):
        TrainLossFn.apply(m, m.clone(), mask, mask, lab, lab, torch.zeros(2, 3), band, w, w)
    with pytest.raises(ValueError, match="masks"):
        TrainLossFn.apply(m, m.clone(), mask[:1], mask[:1], lab, lab, band, band, w, w)
    with pytest.raises(RuntimeError, match="no CPU"):
        TrainLossFn.apply(m, m.clone(), mask.cpu(), mask, lab, lab, band, band, w, w)


@pytest.mark.parametrize("n", [4096, 1003, 3])
def test_adam_kernel_matches_torch_adam(cuda, lib, n):
    """K12 against torch.optim.Adam (models.py:685-698: lr only) on CPU over four steps with fresh gradients, one of
    them with a changed learning rate: parameters within 1e-6 relative + 1e-7, moments within 1e-6 relative."""
    from dram_b200.backward import FlatAdam

    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    flat_p, flat_g = p0.to(cuda), torch.zeros(n, device=cuda)
    adam = FlatAdam(flat_p, flat_g, lr=1e-3)
    for it in range(4):
        grad = torch.randn(n, generator=g) * (10.0 ** (it - 2))
        grad[::7] = 0.0
        if it == 2:
            adam.decay_lr(0.95)
            opt.param_groups[0]["lr"] *= 0.95
        ref.grad = grad.clone()
        opt.step()
        flat_g.copy_(grad)
        adam.step()
    torch.cuda.synchronize()
    st = opt.state[ref]
    assert torch.allclose(flat_p.cpu(), ref.detach(), rtol=1e-6, atol=1e-7), (flat_p.cpu() - ref.detach()).abs().max().item()
    assert torch.allclose(adam.exp_avg.cpu(), st["exp_avg"], rtol=1e-6, atol=1e-12)
    assert torch.allclose(adam.exp_avg_sq.cpu(), st["exp_avg_sq"], rtol=1e-6, atol=1e-20)
    # grad_scale folds a division (gradient averaging) into the update
    p1, p2 = p0.to(cuda), p0.to(cuda)
    a1, a2 = FlatAdam(p1, flat_g, lr=1e-3), FlatAdam(p2, flat_g * 0.25, lr=1e-3)
    a1.step(grad_scale=0.25)
    a2.step()
    assert torch.equal(p1, p2)


def test_peer_allreduce_protocol_two_ranks_on_one_device(cuda, lib):
    """K10x's protocol (push rows, release flags, acquire-wait, rank-order sum, four rotating slots, sequence numbers)
    exercised without a second process: two exchange buffers on one GPU play rank 0 and rank 1, their kernels run on
    two streams and wait for each other.  40 consecutive exchanges of varying length, in place and out of place; both
    'ranks' must end with exactly a + b (two-term sums commute, so the bits are the plain fp64 sum).  The real
    two-process run over CUDA IPC is tools/train_ddp_smoke.py (profiles/train_ddp_n2_*.log)."""
    import ctypes as C

    from dram_b200 import _capi

    nbytes = lib.dram_peer_exchange_bytes()
    bufs, handle = [], (C.c_ubyte * _capi.PEER_HANDLE_BYTES)()
    for _ in range(2):
        ptr = C.c_void_p()
        _capi.check(lib.dram_peer_alloc(nbytes, C.byref(ptr), handle), "dram_peer_alloc")
        bufs.append(ptr.value)
    peers = (C.c_void_p * 8)()
    peers[0], peers[1] = bufs
    status = [torch.zeros(1, dtype=torch.int32, device=cuda) for _ in range(2)]
    streams = [torch.cuda.Stream(device=cuda) for _ in range(2)]
    g = torch.Generator().manual_seed(5)
    try:
        for seq in range(1, 41):
            n = [128, 256, 4096, 2, 1026][seq % 5]
            a, b = torch.randn(n, generator=g, dtype=torch.float64), torch.randn(n, generator=g, dtype=torch.float64)
            ins = [a.to(cuda), b.to(cuda)]
            outs = ins if seq % 2 else [torch.empty_like(t) for t in ins]
            torch.cuda.synchronize()
            for r in (1, 0) if seq % 3 == 0 else (0, 1):  # either 'rank' may arrive first
                with torch.cuda.stream(streams[r]):
                    _capi.check(lib.dram_peer_allreduce_f64(ins[r].data_ptr(), outs[r].data_ptr(), n, peers, 2, r, seq,
                                                            status[r].data_ptr(), streams[r].cuda_stream),
                                "dram_peer_allreduce_f64")
            torch.cuda.synchronize()
            assert int(status[0]) == 0 and int(status[1]) == 0, seq
            assert torch.equal(outs[0].cpu(), a + b) and torch.equal(outs[1].cpu(), a + b), seq
        # a single rank is the identity
        x = torch.randn(130, dtype=torch.float64, device=cuda)
        y = torch.empty_like(x)
        solo = (C.c_void_p * 8)()
        solo[0] = bufs[0]
        _capi.check(lib.dram_peer_allreduce_f64(x.data_ptr(), y.data_ptr(), 130, solo, 1, 0, 41, status[0].data_ptr(), None),
                    "dram_peer_allreduce_f64")
        torch.cuda.synchronize()
        assert torch.equal(x, y)
        assert lib.dram_peer_allreduce_f64(x.data_ptr(), y.data_ptr(), 5000, solo, 1, 0, 42, status[0].data_ptr(), None) == -1
        assert lib.dram_peer_allreduce_f64(x.data_ptr(), y.data_ptr(), 130, solo, 9, 0, 42, status[0].data_ptr(), None) == -1
    finally:
        torch.cuda.synchronize()
        for ptr in bufs:
            lib.dram_peer_free(C.c_void_p(ptr))
