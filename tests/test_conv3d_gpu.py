"""K1 parity: the tcgen05 implicit-GEMM conv3d against torch's CPU fp32 conv3d on the same
bf16-rounded operands (so the only differences are fp32 accumulation order and the single
bf16 rounding of the output).  Tolerance: |err| <= 2^-7 * |ref| + 2^-8 * max|ref|.
Covers every conv shape family of the network: 3^3 with dilation 1/2/4,
stride 2 with shortcut A, 1x1x1, two K-sources (skip concat), 32-channel conv with fused heads,
the unfolded 7^3 stem, ragged volumes (partial tiles) and batch > 1.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


DT = torch.bfloat16  # storage type under test; the fp16 variants re-run a subset


def _rand(shape, gen, scale=1.0, dtype=None):
    return (torch.randn(shape, generator=gen) * scale).to(dtype or DT).float()


def _close(got, ref, what):
    err = (got - ref).abs()
    tol = ref.abs() * 2.0 ** -7 + ref.abs().max() * 2.0 ** -8
    bad = err > tol
    if bad.any():
        idx = torch.nonzero(bad)[0].tolist()
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.numel()} outside tolerance; max err {err.max().item():.4g} "
            f"(ref max {ref.abs().max().item():.4g}); first bad index {idx}: got {got[tuple(idx)].item():.5g} "
            f"ref {ref[tuple(idx)].item():.5g}")


def _run_conv(cuda, n, dims, c1, c2, cout, k, stride, dil, relu=True, residual=None, res_stride=1,
              heads=None, seed=0, tile=None, dtype=torch.bfloat16, normalize=False, algo="auto", max_ctas=0,
              expect_algo=None):
    from dram_b200 import ops

    global DT
    DT = dtype
    g = torch.Generator().manual_seed(seed)
    d, h, w = dims
    k3 = (k, k, k) if isinstance(k, int) else k
    x1 = _rand((n, c1, d, h, w), g)
    x2 = _rand((n, c2, d, h, w), g) if c2 else None
    cin = c1 + c2
    fan = cin * k3[0] * k3[1] * k3[2]
    wgt = _rand((cout, cin) + k3, g, scale=fan ** -0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    pad = tuple(dil * (kk - 1) // 2 for kk in k3)
    xin = x1 if x2 is None else torch.cat([x1, x2], 1)
    ref = F.conv3d(xin, wgt, None, stride=stride, padding=pad, dilation=dil) + bias.view(1, -1, 1, 1, 1)
    res_t = None
    if residual is not None:
        rc = residual
        res = _rand((n, rc, d, h, w), g)
        sub = res[:, :, ::res_stride, ::res_stride, ::res_stride]
        ref[:, :rc] += sub[:, :, :ref.shape[2], :ref.shape[3], :ref.shape[4]]
        res_t = ops.to_ndhwc_16(res.to(cuda), dtype)
    if relu:
        ref = ref.relu()
    head_ref = None
    hk = None
    if heads is not None:
        chs, sig = heads
        hw = torch.randn(sum(chs), 32, generator=g) * 0.3
        hb = torch.randn(sum(chs), generator=g) * 0.1
        dense = torch.einsum("oc,ncdhw->nodhw", hw, ref) + hb.view(1, -1, 1, 1, 1)
        if sig:
            dense = torch.sigmoid(dense)
        head_ref = torch.split(dense, list(chs), dim=1)
        hk = (hw.to(cuda), hb.to(cuda), tuple(chs), sig)
    if normalize:
        wp, mult = ops.pack_conv_weight(wgt.to(cuda), dtype=dtype, normalize=True)
    else:
        wp, mult = ops.pack_conv_weight(wgt, dtype=dtype).to(cuda), None
    plan = ops.Conv3dPlan(
        ops.to_ndhwc_16(x1.to(cuda), dtype), wp, bias.to(cuda), scale=mult,
        x2=None if x2 is None else ops.to_ndhwc_16(x2.to(cuda), dtype), kernel=k3, stride=stride, dilation=dil,
        relu=relu, residual=res_t, res_stride=res_stride, heads=hk, tile=tile, algo=algo)
    if expect_algo is not None:
        assert plan.algo == expect_algo, (plan.algo, expect_algo)
    out = plan.run(max_ctas)
    torch.cuda.synchronize()
    got = ops.to_ncdhw_f32(out).cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    _close(got, ref, f"conv {c1}+{c2}->{cout} k{k} s{stride} d{dil} {dims}")
    if head_ref is not None:
        for i, hr in enumerate(head_ref):
            hg = plan.head_outs[i].cpu()
            assert hg.shape == hr.shape
            err = (hg - hr).abs().max().item()
            assert err < 2e-2 * max(1.0, hr.abs().max().item()), f"head {i}: max err {err}"
    return plan


def test_conv_64_64_plain(cuda, lib):
    _run_conv(cuda, 1, (16, 16, 16), 64, 0, 64, 3, 1, 1)


def test_conv_64_64_residual_ragged(cuda, lib):
    _run_conv(cuda, 1, (10, 12, 20), 64, 0, 64, 3, 1, 1, residual=64)


def test_conv_128_256_dil2(cuda, lib):
    _run_conv(cuda, 1, (12, 16, 20), 128, 0, 256, 3, 1, 2, residual=128)  # layer3.0: shortcut A, stride 1


def test_conv_256_512_dil4_batch2(cuda, lib):
    _run_conv(cuda, 2, (8, 12, 16), 256, 0, 512, 3, 1, 4, residual=512)


def test_conv_stride2_shortcut_a(cuda, lib):
    _run_conv(cuda, 1, (16, 16, 16), 64, 0, 128, 3, 2, 1, relu=True)
    p = _run_conv(cuda, 1, (16, 24, 32), 64, 0, 128, 3, 2, 1, relu=False)
    assert p.out_shape == (1, 8, 12, 16, 128)


def test_conv_stride2_with_residual(cuda, lib):
    # conv2 of layer2.0: stride-1 conv whose residual is the stride-2 subsample of the block input
    from dram_b200 import ops

    g = torch.Generator().manual_seed(5)
    x = _rand((1, 128, 8, 8, 8), g)
    res = _rand((1, 64, 16, 16, 16), g)
    wgt = _rand((128, 128, 3, 3, 3), g, scale=(128 * 27) ** -0.5)
    bias = torch.randn(128, generator=g) * 0.1
    ref = F.conv3d(x, wgt, None, padding=1) + bias.view(1, -1, 1, 1, 1)
    ref[:, :64] += res[:, :, ::2, ::2, ::2]
    ref = ref.relu()
    plan = ops.Conv3dPlan(ops.to_ndhwc_bf16(x.to(cuda)), ops.pack_conv_weight(wgt).to(cuda), bias.to(cuda),
                          residual=ops.to_ndhwc_bf16(res.to(cuda)), res_stride=2)
    got = ops.to_ncdhw_f32(plan.run()).cpu()
    _close(got, ref, "layer2.0.conv2 + shortcut A")


def test_conv_two_sources(cuda, lib):
    _run_conv(cuda, 1, (8, 16, 16), 128, 64, 64, 3, 1, 1)


def test_conv_1x1(cuda, lib):
    _run_conv(cuda, 1, (8, 12, 16), 256, 0, 128, 1, 1, 1)
    _run_conv(cuda, 1, (8, 8, 16), 64, 0, 256, 1, 1, 1, residual=64)  # R50 layer1.0 conv3 + shortcut A


def test_conv_heads_reg_and_cls(cuda, lib):
    _run_conv(cuda, 1, (8, 16, 16), 64, 0, 32, 3, 1, 1, heads=((1, 1), True))
    _run_conv(cuda, 2, (8, 8, 12), 64, 0, 32, 3, 1, 1, heads=((6, 3), False))


def test_conv_fp16_storage(cuda, lib):
    """Same kernel with IEEE fp16 operands/activations (DRAM_DTYPE_F16) and the per-channel power-of-two
    weight normaliser applied in the epilogue."""
    h = torch.float16
    _run_conv(cuda, 1, (10, 12, 20), 64, 0, 64, 3, 1, 1, residual=64, dtype=h)
    _run_conv(cuda, 2, (8, 12, 16), 256, 0, 512, 3, 1, 4, residual=512, dtype=h, normalize=True)
    _run_conv(cuda, 1, (16, 16, 16), 64, 0, 128, 3, 2, 1, residual=64, res_stride=1, dtype=h, normalize=True)
    _run_conv(cuda, 1, (8, 16, 16), 128, 64, 64, 3, 1, 1, dtype=h)
    _run_conv(cuda, 1, (8, 16, 16), 64, 0, 32, 3, 1, 1, heads=((1, 1), True), dtype=h, normalize=True)
    _run_conv(cuda, 1, (8, 16, 16), 64, 0, 64, 3, 1, 1, dtype=torch.bfloat16, normalize=True)


PLANE_CASES = [
    # n, dims, c1, c2, cout, kwargs
    (1, (8, 16, 16), 64, 0, 64, {}),
    (1, (10, 20, 12), 64, 0, 64, {"residual": 64}),                   # ragged: partial slabs and groups
    (2, (12, 32, 24), 64, 0, 64, {"max_ctas": 3}),                    # several items per CTA: plane reuse, column changes
    (1, (16, 16, 8), 64, 0, 64, {"max_ctas": 1, "residual": 64}),     # one CTA marches through everything
    (1, (8, 16, 16), 128, 64, 64, {"max_ctas": 2}),                   # two sources, 3 K-chunks (no plane reuse)
    (1, (9, 18, 10), 64, 64, 64, {}),
    (1, (8, 16, 16), 64, 0, 32, {"heads": ((1, 1), True)}),           # us3 + regression heads
    (2, (6, 16, 24), 64, 0, 32, {"heads": ((6, 3), False), "max_ctas": 2}),
    (1, (8, 16, 16), 128, 0, 128, {}),                                # layer2: Cout 128, two output planes per item
    (2, (9, 18, 20), 128, 0, 128, {"residual": 128}),                 # ragged groups (odd D), residual
    (1, (12, 32, 16), 128, 0, 128, {"max_ctas": 2, "residual": 128}), # several items per CTA
    (1, (7, 16, 8), 64, 0, 128, {"max_ctas": 1}),                     # one chunk: plane reuse between items of a column
    (1, (6, 10, 12), 128, 64, 128, {}),                               # two sources
]


@pytest.mark.parametrize("case", range(len(PLANE_CASES)))
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_conv_plane_ring_kernel(cuda, lib, case, dt):
    """DRAM_CONV_ALGO_PLANES: same results as the reference conv for every shape family it serves."""
    n, dims, c1, c2, cout, kw = PLANE_CASES[case]
    _run_conv(cuda, n, dims, c1, c2, cout, 3, 1, 1, dtype=dt, algo="planes", expect_algo="planes", seed=11 + case,
              normalize=(case % 2 == 0), **kw)


STREAM_CASES = [
    # n, dims, kwargs: Cin 64 -> Cout 32, one source = the weights-resident streaming kernel (us3)
    (1, (40, 16, 8), {"max_ctas": 1}),                                 # one CTA, one column: the 16-block TMEM ring wraps twice
    (1, (40, 16, 16), {"max_ctas": 3, "heads": ((1, 1), True)}),       # columns cut mid-way: halo planes at both cuts
    (2, (21, 20, 12), {"max_ctas": 5, "heads": ((6, 3), False)}),      # ragged H / W, odd D, two samples
    (1, (1, 16, 8), {}),                                               # D = 1: the only plane is first and last
    (1, (2, 16, 8), {"max_ctas": 2}),                                  # one output plane per CTA
    (1, (19, 32, 24), {"residual": 32}),                               # residual rows, 148 > steps / few steps per CTA
    (2, (33, 16, 8), {"max_ctas": 7, "residual": 32, "heads": ((1, 1), True)}),
]


@pytest.mark.parametrize("case", range(len(STREAM_CASES)))
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_conv_streaming_kernel(cuda, lib, case, dt):
    """us3 (med3d.py:90): planes stream past resident weights; every way a CTA's range can cut a column."""
    n, dims, kw = STREAM_CASES[case]
    _run_conv(cuda, n, dims, 64, 0, 32, 3, 1, 1, dtype=dt, algo="planes", expect_algo="planes", seed=140 + case,
              normalize=(case % 2 == 1), **kw)


def test_conv_streaming_equals_plane_items(cuda, lib, monkeypatch):
    """DRAM_B200_US3=ring keeps the 4-plane work items for Cout 32: same operands, same fp32 sums up to their order."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(151)
    x = ops.to_ndhwc_16(_rand((1, 64, 24, 32, 16), g, dtype=torch.float16).to(cuda), torch.float16)
    wp = ops.pack_conv_weight(_rand((32, 64, 3, 3, 3), g, scale=1728 ** -0.5, dtype=torch.float16),
                              dtype=torch.float16).to(cuda)
    bias = (torch.randn(32, generator=g) * 0.1).to(cuda)
    a = ops.Conv3dPlan(x, wp, bias, algo="planes").run(5).float().clone()
    monkeypatch.setenv("DRAM_B200_US3", "ring")
    b = ops.Conv3dPlan(x, wp, bias, algo="planes").run(5).float().clone()
    torch.cuda.synchronize()
    assert (a - b).abs().max().item() <= 2.0 ** -9 * max(1.0, b.abs().max().item())
    assert a.abs().sum().item() > 0


def test_conv_algo_dispatch(cuda, lib):
    p = _run_conv(cuda, 1, (8, 16, 16), 64, 0, 64, 3, 1, 1)                 # auto -> planes
    assert p.algo == "planes"
    p = _run_conv(cuda, 1, (8, 16, 16), 64, 0, 64, 3, 1, 1, algo="tiles")
    assert p.algo == "tiles"
    p = _run_conv(cuda, 1, (8, 16, 16), 64, 0, 64, 3, 1, 2)                 # dilation 2 -> tiles
    assert p.algo == "tiles"
    p = _run_conv(cuda, 1, (8, 16, 16), 64, 0, 128, 3, 1, 1)                # cout 128 -> planes (two planes per item)
    assert p.algo == "planes"
    p = _run_conv(cuda, 1, (8, 16, 16), 64, 0, 256, 3, 1, 1)                # cout 256 -> tiles
    assert p.algo == "tiles"


def test_conv_tile_shapes(cuda, lib):
    for tile in [(8, 4, 4), (4, 4, 8), (16, 8, 1), (32, 2, 2), (128, 1, 1)]:
        _run_conv(cuda, 1, (8, 16, 36), 64, 0, 64, 3, 1, 1, tile=tile, seed=7)


def test_stem_unfold(cuda, lib):
    """7^3 stride-2 conv (med3d.py:296-304) = stem_expand + 7x1x1 conv over 64 pseudo-channels."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(3)
    for dims in [(16, 16, 32), (12, 20, 24)]:
        x = _rand((2, 1) + dims, g)
        wgt = _rand((64, 1, 7, 7, 7), g, scale=343 ** -0.5)
        bias = torch.randn(64, generator=g) * 0.1
        ref = (F.conv3d(x, wgt, None, stride=2, padding=3) + bias.view(1, -1, 1, 1, 1)).relu()
        xe = ops.stem_expand(x[:, 0].contiguous().to(cuda))
        plan = ops.Conv3dPlan(xe, ops.pack_stem_weight(wgt).to(cuda), bias.to(cuda), kernel=(7, 1, 1),
                              stride=(2, 1, 1), padding=(3, 0, 0), tile=(16, 8, 1))
        got = ops.to_ncdhw_f32(plan.run()).cpu()
        assert got.shape == ref.shape, (got.shape, ref.shape)
        _close(got, ref, f"stem {dims}")


STEM_CASES = [
    # (batch, dims, max_ctas): even/odd sizes, several columns and D-groups, partial tiles, CTA counts that
    # split columns mid-way (plane reuse across groups on and off)
    (1, (16, 16, 32), 0),
    (2, (12, 20, 24), 0),
    (1, (9, 13, 17), 0),
    (1, (40, 34, 36), 3),
    (2, (33, 40, 20), 5),
    (1, (64, 32, 16), 1),
    (1, (7, 8, 8), 0),
]


@pytest.mark.parametrize("case", range(len(STEM_CASES)))
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_stem_fused(cuda, lib, case, dt):
    """K2: Conv3d(1,64,k7,s2,p3)+shift+ReLU straight from the fp32 image (med3d.py:296-304, 371-373)."""
    from dram_b200 import ops

    n, dims, max_ctas = STEM_CASES[case]
    g = torch.Generator().manual_seed(40 + case)
    x = _rand((n, 1) + dims, g, dtype=dt)
    wgt = _rand((64, 1, 7, 7, 7), g, scale=343 ** -0.5, dtype=dt)
    bias = torch.randn(64, generator=g) * 0.1
    ref = (F.conv3d(x, wgt, None, stride=2, padding=3) + bias.view(1, -1, 1, 1, 1)).relu()
    wp, mult = ops.pack_stem_weight_fused(wgt.to(cuda), dtype=dt, normalize=True)
    out = ops.stem_conv7(x[:, 0].contiguous().to(cuda), wp, bias.to(cuda), mult, max_ctas=max_ctas)
    torch.cuda.synchronize()
    got = ops.to_ncdhw_f32(out).cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    _close(got, ref, f"fused stem {dims} x{n} ctas {max_ctas}")


def test_stem_fused_matches_unfold_route(cuda, lib):
    """Both stem routes compute the same dot products from the same 16-bit operands."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(77)
    x = _rand((1, 1, 24, 32, 40), g, dtype=torch.float16)
    wgt = _rand((64, 1, 7, 7, 7), g, scale=343 ** -0.5, dtype=torch.float16)
    bias = (torch.randn(64, generator=g) * 0.1).to(cuda)
    xc = x[:, 0].contiguous().to(cuda)
    xe = ops.stem_expand(xc, dtype=torch.float16)
    plan = ops.Conv3dPlan(xe, ops.pack_stem_weight(wgt, dtype=torch.float16).to(cuda), bias, kernel=(7, 1, 1),
                          stride=(2, 1, 1), padding=(3, 0, 0), tile=(16, 8, 1))
    a = plan.run().float().cpu()
    b = ops.stem_conv7(xc, ops.pack_stem_weight_fused(wgt.to(cuda), dtype=torch.float16), bias).float().cpu()
    torch.cuda.synchronize()
    assert (a - b).abs().max().item() <= 2.0 ** -9 * a.abs().max().item()


UP_CASES = [
    # (batch, full-res dims, c1 (up-sampled source), c2 (skip), max_ctas)
    (1, (8, 16, 8), 64, 64, 0),
    (2, (12, 36, 20), 128, 64, 0),
    (1, (16, 32, 48), 64, 64, 3),
    (1, (6, 18, 10), 192, 64, 2),
    (1, (40, 20, 24), 64, 128, 5),
]


@pytest.mark.parametrize("case", range(len(UP_CASES)))
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_conv_fused_upsample(cuda, lib, case, dt):
    """Decoder stage (med3d.py:83-89): conv3x3x3(cat([upsample_x2_trilinear_ac(x), skip])).  The plane-ring
    kernel that up-samples its first source on the fly must equal K4 + plain convolution bit for bit (same
    fp32 operation order, one rounding to the storage type) and match torch within the conv tolerance."""
    from dram_b200 import ops

    global DT
    DT = dt
    n, dims, c1, c2, max_ctas = UP_CASES[case]
    d, h, w = dims
    g = torch.Generator().manual_seed(60 + case)
    lo = _rand((n, c1, d // 2, h // 2, w // 2), g)
    skip = _rand((n, c2, d, h, w), g)
    wgt = _rand((64, c1 + c2, 3, 3, 3), g, scale=(27 * (c1 + c2)) ** -0.5)
    bias = torch.randn(64, generator=g) * 0.1
    lo_d, skip_d = ops.to_ndhwc_16(lo.to(cuda), dt), ops.to_ndhwc_16(skip.to(cuda), dt)
    wp = ops.pack_conv_weight(wgt, dtype=dt).to(cuda)
    fused = ops.Conv3dPlan(lo_d, wp, bias.to(cuda), x2=skip_d, upsample_x1=True)
    assert fused.algo == "planes"
    got = fused.run(max_ctas).clone()
    up = ops.upsample2x(lo_d)
    plain = ops.Conv3dPlan(up, wp, bias.to(cuda), x2=skip_d, algo="planes").run()
    torch.cuda.synchronize()
    assert torch.equal(got, plain), f"max diff {(got.float() - plain.float()).abs().max().item()}"
    up_ref = ops.to_ncdhw_f32(up).cpu()  # the 16-bit rounded up-sampled tensor the convolution sees
    ref = (F.conv3d(torch.cat([up_ref, skip], 1), wgt, None, padding=1) + bias.view(1, -1, 1, 1, 1)).relu()
    _close(ops.to_ncdhw_f32(got).cpu(), ref, f"fused upsample conv {dims} {c1}+{c2}")


def test_conv_fused_upsample_argument_errors(cuda, lib):
    from dram_b200 import ops
    from dram_b200._capi import DramError

    lo = torch.zeros((1, 4, 8, 4, 64), dtype=torch.float16, device=cuda)
    skip = torch.zeros((1, 8, 16, 8, 64), dtype=torch.float16, device=cuda)
    b = torch.zeros(64, device=cuda)
    with pytest.raises(DramError):   # no skip tensor
        ops.Conv3dPlan(lo, torch.zeros((64, 27 * 64), dtype=torch.float16, device=cuda), b, upsample_x1=True)
    with pytest.raises(DramError):   # cout 128 is not a plane-ring shape
        ops.Conv3dPlan(lo, torch.zeros((128, 27 * 128), dtype=torch.float16, device=cuda),
                       torch.zeros(128, device=cuda), x2=skip, upsample_x1=True)


STAGED_CASES = [
    # (batch, dims, cin, cout, residual channels, max_ctas): the 1x1x1 convolutions of the bottleneck blocks
    (1, (8, 8, 8), 64, 256, 256, 0),        # layer1.x.conv3: expand + full residual
    (2, (5, 9, 12), 256, 64, 0, 0),         # conv1: reduce, ragged volume (partial tiles, TMA store clipping)
    (1, (6, 10, 20), 128, 512, 512, 3),     # several tiles per CTA: residual prefetch / store double buffering
    (1, (4, 8, 8), 512, 128, 0, 1),         # one CTA walks every tile, two column groups per tile
    (1, (3, 5, 7), 64, 64, 64, 2),          # single 64-channel group per tile
    (1, (6, 9, 10), 64, 256, 64, 0),        # layer1.0.conv3: shortcut type A, residual has fewer channels
    (2, (5, 8, 12), 256, 512, 256, 4),      # layer3.0-like: shortcut A through the residual tensor map
]


@pytest.mark.parametrize("case", range(len(STAGED_CASES)))
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_conv_1x1_staged_epilogue(cuda, lib, case, dt):
    """Bottleneck 1x1x1 convolutions (med3d.py:152-157, 164-184) through the shared-memory + TMA epilogue."""
    n, dims, cin, cout, res, max_ctas = STAGED_CASES[case]
    _run_conv(cuda, n, dims, cin, 0, cout, 1, 1, 1, residual=res or None, seed=80 + case, dtype=dt, max_ctas=max_ctas,
              normalize=(case % 2 == 0))


def test_conv_1x1_staged_equals_direct_epilogue(cuda, lib, monkeypatch):
    """Both epilogues of the tile kernel round the same fp32 values once: identical 16-bit results."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(91)
    x = ops.to_ndhwc_16(_rand((1, 128, 6, 9, 11), g, dtype=torch.float16).to(cuda), torch.float16)
    res = ops.to_ndhwc_16(_rand((1, 256, 6, 9, 11), g, dtype=torch.float16).to(cuda), torch.float16)
    w = ops.pack_conv_weight(_rand((256, 128, 1, 1, 1), g, scale=128 ** -0.5), dtype=torch.float16).to(cuda)
    b = (torch.randn(256, generator=g) * 0.1).to(cuda)
    staged = ops.Conv3dPlan(x, w, b, kernel=1, residual=res).run().clone()
    monkeypatch.setenv("DRAM_B200_STAGED_EPILOGUE", "0")
    direct = ops.Conv3dPlan(x, w, b, kernel=1, residual=res).run().clone()
    torch.cuda.synchronize()
    assert torch.equal(staged, direct)


def test_conv_1x1_staged_strided_shortcut_a(cuda, lib):
    """R50 layer2.0.conv3 (med3d.py:103-112, 181): 1x1x1 expand whose residual is the stride-2 subsample of a
    tensor with fewer channels — through the staged epilogue's residual tensor map (element strides 2, channel
    groups beyond res_c out of bounds = zeros)."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(93)
    for dt in (torch.float16, torch.bfloat16):
        x = _rand((2, 128, 5, 6, 9), g, dtype=dt)
        res = _rand((2, 256, 10, 12, 18), g, dtype=dt)
        wgt = _rand((512, 128, 1, 1, 1), g, scale=128 ** -0.5, dtype=dt)
        bias = torch.randn(512, generator=g) * 0.1
        ref = F.conv3d(x, wgt, None) + bias.view(1, -1, 1, 1, 1)
        ref[:, :256] += res[:, :, ::2, ::2, ::2]
        ref = ref.relu()
        plan = ops.Conv3dPlan(ops.to_ndhwc_16(x.to(cuda), dt), ops.pack_conv_weight(wgt, dtype=dt).to(cuda),
                              bias.to(cuda), kernel=1, residual=ops.to_ndhwc_16(res.to(cuda), dt), res_stride=2)
        got = ops.to_ncdhw_f32(plan.run(3)).cpu()
        global DT
        DT = dt
        _close(got, ref, f"1x1 + strided shortcut A {dt}")


PAIR_CASES = [
    # (batch, dims, cin, cout, dilation, residual, max_ctas): wide layers with a long K loop -> M256 kernel
    (1, (8, 8, 8), 256, 256, 2, 256, 0),      # layer3 block: 4 bricks -> 2 pairs
    (1, (4, 12, 8), 256, 512, 4, 0, 0),       # 3 bricks: the last pair has one brick past the end
    (2, (6, 10, 10), 512, 512, 4, 512, 3),    # ragged bricks, two n-tiles, several tiles per CTA, batch 2
    (1, (16, 8, 8), 256, 256, 1, 0, 1),       # one CTA walks every tile: accumulator hand-over across tiles
]


@pytest.mark.parametrize("case", range(len(PAIR_CASES)))
def test_conv_pair_kernel(cuda, lib, case, monkeypatch):
    """M256 tile kernel (two bricks per weight tile) against torch, and bit-equal to the M128 tile kernel."""
    n, dims, cin, cout, dil, res, max_ctas = PAIR_CASES[case]
    dt = torch.float16 if case % 2 == 0 else torch.bfloat16
    monkeypatch.setenv("DRAM_B200_PAIR_TILES", "1")   # small test shapes: force what the heuristic keeps for big ones
    plan = _run_conv(cuda, n, dims, cin, 0, cout, 3, 1, dil, residual=res or None, seed=120 + case, dtype=dt,
                     max_ctas=max_ctas)
    assert plan.stages == 3 and plan.algo == "tiles"          # the pair kernel's stage count
    a = plan.run(max_ctas).clone()
    monkeypatch.setenv("DRAM_B200_PAIR_TILES", "0")
    plan128 = _run_conv(cuda, n, dims, cin, 0, cout, 3, 1, dil, residual=res or None, seed=120 + case, dtype=dt)
    assert plan128.stages == 4
    torch.cuda.synchronize()
    assert torch.equal(a, plan128.run())


# ------------------------------------------------------------------------------------------------
# K13: the commuted convolution of an up-sampled tensor (us1.0)
# ------------------------------------------------------------------------------------------------
def _lerp_matrix(l_lo, l_hi):
    """[l_hi, l_lo] matrix of ATen's linear interpolation with align_corners=True."""
    eye = torch.eye(l_lo).view(1, l_lo, l_lo)  # [N=1, C=l_lo, L=l_lo]
    return F.interpolate(eye, size=l_hi, mode="linear", align_corners=True)[0].t().contiguous()


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("dims,axis,groups,pad_c", [((3, 4, 5), 3, 9, 0), ((3, 4, 10), 2, 3, 0), ((3, 8, 10), 1, 1, 0),
                                                    ((2, 5, 3), 3, 9, 64), ((1, 1, 2), 2, 3, 0), ((4, 1, 1), 1, 1, 0)])
def test_upconv_axis_pass(cuda, lib, dims, axis, groups, pad_c, dt):
    """One axis of K13 against its definition: out[o, g, c] = sum_t (M x[.., g, t, c])[o + t - 1] with M ATen's
    align_corners interpolation matrix and zero for positions outside the doubled axis."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(31)
    n = 2
    c = groups * 192 + pad_c
    x = torch.randn((n,) + dims + (c,), generator=g).to(dt)
    got = ops.upconv_axis(x.to(cuda), axis, groups).float().cpu()
    l_lo = dims[axis - 1]
    M = _lerp_matrix(l_lo, 2 * l_lo)                                    # [l_hi, l_lo]
    xv = x.float()[..., :groups * 192].reshape((n,) + dims + (groups, 3, 64)).movedim(axis, -1)   # [..., g, t, c, L_lo]
    up = torch.einsum("...l,ol->...o", xv, M)                            # [..., g, t, c, L_hi]
    ref = torch.zeros(up.shape[:-3] + (64, 2 * l_lo))
    for t in range(3):
        lo, hi = max(0, 1 - t), min(2 * l_lo, 2 * l_lo + 1 - t)           # o with 0 <= o + t - 1 < l_hi
        ref[..., lo:hi] += up[..., t, :, lo + t - 1:hi + t - 1]
    ref = ref.movedim(-1, axis)                                          # [n, d, h, w, g, c] with the axis doubled
    ref = ref.reshape(ref.shape[:4] + (groups * 64,))
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 2 * ulp * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("lo_dims,c_up,n", [((4, 4, 4), 128, 1), ((3, 5, 4), 512, 2), ((8, 4, 6), 256, 1)])
def test_commuted_upsampled_conv_equals_direct(cuda, lib, lo_dims, c_up, n, dt):
    """The whole of K13 — low-resolution 1x1x1 product with 27*64 outputs, three gather passes, 64->64 convolution of the
    skip tensor with the gathered tensor as residual — against conv3d(cat[upsample(x4), x1]) in fp32 on the CPU."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(32)
    hi_dims = tuple(2 * v for v in lo_dims)
    x4 = _rand((n, c_up) + lo_dims, g, dtype=dt)
    x1 = _rand((n, 64) + hi_dims, g, dtype=dt)
    wgt = torch.randn((64, c_up + 64, 3, 3, 3), generator=g) * ((c_up + 64) * 27) ** -0.5
    scale = torch.rand(64, generator=g) + 0.5
    shift = torch.randn(64, generator=g) * 0.1
    up = F.interpolate(x4, scale_factor=2, mode="trilinear", align_corners=True)
    ref = torch.relu(F.conv3d(torch.cat([up, x1], 1), wgt, None, padding=1) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    x4d, x1d = ops.to_ndhwc_16(x4.to(cuda), dt), ops.to_ndhwc_16(x1.to(cuda), dt)
    wz, mz = ops.pack_upconv_weight(wgt.to(cuda), c_up, scale.to(cuda), dtype=dt)
    assert wz.shape == (1792, c_up) and not bool(wz[1728:].any())    # padded to 7 N-tiles of 256 with zero rows
    zplan = ops.Conv3dPlan(x4d, wz, torch.zeros(wz.shape[0], device=cuda), scale=mz, kernel=1, relu=False, epilogue="staged")
    assert zplan.block_n == 256    # the staged epilogue without residual tiles keeps the 256-wide N tile
    z = zplan.run()
    zd = ops.Conv3dPlan(x4d, wz, torch.zeros(wz.shape[0], device=cuda), scale=mz, kernel=1, relu=False, epilogue="direct").run()
    assert torch.equal(z, zd)      # staged and direct epilogues: same arithmetic, same bits
    gath = ops.upconv_axis(ops.upconv_axis(ops.upconv_axis(z, 3, 9), 2, 3), 1, 1)
    ws, ms = ops.pack_conv_weight(wgt.to(cuda)[:, c_up:], scale.to(cuda), dtype=dt, normalize=True)
    out = ops.Conv3dPlan(x1d, ws, shift.to(cuda), scale=ms, residual=gath).run()
    got = ops.to_ncdhw_f32(out).cpu()
    err = (got - ref).abs()
    # rounding points: z, two intermediate gathers, the gathered tensor, the output (direct route: up(x4), the output)
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    assert err.max().item() <= 6 * ulp * ref.abs().max().item(), (err.max().item(), ref.abs().max().item())


@pytest.mark.parametrize("c1,c2,cout,k,dil,dims", [(64, 0, 64, 3, 1, (24, 64, 64)), (64, 64, 64, 3, 1, (16, 48, 64)),
                                                   (128, 0, 128, 3, 1, (32, 32, 32)),
                                                   (64, 0, 32, 3, 1, (24, 64, 64)), (128, 0, 256, 3, 2, (16, 16, 16)),
                                                   (256, 0, 128, 1, 1, (16, 16, 32))])
def test_convolutions_are_reproducible_bit_for_bit(cuda, lib, c1, c2, cout, k, dil, dims):
    """Same buffers, same plan, six launches: identical bits — on shapes large enough to keep every SM busy.  Guards the
    issue order of the plane-ring kernel (K steps outermost, planes interleaved: instructions with partially overlapping
    accumulator ranges in flight) and every other ordering assumption of the convolution kernels."""
    from dram_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(41)
    x1 = torch.randn((2,) + dims + (c1,), generator=g, device=cuda).half()
    x2 = torch.randn((2,) + dims + (c2,), generator=g, device=cuda).half() if c2 else None
    w = (torch.randn((cout, k ** 3 * (c1 + c2)), generator=g, device=cuda) * 0.02).half()
    heads = (torch.randn(2, 32, device=cuda), torch.zeros(2, device=cuda), (1, 1), True) if cout == 32 else None
    plan = ops.Conv3dPlan(x1, w, torch.zeros(cout, device=cuda), x2=x2, kernel=k, dilation=dil, heads=heads,
                          store_out=heads is None)
    outs = []
    for _ in range(6):
        plan.run()
        outs.append((plan.head_outs[0] if heads else plan.out).clone())
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:])


def test_stem_is_reproducible_bit_for_bit(cuda, lib):
    from dram_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.randn((2, 48, 96, 128), generator=g, device=cuda)
    w = torch.randn((64, 1, 7, 7, 7), generator=g, device=cuda) * 0.05
    packed, mult = ops.pack_stem_weight_fused(w, dtype=torch.float16, normalize=True)
    bias = torch.zeros(64, device=cuda)
    outs = [ops.stem_conv7(x, packed, bias, mult).clone() for _ in range(6)]
    assert all(torch.equal(outs[0], o) for o in outs[1:])
