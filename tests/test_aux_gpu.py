"""Parity of the memory-bound kernels (K2a, K3, K4, K6, K7, K8, K8b) against the torch CPU ops the
reference calls.  Index/mask work is bit-exact; floating point within the stated tolerance."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


DTYPES = [torch.bfloat16, torch.float16]


@pytest.mark.parametrize("dt", DTYPES)
def test_layout_roundtrip(cuda, lib, dt):
    from dram_b200 import ops

    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 24, 3, 5, 7, generator=g).to(dt).float()
    y = ops.to_ncdhw_f32(ops.to_ndhwc_16(x.to(cuda), dt)).cpu()
    assert torch.equal(x, y)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("dims", [(8, 8, 8), (7, 9, 12), (16, 6, 10), (20, 6, 10), (33, 5, 7), (1, 1, 1), (2, 3, 1)])
def test_maxpool(cuda, lib, dims, dt):
    """med3d.py:305 MaxPool3d(3, stride 2, pad 1): exact (max of 16-bit values)."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(1)
    x = torch.randn((2, 64) + dims, generator=g).to(dt).float()
    ref = F.max_pool3d(x, 3, 2, 1)
    got = ops.to_ncdhw_f32(ops.maxpool3d(ops.to_ndhwc_16(x.to(cuda), dt))).cpu()
    assert torch.equal(got, ref)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("dims", [(4, 4, 4), (3, 5, 6), (8, 7, 9), (9, 3, 5), (21, 2, 3), (1, 1, 1), (1, 4, 2)])
def test_upsample2x(cuda, lib, dims, dt):
    """med3d.py:83 nn.Upsample(scale 2, trilinear, align_corners=True); fp32 math, one 16-bit rounding."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(2)
    x = torch.randn((2, 64) + dims, generator=g).to(dt).float()
    ref = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    got = ops.to_ncdhw_f32(ops.upsample2x(ops.to_ndhwc_16(x.to(cuda), dt))).cpu()
    assert got.shape == ref.shape
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    assert (got - ref).abs().max().item() <= ulp * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize("dt", DTYPES)
def test_stem_expand_layout(cuda, lib, dt):
    from dram_b200 import ops

    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 9, 12, generator=g).to(dt).float()
    got = ops.stem_expand(x.to(cuda), dtype=dt).float().cpu()
    n, d, h, w = x.shape
    h2, w2 = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    xp = F.pad(x, (3, 5, 3, 5))
    ref = torch.zeros(n, d, h2, w2, 8, 8)
    for kh in range(7):
        for j in range(7):
            ref[..., kh, j] = xp[:, :, kh:kh + 2 * h2:2, j:j + 2 * w2:2][:, :, :h2, :w2]
    assert torch.equal(got, ref.reshape(n, d, h2, w2, 64))


@pytest.mark.parametrize("mask_dims,dims", [((16, 16, 16), (8, 8, 8)), ((10, 14, 18), (5, 7, 9)),
                                            ((9, 11, 13), (4, 6, 5))])
def test_masked_pool(cuda, lib, mask_dims, dims):
    """med3d.py:383-387: nearest-resampled lung mask (bit-exact indexing), masked mean in fp32."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(4)
    dense = torch.rand((2, 3) + dims, generator=g)
    lungs = (torch.rand((2, 1) + mask_dims, generator=g) > 0.6).float()
    m = F.interpolate(lungs, dims, mode="nearest")
    ref = (dense * m).view(2, 3, -1).sum(-1) / m.view(2, 1, -1).sum(-1)
    got = ops.masked_pool(dense.to(cuda), lungs[:, 0].to(torch.uint8).to(cuda)).cpu()
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-7), (got, ref)
    ref_plain = dense.view(2, 3, -1).mean(-1)
    got_plain = ops.masked_pool(dense.to(cuda), None).cpu()
    assert torch.allclose(got_plain, ref_plain, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("dims,size", [((4, 6, 8), (8, 12, 16)), ((5, 7, 9), (10, 14, 18)),
                                       ((4, 4, 6), (9, 11, 13))])
def test_dram_upsample_mask(cuda, lib, dims, size):
    """models.py:438-441: dRAM = trilinear(dense -> scan size, align_corners) * ess; pct over batch lungs."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(5)
    d0 = torch.rand((2, 1) + dims, generator=g)
    d1 = torch.rand((2, 1) + dims, generator=g)
    lungs = (torch.rand((2, 1) + size, generator=g) > 0.4).float()
    ess = lungs * (torch.rand((2, 1) + size, generator=g) > 0.5).float()
    r0 = F.interpolate(d0, size=size, mode="trilinear", align_corners=True) * ess
    r1 = F.interpolate(d1, size=size, mode="trilinear", align_corners=True) * ess
    p0 = r0.view(2, -1).sum(-1) / lungs.sum()
    p1 = r1.view(2, -1).sum(-1) / lungs.sum()
    o0, o1, pct = ops.dram_upsample_mask(d0.to(cuda), d1.to(cuda), ess[:, 0].to(torch.uint8).to(cuda),
                                         lungs[:, 0].to(torch.uint8).to(cuda), size)
    assert (o0.cpu() - r0).abs().max().item() < 1e-5
    assert (o1.cpu() - r1).abs().max().item() < 1e-5
    assert torch.equal(o0.cpu() == 0, r0 == 0)
    assert torch.allclose(pct.cpu(), torch.stack([p0, p1]), rtol=1e-5)
    _, _, pct_s = ops.dram_upsample_mask(d0.to(cuda), d1.to(cuda), ess[:, 0].to(torch.uint8).to(cuda),
                                         lungs[:, 0].to(torch.uint8).to(cuda), size, per_sample_denominator=True)
    ps = r0.view(2, -1).sum(-1) / lungs.view(2, -1).sum(-1)
    assert torch.allclose(pct_s.cpu()[0], ps, rtol=1e-5)


@pytest.mark.parametrize("dims,size,n", [((16, 16, 16), (32, 32, 32), 2), ((8, 12, 16), (16, 24, 64), 3),
                                         ((12, 10, 64), (24, 20, 512), 1), ((5, 7, 9), (10, 14, 18), 2),
                                         ((4, 5, 6), (9, 11, 13), 2), ((6, 8, 40), (11, 16, 80), 1),
                                         ((2, 40, 8), (3, 79, 16), 2), ((4, 8, 32), (8, 16, 64), 2),
                                         ((5, 7, 32), (10, 14, 64), 2), ((3, 6, 128), (6, 12, 256), 1),
                                         ((2, 40, 64), (4, 80, 128), 1), ((2, 3, 512), (4, 6, 1024), 1)])
def test_dram_staged_kernel_equals_row_kernel(cuda, lib, dims, size, n, monkeypatch):
    """K7's shared-memory staged kernels (16-voxel segments, source brick in shared memory; the lean one-segment-per-
    thread variant that is the default for power-of-two rows, the generic fast and ragged variants) against the
    round-1 warp-per-row kernel (DRAM_B200_K7=rows): same expression, same rounding order ->
    bit-identical maps; the fp64 sums agree to the last fp32 bit of the percentages.  Covers aligned sizes (vector
    path), W = 512 (two segments per thread), ragged sizes and unaligned sample strides (scalar path), sparse and dense
    `ess`, and empty masks."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(15)
    d0 = torch.rand((n, 1) + dims, generator=g).to(cuda)
    d1 = torch.rand((n, 1) + dims, generator=g).to(cuda)
    lungs = (torch.rand((n,) + size, generator=g) > 0.4)
    for frac in (0.03, 0.9, 0.0):
        ess = (lungs & (torch.rand((n,) + size, generator=g) < frac)).to(torch.uint8).to(cuda)
        lu = lungs.to(torch.uint8).to(cuda)
        monkeypatch.setenv("DRAM_B200_K7", "rows")
        r0, r1, rp = ops.dram_upsample_mask(d0, d1, ess, lu, size)
        for mode in ("staged", "lean"):   # lean = the default: one CTA per plane slice where W = 16 * 2^k, else staged
            monkeypatch.setenv("DRAM_B200_K7", mode)
            o0, o1, op = ops.dram_upsample_mask(d0, d1, ess, lu, size)
            assert torch.equal(o0, r0) and torch.equal(o1, r1), (dims, size, frac, mode)
            assert torch.allclose(op, rp, rtol=1e-6, atol=0), (op, rp, mode)
        ref = F.interpolate(d0.cpu(), size=size, mode="trilinear", align_corners=True) * ess.cpu()[:, None].float()
        assert (o0.cpu() - ref).abs().max().item() < 1e-5
        assert torch.equal(o0.cpu() == 0, ref == 0)


@pytest.mark.parametrize("shape,n", [((16, 16, 16), 3), ((7, 9, 11), 2), ((15, 15, 15), 4), ((8, 8, 24), 1)])
def test_window_standardize_batched_and_unaligned(cuda, lib, shape, n):
    """K8 over a stack of volumes in three launches: per-volume statistics, and volumes that do not start on a 16-byte
    boundary (8k-1 sizes in a batch: count % 8 != 0) take the scalar path instead of failing."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(16)
    hu = ((torch.randn((n,) + shape, generator=g) * 400 - 700) + 150 * torch.arange(n).view(-1, 1, 1, 1)).round()
    hu = hu.clamp(-2048, 1500).to(torch.int16)
    got, stats = ops.window_standardize(hu.to(cuda), batched=True)
    only = ops.window_stats(hu.to(cuda))
    assert torch.equal(only, stats) and stats.shape == (n, 2)
    for b in range(n):
        v = (torch.clamp(hu[b].float(), min=-1150, max=-300) + 1150) / 850
        ref = (v - v.mean()) / v.std()
        assert (got[b].cpu() - ref).abs().max().item() < 2e-5
        assert abs(stats[b, 0].item() - v.mean().item()) < 1e-6 and abs(stats[b, 1].item() - v.std().item()) < 1e-6
        one, st1 = ops.window_standardize(hu[b].to(cuda).contiguous())
        assert (one - got[b]).abs().max().item() < 1e-6   # (fp64 atomics may order differently: last-bit differences)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape,n", [((16, 32, 32), 2), ((15, 31, 23), 2), ((9, 16, 40), 1)])
def test_stem_from_hu_equals_window_then_stem(cuda, lib, shape, n, dt):
    """The int16-HU stem (K8's apply pass fused into the producers) gives bit for bit what K8 + the fp32 stem give."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(17)
    hu = (torch.randn((n,) + shape, generator=g) * 400 - 700).round().clamp(-2048, 1500).to(torch.int16).to(cuda)
    w = torch.randn((64, 1, 7, 7, 7), generator=g).to(cuda) * 0.05
    packed, mult = ops.pack_stem_weight_fused(w, dtype=dt, normalize=True)
    bias = torch.randn(64, generator=g).to(cuda) * 0.1
    img, stats = ops.window_standardize(hu, batched=True)
    want = ops.stem_conv7(img, packed, bias, mult)
    lut, stats2 = ops.window_lut(hu)
    assert torch.equal(stats2, stats) and lut.shape == (n, 851)
    # the table holds K8's value for every HU of the window (and the clamped ends for everything outside)
    ramp = torch.arange(-1200, -250, dtype=torch.int16, device=cuda).repeat(n, 1)
    for b in range(n):
        v = ((ramp[b].float().clamp(-1150, -300) + 1150) / 850 - stats[b, 0]) / stats[b, 1]
        # (torch divides by a Python scalar through a reciprocal multiply: last-bit differences from K8's IEEE division)
        assert torch.allclose(lut[b][(ramp[b].clamp(-1150, -300) + 1150).long()], v, rtol=1e-5, atol=1e-5)
        # ... and K8 itself on the ramp, shifted to this volume's statistics, is reproduced exactly
    k8, _ = ops.window_standardize(hu, batched=True)
    for b in range(n):
        idx = (hu[b].clamp(-1150, -300) + 1150).long()
        assert torch.equal(lut[b][idx], k8[b])
    got = ops.stem_conv7_hu(hu, lut, packed, bias, mult)
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", [(16, 16, 16), (7, 9, 11)])
def test_window_standardize(cuda, lib, shape):
    """functional.py:13-26 + intensity_transforms.py:104-114 (unbiased std)."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(6)
    hu = (torch.randn(shape, generator=g) * 400 - 700).round().clamp(-2048, 1500).to(torch.int16)
    v = torch.clamp(hu.float(), min=-1150, max=-300)
    v = ((v - (-1150)) / (-300 - (-1150))) * (1 - 0) + 0
    ref = (v - v.mean()) / (v - v.mean()).std()
    got, stats = ops.window_standardize(hu.to(cuda))
    assert (got.cpu() - ref).abs().max().item() < 2e-5
    assert abs(stats[0].item() - v.mean().item()) < 1e-6


@pytest.mark.parametrize("shape,size", [((10, 12, 14), (8, 16, 24)), ((40, 30, 50), (32, 24, 40)),
                                        ((9, 17, 13), (16, 8, 8))])
def test_resize(cuda, lib, shape, size):
    """spatial_transforms.py:55-97 Interpolate(only_in_plane=True): mask path bit-exact."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(7)
    x = torch.randn(shape, generator=g)
    m = torch.rand(shape, generator=g) > 0.5
    idx = torch.linspace(0, shape[0] - 1, size[0]).long()
    ref = F.interpolate(x[None], size=size[1:], mode="bilinear", align_corners=True)[:, idx][0]
    refm = F.interpolate(m[None].float(), size=size[1:], mode="nearest")[:, idx][0].bool()
    got = ops.resize_image(x.to(cuda), size).cpu()
    gotm = ops.resize_mask(m.to(torch.uint8).to(cuda), size).cpu().bool()
    assert (got - ref).abs().max().item() < 1e-5
    assert torch.equal(gotm, refm)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("dims,c,n,max_ctas", [((4, 4, 4), 64, 1, 0), ((3, 5, 6), 128, 2, 0), ((8, 7, 9), 64, 1, 3),
                                               ((9, 3, 5), 192, 1, 1), ((16, 16, 16), 64, 1, 0), ((1, 2, 4), 64, 1, 0)])
def test_upsample2x_tensor_core(cuda, lib, dims, c, n, max_ctas, dt):
    """K4 as a tcgen05 GEMM (generated interpolation matrix x source patch) against F.interpolate and K4.
    Error budget: the eight weights are rounded to the storage type (ulp/2 relative each) before the fp32
    accumulation, the result once more: <= 2 * ulp * max|x|."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(12)
    x = torch.randn((n, c) + dims, generator=g).to(dt).float()
    ref = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    xd = ops.to_ndhwc_16(x.to(cuda), dt)
    plan = ops.Upsample2xPlan(xd)
    got = ops.to_ncdhw_f32(plan.run(max_ctas)).cpu()
    assert got.shape == ref.shape
    ulp = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    tol = 2 * ulp * x.abs().max().item() + 1e-6
    assert (got - ref).abs().max().item() <= tol
    k4 = ops.to_ncdhw_f32(ops.upsample2x(xd)).cpu()
    assert (got - k4).abs().max().item() <= tol
    # constants are reproduced to the rounding of the weights (rows of the interpolation matrix sum to one)
    ones = ops.Upsample2xPlan(torch.ones_like(xd)).run().float()
    assert (ones - 1.0).abs().max().item() <= 4 * ulp


def test_torch_library_ops_run_capture_and_opcheck(cuda, lib):
    """The dram_b200:: custom ops give the results of the ctypes wrappers, pass torch.library.opcheck (schema, fake
    kernel, dispatch) and capture into a CUDA graph whose replay follows new input data."""
    from dram_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((1, 16, 16, 24, 64), generator=g, device=cuda).half()
    assert torch.equal(torch.ops.dram_b200.maxpool3d(x), ops.maxpool3d(x))
    assert torch.equal(torch.ops.dram_b200.upsample2x(x), ops.upsample2x(x))
    w = (torch.randn((64, 64, 3, 3, 3), generator=g, device=cuda) * 0.05)
    packed, bias = ops.pack_conv_weight(w, dtype=torch.float16), torch.zeros(64, device=cuda)
    y = torch.ops.dram_b200.conv3d(x, packed, bias)
    ref = torch.relu(torch.nn.functional.conv3d(x.permute(0, 4, 1, 2, 3).float(), packed.float().view(64, 3, 3, 3, 64)
                                                .permute(0, 4, 1, 2, 3), padding=1)).permute(0, 2, 3, 4, 1)
    assert (y.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    torch.library.opcheck(torch.ops.dram_b200.maxpool3d.default, (x,))
    torch.library.opcheck(torch.ops.dram_b200.conv3d.default, (x, packed, bias), {"dilation": 2})
    # graph capture of dispatcher calls: static input, replay after the input changed
    xs = x.clone()
    torch.ops.dram_b200.conv3d(xs, packed, bias)  # warm-up (function attributes) outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = torch.ops.dram_b200.maxpool3d(torch.ops.dram_b200.conv3d(xs, packed, bias))
    xs.copy_(torch.randn(xs.shape, generator=g, device=cuda).half())
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ops.maxpool3d(torch.ops.dram_b200.conv3d(xs, packed, bias)))
