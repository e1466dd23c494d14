"""Out-of-bounds WRITE check with guard bands.  compute-sanitizer is closed on this GPU pool (the refusal is committed as
profiles/sanitizer_closed_r2g.txt), so this is the bounds check the suite carries itself: every kernel family writes
into a view that sits in the middle of a larger allocation whose borders hold a sentinel pattern; after the launch
the borders must be untouched.  Shapes are ragged on purpose (partial tiles, clipped TMA boxes, odd sizes, tails of
vector loops) — the places where an index slip writes past a buffer."""
import pytest
import torch

pytestmark = pytest.mark.gpu
PAD = 4096  # bytes on either side (keeps 16-byte / 1 KiB alignment of the view)


class Guarded:
    def __init__(self, shape, dtype, device):
        self.numel = 1
        for v in shape:
            self.numel *= v
        self.item = torch.empty((), dtype=dtype).element_size()
        self.raw = torch.full((2 * PAD + self.numel * self.item,), 0xA5, dtype=torch.uint8, device=device)
        self.view = self.raw[PAD:PAD + self.numel * self.item].view(dtype).view(shape)

    def check(self, what):
        torch.cuda.synchronize()
        lo, hi = self.raw[:PAD], self.raw[PAD + self.numel * self.item:]
        assert bool((lo == 0xA5).all()), f"{what}: wrote BEFORE its output buffer"
        assert bool((hi == 0xA5).all()), f"{what}: wrote PAST its output buffer"
        return self.view


def _conv(cuda, dims, c1, cout, k=3, stride=1, dil=1, c2=0, res=False, dt=torch.float16, epilogue="auto", n=1):
    from dram_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(7)
    x1 = torch.randn((n,) + dims + (c1,), generator=g, device=cuda).to(dt)
    x2 = torch.randn((n,) + dims + (c2,), generator=g, device=cuda).to(dt) if c2 else None
    w = (torch.randn((cout, k ** 3 * (c1 + c2)), generator=g, device=cuda) * 0.02).to(dt)
    probe = ops.Conv3dPlan(x1, w, torch.zeros(cout, device=cuda), x2=x2, kernel=k, stride=stride, dilation=dil, epilogue=epilogue)
    out = Guarded(probe.out_shape, dt, cuda)
    r = torch.randn(probe.out_shape, generator=g, device=cuda).to(dt) if res else None
    plan = ops.Conv3dPlan(x1, w, torch.zeros(cout, device=cuda), x2=x2, kernel=k, stride=stride, dilation=dil,
                          residual=r, out=out.view, epilogue=epilogue)
    plan.run()
    got = out.check(f"conv {c1}+{c2}->{cout} k{k} s{stride} d{dil} {dims} [{plan.algo}, {epilogue}]")
    assert bool(torch.isfinite(got.float()).all())


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_convolution_kernels_stay_inside_their_output(cuda, lib, dt):
    _conv(cuda, (5, 9, 11), 64, 64, dt=dt)                      # plane ring, ragged slabs
    _conv(cuda, (7, 17, 9), 64, 32, dt=dt)                      # plane ring N = 32 (output stored)
    _conv(cuda, (6, 10, 14), 64, 64, c2=64, res=True, dt=dt)    # two sources + residual
    _conv(cuda, (5, 7, 9), 128, 128, dt=dt)                     # plane ring N = 128 (two planes per item), ragged
    _conv(cuda, (5, 7, 9), 128, 128, dil=2, dt=dt)              # tiles N = 128, partial tiles
    _conv(cuda, (9, 7, 5), 64, 128, stride=2, dt=dt)            # stride 2
    _conv(cuda, (6, 6, 10), 128, 256, dil=2, dt=dt)             # N = 256, dilation (tap skipping)
    _conv(cuda, (5, 6, 7), 256, 512, dil=4, res=True, dt=dt, n=2)
    _conv(cuda, (5, 7, 9), 64, 256, k=1, res=True, dt=dt)       # 1x1x1 staged epilogue (TMA stores clip the tile)
    _conv(cuda, (5, 7, 9), 128, 256, k=1, dt=dt, epilogue="staged")   # staged, N = 256, no residual
    _conv(cuda, (5, 7, 9), 128, 1792, k=1, dt=dt, epilogue="staged")  # K13's product shape
    _conv(cuda, (5, 7, 9), 128, 256, k=1, dt=dt, epilogue="direct")


@pytest.mark.parametrize("dims", [(9, 17, 23), (16, 32, 30), (7, 8, 8)])
def test_stem_and_pooling_stay_inside_their_output(cuda, lib, dims):
    from dram_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(8)
    n = 2
    x = torch.randn((n,) + dims, generator=g, device=cuda)
    w = torch.randn((64, 1, 7, 7, 7), generator=g, device=cuda) * 0.05
    packed, mult = ops.pack_stem_weight_fused(w, dtype=torch.float16, normalize=True)
    half = tuple((v - 1) // 2 + 1 for v in dims)
    out = Guarded((n,) + half + (64,), torch.float16, cuda)
    ops.stem_conv7(x, packed, torch.zeros(64, device=cuda), mult, out=out.view)
    stem = out.check(f"stem {dims}")
    hu = (x * 300 - 700).round().clamp(-2000, 1500).to(torch.int16)
    lut, _ = ops.window_lut(hu)
    out = Guarded((n,) + half + (64,), torch.float16, cuda)
    ops.stem_conv7_hu(hu, lut, packed, torch.zeros(64, device=cuda), mult, out=out.view)
    out.check(f"stem from HU {dims}")
    quarter = tuple((v - 1) // 2 + 1 for v in half)
    out = Guarded((n,) + quarter + (64,), torch.float16, cuda)
    ops.maxpool3d(stem.contiguous(), out=out.view)
    pooled = out.check(f"maxpool {half}")
    up_shape = (n,) + tuple(2 * v for v in quarter) + (64,)
    out = Guarded(up_shape, torch.float16, cuda)
    ops.upsample2x(pooled.contiguous(), out=out.view)
    out.check("upsample2x (CUDA cores)")
    out = Guarded(up_shape, torch.float16, cuda)
    ops.Upsample2xPlan(pooled.contiguous(), out=out.view).run()
    out.check("upsample2x (tensor cores, TMA stores)")
    out = Guarded((n,) + dims, torch.float32, cuda)
    ops.window_standardize(hu, out=out.view, batched=True)
    out.check("window_standardize")


@pytest.mark.parametrize("lo,axis,groups", [((3, 5, 7), 3, 9), ((3, 5, 14), 2, 3), ((3, 10, 14), 1, 1)])
def test_upconv_passes_stay_inside_their_output(cuda, lib, lo, axis, groups):
    from dram_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn((2,) + lo + (groups * 192,), generator=g, device=cuda).half()
    shape = [2] + list(lo) + [groups * 64]
    shape[axis] *= 2
    out = Guarded(tuple(shape), torch.float16, cuda)
    ops.upconv_axis(x, axis, groups, out=out.view)
    out.check(f"upconv axis {axis}")


@pytest.mark.parametrize("dims,size", [((4, 5, 32), (8, 10, 64)), ((5, 7, 9), (10, 14, 18)), ((3, 9, 128), (6, 18, 256)),
                                       ((4, 5, 6), (9, 11, 13))])
def test_dram_and_heatmap_stay_inside_their_output(cuda, lib, dims, size, monkeypatch):
    from dram_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(10)
    n = 2
    d0 = torch.rand((n, 1) + dims, generator=g, device=cuda)
    d1 = torch.rand((n, 1) + dims, generator=g, device=cuda)
    lungs = (torch.rand((n,) + size, generator=g, device=cuda) > 0.4).to(torch.uint8)
    ess = lungs * (torch.rand((n,) + size, generator=g, device=cuda) < 0.2).to(torch.uint8)
    for mode in ("rows", "staged", "lean"):
        monkeypatch.setenv("DRAM_B200_K7", mode)
        o0, o1 = Guarded((n, 1) + size, torch.float32, cuda), Guarded((n, 1) + size, torch.float32, cuda)
        ops.dram_upsample_mask(d0, d1, ess, lungs, size, out=(o0.view, o1.view))
        o0.check(f"dRAM map 0 [{mode}]")
        o1.check(f"dRAM map 1 [{mode}]")
    full = tuple(v + 9 for v in size)
    crop = [(2, 2 + size[0] + 3), (4, 4 + size[1] + 1), (1, 1 + size[2] + 5)]
    out = Guarded(full, torch.uint8, cuda)
    ops.heatmap_u8(o0.view[0, 0].contiguous(), crop, full, out=out.view)
    out.check("heatmap_u8")
    scan = (torch.randn(full, generator=g, device=cuda) * 300 - 600).round().to(torch.int16)
    lobe = (torch.rand(full, generator=g, device=cuda) > 0.5).to(torch.uint8)
    cs = tuple(b - a for a, b in crop)
    gi, gl, ge = Guarded(cs, torch.int16, cuda), Guarded(cs, torch.uint8, cuda), Guarded(cs, torch.uint8, cuda)
    ops.lung_crop(scan, lobe, crop, out=(gi.view, gl.view, ge.view))
    for gb, name in ((gi, "image"), (gl, "lung"), (ge, "ess")):
        gb.check(f"lung_crop {name}")
