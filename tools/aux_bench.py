"""Times the memory-bound kernels and the fused stem at the shapes of resnet34segreg on a cubic volume.

Dev tool: prints per-kernel time and achieved GB/s against the ALGORITHMIC bytes (compulsory reads + writes).
    python tools/aux_bench.py [size] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dram_b200  # noqa: E402,F401
from dram_b200 import ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
DT = torch.float16
dev = torch.device("cuda:0")


def timeit(fn, reps=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()  # evict L2 between repetitions
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes, flops=None):
    extra = f"  {flops / ms / 1e9:8.1f} TFLOP/s" if flops else ""
    print(f"{name:34s} {ms:8.3f} ms  {nbytes / 1e6:9.1f} MB  {nbytes / ms / 1e6:8.1f} GB/s{extra}", flush=True)


def main():
    V = S ** 3
    h, q, e = S // 2, S // 4, S // 8
    img = torch.randn((B, S, S, S), device=dev)
    # K2 fused stem
    wgt = torch.randn(64, 1, 7, 7, 7, device=dev) * 343 ** -0.5
    wp, mult = ops.pack_stem_weight_fused(wgt, dtype=DT, normalize=True)
    bias = torch.zeros(64, device=dev)
    x = torch.empty((B, h, h, h, 64), dtype=DT, device=dev)
    ms = timeit(lambda: ops.stem_conv7(img, wp, bias, mult, out=x))
    report("K2 stem fused 7^3 s2 1->64", ms, B * (V * 4 + h ** 3 * 128), flops=2.0 * B * h ** 3 * 64 * 343)
    # K2a + K1 route for comparison
    xe = torch.empty((B, S, h, h, 64), dtype=DT, device=dev)
    ms = timeit(lambda: ops.stem_expand(img, out=xe))
    report("K2a stem_expand", ms, B * (V * 4 + S * h * h * 128))
    # K3 maxpool
    xp = torch.empty((B, q, q, q, 64), dtype=DT, device=dev)
    ms = timeit(lambda: ops.maxpool3d(x, out=xp))
    report("K3 maxpool 64ch /2->/4", ms, B * (h ** 3 + q ** 3) * 128)
    # K4 upsample
    x4 = torch.randn((B, e, e, e, 512), device=dev).to(DT)
    up1 = torch.empty((B, q, q, q, 512), dtype=DT, device=dev)
    ms = timeit(lambda: ops.upsample2x(x4, out=up1))
    report("K4 upsample2x 512ch /8->/4", ms, B * (e ** 3 + q ** 3) * 1024)
    x1 = torch.randn((B, q, q, q, 64), device=dev).to(DT)
    up2 = torch.empty((B, h, h, h, 64), dtype=DT, device=dev)
    ms = timeit(lambda: ops.upsample2x(x1, out=up2))
    report("K4 upsample2x 64ch /4->/2", ms, B * (q ** 3 + h ** 3) * 128)
    plan1 = ops.Upsample2xPlan(x4, out=up1)
    ms = timeit(lambda: plan1.run())
    report("K4 tensor-core 512ch /8->/4", ms, B * (e ** 3 + q ** 3) * 1024)
    plan2 = ops.Upsample2xPlan(x1, out=up2)
    ms = timeit(lambda: plan2.run())
    report("K4 tensor-core 64ch /4->/2", ms, B * (q ** 3 + h ** 3) * 128)
    # K6 pooling, K7 dRAM
    dense0 = torch.rand((B, 1, h, h, h), device=dev)
    dense1 = torch.rand((B, 1, h, h, h), device=dev)
    lungs = (torch.rand((B, S, S, S), device=dev) < 0.25).to(torch.uint8)
    ess = (lungs & (torch.rand((B, S, S, S), device=dev) < 0.3).to(torch.uint8))
    ms = timeit(lambda: ops.masked_pool(dense0, lungs))
    report("K6 masked_pool (one map)", ms, B * (h ** 3 * 4 + h ** 3))
    # realistic masks: ellipsoid lungs (24 % of the volume), ess = 8 % of the lung in blobs (bench.make_volumes)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    hu_b, lungs_b, ess_b = bench.make_volumes(B, (S, S, S), dev, seed=0)
    for tag, e_, l_ in (("random 7.5 % ess", ess, lungs), ("synthetic CT masks", ess_b, lungs_b)):
        for mode in ("rows", "staged", "lean"):
            os.environ["DRAM_B200_K7"] = mode
            ms = timeit(lambda: ops.dram_upsample_mask(dense0, dense1, e_, l_, (S, S, S)))
            report(f"K7 dRAM {mode}, {tag}", ms, B * (2 * h ** 3 * 4 + 2 * V + 2 * V * 4))
    os.environ.pop("DRAM_B200_K7", None)
    # K8 window + standardise
    hu = hu_b
    out = torch.empty((B, S, S, S), dtype=torch.float32, device=dev)
    ms = timeit(lambda: ops.window_standardize(hu, out=out, batched=True))
    report(f"K8 window+standardise ({B} vol, 3 launches)", ms, B * V * (2 + 2 + 4))
    ms = timeit(lambda: ops.window_stats(hu))
    report(f"K8 statistics pass only ({B} vol)", ms, B * V * 2)
    ms = timeit(lambda: ops.window_lut(hu))
    report(f"K8 statistics + table ({B} vol)", ms, B * V * 2)
    lut, _ = ops.window_lut(hu)
    ms = timeit(lambda: ops.stem_conv7_hu(hu, lut, wp, bias, mult, out=x))
    report("K2 stem fused from int16 HU", ms, B * (V * 2 + h ** 3 * 128), flops=2.0 * B * h ** 3 * 64 * 343)
    # f1 / f2: the device pre-/post-steps of the product path (whole volume as the crop, and an interior crop box)
    lobes = (lungs_b[0] * 3).to(torch.uint8)
    cle, _, _ = ops.dram_upsample_mask(dense0, dense1, ess_b, lungs_b, (S, S, S))
    dram0 = cle[0, 0].contiguous()
    c0, c1 = S // 8, S - S // 8 - 3
    for tag, box in (("whole volume", ((0, S), (0, S), (0, S))), ("interior crop", ((c0, c1), (c0 + 1, c1 - 2), (c0 + 3, c1)))):
        cd, chh, cww = (b - a for a, b in box)
        outs = (torch.empty((cd, chh, cww), dtype=torch.int16, device=dev),
                torch.empty((cd, chh, cww), dtype=torch.uint8, device=dev),
                torch.empty((cd, chh, cww), dtype=torch.uint8, device=dev))
        ms = timeit(lambda: ops.lung_crop(hu_b[0], lobes, box, out=outs))
        report(f"f1 lung_crop, {tag}", ms, cd * chh * cww * (2 + 1 + 2 + 1 + 1))
        heat = torch.empty((S, S, S), dtype=torch.uint8, device=dev)
        ms = timeit(lambda: ops.heatmap_u8(dram0, box, (S, S, S), out=heat))
        report(f"f2 heatmap_u8, {tag}", ms, V * 4 + V)


if __name__ == "__main__":
    main()
