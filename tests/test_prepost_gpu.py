"""the device pre-steps (lung bounding box, 2x dilation blanking, crop, LAA-910 mask)
and post-steps (resample to crop, paste, uint8 windowing) against the oracle's restatement of
dataset.py:66-83 / utils.py:53-63 (scipy binary_dilation, numpy) and processor.py:111-158.
Integer/byte results are bit-exact; the uint8 heat-map may differ by one level where the fp32 trilinear
value times 255 lands within rounding distance of an integer (the reference truncates)."""
import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as P
from oracle import synthetic

pytestmark = pytest.mark.gpu


def _lobes_case(case):
    g = np.random.default_rng(100 + case)
    if case == 0:      # synthetic lungs, isotropic
        ct, lobes = synthetic.make_volume(1, (40, 48, 56))
        return ct.numpy(), lobes.numpy(), (1.0, 1.0, 1.0)
    if case == 1:      # anisotropic spacing -> different border per axis, lung touching the volume faces
        ct, lobes = synthetic.make_volume(2, (24, 64, 40))
        lobes = lobes.numpy().copy()
        lobes[0, 10:30, 5:20] = 3
        lobes[:, 63, 7] = 5
        return ct.numpy(), lobes, (2.5, 0.7, 0.7)
    if case == 2:      # sparse specks: the dilation halo matters everywhere, odd sizes
        ct = g.integers(-1024, 400, size=(19, 23, 29)).astype(np.int16)
        lobes = (g.random((19, 23, 29)) < 0.003).astype(np.uint8) * 2
        lobes[9, 11, 14] = 1
        return ct, lobes, (1.0, 1.25, 5.0)
    ct = g.integers(-1024, 400, size=(8, 9, 10)).astype(np.int16)   # single voxel in a corner
    lobes = np.zeros((8, 9, 10), np.uint8)
    lobes[7, 0, 9] = 4
    return ct, lobes, (0.5, 0.5, 0.5)


@pytest.mark.parametrize("case", range(4))
def test_lung_presteps_match_oracle_bit_exact(cuda, lib, case):
    from dram_b200 import ops
    from dram_b200.dataset import grow_bbox

    ct, lobes, spacing = _lobes_case(case)
    ref = P.lung_crop_sample(ct, lobes, spacing=spacing, crop_border=5)
    lobe_d = torch.from_numpy(lobes).to(cuda)
    bbox = ops.mask_bbox(lobe_d).cpu().tolist()
    crop = grow_bbox(bbox, ct.shape, spacing, 5)
    assert np.array_equal(np.asarray(crop), ref["crop_slice"])
    image, lung, ess = ops.lung_crop(torch.from_numpy(ct).to(cuda), lobe_d, crop)
    assert np.array_equal(image.cpu().numpy(), ref["image"])
    assert np.array_equal(lung.cpu().numpy().astype(bool), ref["lung_mask"])
    assert np.array_equal(ess.cpu().numpy().astype(bool), ref["ess_mask"])


@pytest.mark.parametrize("shape,crop,density", [
    ((20, 40, 150), ((3, 17), (5, 33), (37, 129)), 0.02),    # crop inside the volume: the 2-voxel halo outside it counts
    ((12, 33, 70), ((0, 12), (0, 33), (0, 70)), 0.3),        # whole volume, zero border at every face, dense mask
    ((9, 24, 64), ((1, 9), (0, 24), (0, 64)), 0.05),         # crop width 64: 8-voxel vector stores, word-aligned rows
    ((7, 11, 40), ((2, 5), (3, 4), (33, 40)), 0.1),          # one row, 7 voxels wide at the right face
    ((6, 9, 131), ((0, 6), (2, 9), (1, 131)), 0.01),         # 130 wide: five bit words, ragged tail
])
def test_lung_crop_any_box_against_scipy(cuda, lib, shape, crop, density):
    """f1 for arbitrary crop boxes (dataset.py:66-80 only produces boxes around the lung): two dilations with the full
    3x3x3 structure computed by scipy on the whole volume, then cropped — the bit-mask kernels must agree exactly,
    including for lung voxels just outside the box and at the volume faces."""
    from scipy import ndimage

    from dram_b200 import ops

    g = np.random.default_rng(sum(shape))
    ct = g.integers(-1100, 300, size=shape).astype(np.int16)
    lobes = ((g.random(shape) < density) * g.integers(1, 6, size=shape)).astype(np.uint8)
    lung = lobes > 0
    dil = ndimage.binary_dilation(lung, structure=np.ones((3, 3, 3), bool), iterations=2)
    sl = tuple(slice(a, b) for a, b in crop)
    want_image = np.where(dil, ct, np.int16(-2048))[sl]
    want_lung = lung[sl]
    want_ess = (want_image < -910) & want_lung
    image, lung_c, ess = ops.lung_crop(torch.from_numpy(ct).to(cuda), torch.from_numpy(lobes).to(cuda), crop)
    assert np.array_equal(image.cpu().numpy(), want_image)
    assert np.array_equal(lung_c.cpu().numpy().astype(bool), want_lung)
    assert np.array_equal(ess.cpu().numpy().astype(bool), want_ess)


def test_mask_bbox_empty_and_full(cuda, lib):
    from dram_b200 import ops

    assert ops.mask_bbox(torch.zeros((5, 6, 7), dtype=torch.uint8, device=cuda)).cpu().tolist() == [5, 0, 6, 0, 7, 0]
    assert ops.mask_bbox(torch.ones((5, 6, 70), dtype=torch.uint8, device=cuda)).cpu().tolist() == [0, 5, 0, 6, 0, 70]


@pytest.mark.parametrize("dims,orig,crop", [
    ((16, 24, 20), (40, 50, 60), ((3, 35), (5, 45), (10, 58))),
    ((8, 8, 8), (9, 9, 9), ((0, 9), (0, 9), (0, 9))),
    ((12, 10, 14), (30, 31, 33), ((29, 30), (2, 3), (4, 33))),   # one-voxel extents: scale 0 branches
])
def test_heatmap_u8_matches_oracle(cuda, lib, dims, orig, crop):
    from dram_b200 import ops

    g = torch.Generator().manual_seed(9)
    dense = torch.rand((1,) + dims, generator=g)
    dense[0, :2] = 0.0
    dense[0, -1] = 1.0
    ref = P.postprocess_scan(dense, np.asarray(crop), orig)
    got = ops.heatmap_u8(dense[0].contiguous().to(cuda), crop, orig).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.uint8
    diff = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1
    assert (diff != 0).mean() <= 1e-3
    outside = np.ones(orig, bool)
    outside[tuple(slice(a, b) for a, b in crop)] = False
    assert not got[outside].any()


@pytest.mark.parametrize("dims,orig,crop", [
    ((20, 24, 160), (30, 40, 300), ((2, 29), (1, 38), (7, 291))),    # wide rows: several 128-voxel trips, ragged tail
    ((16, 16, 16), (20, 20, 150), ((0, 20), (0, 20), (0, 150))),     # whole volume, strong up-sampling along x
])
def test_heatmap_u8_sparse_map(cuda, lib, dims, orig, crop):
    """A dRAM is zero outside `ess`: the kernel skips the arithmetic of 32-voxel stretches whose eight sources are all
    zero.  A map of isolated blobs (most stretches empty, blobs cut by stretch borders) against the oracle."""
    from dram_b200 import ops

    g = torch.Generator().manual_seed(21)
    dense = torch.zeros((1,) + dims)
    for _ in range(12):
        z, y, x = (int(torch.randint(0, n - 2, (1,), generator=g)) for n in dims)
        dense[0, z:z + 3, y:y + 2, x:x + 5] = torch.rand((3, 2, 5), generator=g)[: dims[0] - z, : dims[1] - y, : dims[2] - x]
    ref = P.postprocess_scan(dense, np.asarray(crop), orig)
    got = ops.heatmap_u8(dense[0].contiguous().to(cuda), crop, orig).cpu().numpy()
    diff = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-3
    assert (got != 0).any() and (got == 0).mean() > 0.5
    outside = np.ones(orig, bool)
    outside[tuple(slice(a, b) for a, b in crop)] = False
    assert not got[outside].any()
