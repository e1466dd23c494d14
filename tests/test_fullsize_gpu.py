"""The BASELINE.json configurations at their full sizes.

C3 (med3ddram / ResNet-34, one 256^3 volume) is small enough for the CPU oracle (a few seconds on the GPU
box's host cores), so it is compared voxel by voxel.  C2 (med3ddram18, 256^3, batch 4) and C4 (med3ddram50,
400x512x512) are checked through size-independent properties: exact zeros outside the `ess` mask, lesion
percentages that equal the sums of the returned maps (models.py:440-441), identical results for identical
volumes at different batch positions, run-to-run determinism, and sigmoid range.
Tolerances (north star): dRAM voxels <= 2e-2 max-abs, percentages <= 1e-2 relative, mask support bit-exact.
"""
import os
import sys
from argparse import Namespace

import pytest
import torch

from oracle import pipeline_oracle as P
from oracle import synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _module(arch, sd, cuda):
    from dram_b200.models import ScanRegLightningModule

    module = ScanRegLightningModule(Namespace(model_arch=arch))
    module.model.load_state_dict(sd)
    return module.to(cuda).eval()


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench

    return bench


def test_c3_resnet34_256cube_matches_oracle(cuda, lib):
    arch, dims = "med3ddram", (256, 256, 256)
    sd = synthetic.make_state_dict(arch, seed=0, calib_dims=(64, 64, 64))
    module = _module(arch, sd, cuda)
    x, lung, ess = synthetic.make_network_input(3, dims)
    batch = {"image": x[None], "lung_mask": lung[None].bool(), "ess_mask": ess[None].bool()}
    torch.set_num_threads(os.cpu_count() or 1)
    ref = P.predict_step(sd, arch, batch)
    got = module.predict_step({k: v.to(cuda) for k, v in batch.items()}, 0)
    for k in ("cle_dense_outs", "pse_dense_outs"):
        g = got[k].cpu()
        assert g.shape == ref[k].shape == (1, 1) + dims
        assert (g - ref[k]).abs().max().item() <= 2e-2, k
        assert torch.equal(g == 0, ref[k] == 0), f"{k}: ess-mask support differs"
    for k in ("cle_precentages", "pse_precentages"):
        rel = ((got[k].cpu() - ref[k]).abs() / ref[k].abs().clamp_min(1e-6)).max().item()
        assert rel <= 1e-2, (k, rel)


def _check_properties(module, hu, lungs, ess, cuda):
    out = module.predict_step_from_hu(hu, lungs, ess)
    torch.cuda.synchronize()
    lung_total = lungs.sum(dtype=torch.float64)
    for k, pk in (("cle_dense_outs", "cle_precentages"), ("pse_dense_outs", "pse_precentages")):
        m = out[k]
        assert m.shape == (hu.shape[0], 1) + tuple(hu.shape[1:]) and m.dtype == torch.float32
        assert torch.isfinite(m).all()
        assert float(m.min()) >= 0.0 and float(m.max()) <= 1.0            # sigmoid maps times a 0/1 mask
        assert not bool((m[:, 0][ess == 0] != 0).any())                    # exact zeros outside ess
        assert bool((m[:, 0][ess != 0] > 0).all())                         # sigmoid is strictly positive inside
        pct = m.sum(dim=(1, 2, 3, 4), dtype=torch.float64) / lung_total    # models.py:440-441 (batch-wide lungs)
        assert torch.allclose(out[pk].double(), pct, rtol=1e-4), (out[pk], pct)
    return out


def test_c2_resnet18_256cube_batch4_properties(cuda, lib):
    bench = _bench()
    arch, dims = "med3ddram18", (256, 256, 256)
    module = bench.build_module(cuda, arch)
    hu2, lungs2, ess2 = bench.make_volumes(2, dims, cuda, seed=11)
    # volumes 0/2 and 1/3 are identical: every per-volume result must be too, bit for bit
    hu, lungs, ess = (torch.cat([t, t]) for t in (hu2, lungs2, ess2))
    out = _check_properties(module, hu, lungs, ess, cuda)
    for k in ("cle_dense_outs", "pse_dense_outs", "cle_precentages", "pse_precentages"):
        assert torch.equal(out[k][0], out[k][2]) and torch.equal(out[k][1], out[k][3]), k
    first = {k: out[k].clone() for k in ("cle_dense_outs", "cle_precentages")}
    again = module.predict_step_from_hu(hu, lungs, ess)
    for k, v in first.items():
        assert torch.equal(again[k], v), f"{k}: not deterministic"
    # a batch of 2 reports twice the percentage of the same volumes in a batch of 4 (quirk Q1) and the same maps
    half = module.predict_step_from_hu(hu2, lungs2, ess2)
    assert torch.equal(half["cle_dense_outs"], out["cle_dense_outs"][:2])
    assert torch.allclose(half["cle_precentages"], 2 * out["cle_precentages"][:2], rtol=1e-5)


def test_c4_resnet50_400x512x512_properties(cuda, lib):
    bench = _bench()
    arch, dims = "med3ddram50", (400, 512, 512)
    module = bench.build_module(cuda, arch)
    hu, lungs, ess = bench.make_volumes(1, dims, cuda, seed=12)
    out = _check_properties(module, hu, lungs, ess, cuda)
    again = module.predict_step_from_hu(hu, lungs, ess)
    assert torch.equal(again["pse_dense_outs"], out["pse_dense_outs"])
    del out, again
    module.model._engines.clear()
    torch.cuda.empty_cache()
