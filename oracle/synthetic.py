"""Seeded synthetic CT volumes, lobe masks and checkpoints shared by the oracle and the B200 path.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Nothing here comes from the reference except the shapes/semantics it documents:
  * volumes follow SURVEY.md §8d (two ellipsoid lungs cut into 5 lobes, HU statistics chosen so
    that the LAA-910 `ess` mask of dataset.py:79 is non-trivial);
  * checkpoints are "calibrated random" (SURVEY.md §7.2 H6): the reference's own init
    (med3d.py:334-339: kaiming-normal fan_out convs, BN gamma=1 beta=0) makes eval-mode
    activations explode and the sigmoid heads saturate, so parity would be vacuous.  Here the BN
    affine parameters are drawn at random, the BN running statistics are set from one fp64
    calibration pass (so every conv output is O(1)), and the heads are rescaled to O(1) logits.
    Calibration runs in fp64 and is rounded to fp32 once, which makes the result reproducible
    across CPUs to ~1e-15 before rounding.
The real paper.ckpt / best.ckpt are Git-LFS pointers in the reference checkout and cannot be loaded.
"""
import math

import torch

from . import med3d_oracle as M

BASE_SEED = 20261018


def make_volume(index, dims, dtype=torch.int16):
    """(ct int16 [D,H,W] in HU, lobes uint8 [D,H,W] with labels 0..5) for volume `index`."""
    D, H, W = dims
    g = torch.Generator().manual_seed(BASE_SEED + index)
    zz = (torch.arange(D, dtype=torch.float32) + 0.5) / D
    yy = (torch.arange(H, dtype=torch.float32) + 0.5) / H
    xx = (torch.arange(W, dtype=torch.float32) + 0.5) / W
    z, y, x = zz[:, None, None], yy[None, :, None], xx[None, None, :]
    a, b, c = 0.42, 0.36, 0.19
    right = ((z - 0.5) / a) ** 2 + ((y - 0.5) / b) ** 2 + ((x - 0.30) / c) ** 2 <= 1.0
    left = ((z - 0.5) / a) ** 2 + ((y - 0.5) / b) ** 2 + ((x - 0.70) / c) ** 2 <= 1.0
    lobes = torch.zeros((D, H, W), dtype=torch.uint8)
    zfull = z.expand(D, H, W)
    lobes[right & (zfull < 0.40)] = 1
    lobes[right & (zfull >= 0.40) & (zfull < 0.62)] = 2
    lobes[right & (zfull >= 0.62)] = 3
    lobes[left & (zfull < 0.50)] = 4
    lobes[left & (zfull >= 0.50)] = 5
    lung = lobes > 0
    ct = torch.randn((D, H, W), generator=g) * 150.0 + 20.0
    inside = torch.randn((D, H, W), generator=g) * 120.0 - 850.0
    ct = torch.where(lung, inside, ct)
    # emphysema-like blobs: smooth low-resolution noise thresholded at the 92nd lung percentile
    coarse = torch.rand((1, 1, max(D // 8, 2), max(H // 8, 2), max(W // 8, 2)), generator=g)
    field = torch.nn.functional.interpolate(coarse, size=(D, H, W), mode="trilinear", align_corners=True)[0, 0]
    vals = field[lung]
    if vals.numel() > 0:
        k = max(1, int(math.ceil(0.92 * vals.numel())))
        thr = torch.kthvalue(vals, k).values
        blob = lung & (field > thr)
        low = torch.randn((D, H, W), generator=g) * 25.0 - 960.0
        ct = torch.where(blob, low, ct)
    ct = ct.round().clamp(-1024, 1500).to(dtype)
    return ct, lobes


def make_network_input(index, dims):
    """(image fp32 [D,H,W], lung uint8, ess uint8) already at network size: the tensors that
    reach predict_step after the transforms (window -> standardise; no resize needed)."""
    ct, lobes = make_volume(index, dims)
    lung = (lobes > 0)
    ess = (ct < -910) & lung  # dataset.py:79 (without the -2048 fill outside the dilated lung)
    v = torch.clamp(ct.float(), -1150.0, -300.0)
    v = (v - (-1150.0)) / (-300.0 - (-1150.0))
    v = (v - v.mean()) / v.std()
    return v, lung.to(torch.uint8), ess.to(torch.uint8)


def _calibrating_bn(sd64, stats):
    def bn(sd, prefix, x):
        mean = x.mean(dim=(0, 2, 3, 4))
        var = x.var(dim=(0, 2, 3, 4), unbiased=False)
        sd[prefix + ".running_mean"] = mean
        sd[prefix + ".running_var"] = var.clamp_min(1e-6)
        stats[prefix] = (mean, var)
        return M.batch_norm_eval(sd, prefix, x)

    return bn


def make_state_dict(arch, seed=0, calib_dims=(32, 32, 32), prefix="", logit_std=1.5, reg_bias=-2.0, device=None):
    """Deterministic calibrated-random `state_dict` (fp32) for conf/<arch>.yaml.

    Keys/shapes/order equal the reference module's; `prefix='model.'` gives
    the Lightning-checkpoint naming that load_state_dict_greedy expects (utils.py:226-249).
    `device`: where the fp64 calibration pass runs (default CPU; a CUDA device makes a large calibration volume
    affordable — the full-size C4 test calibrates on a 1/3-scale volume of the same aspect).  The random draws always
    come from the CPU generator, the result is returned on the CPU.
    """
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    sd = {}
    for key, shape, kind in M.state_layout(arch):
        if kind == "conv_w":
            fan_out = shape[0] * shape[2] * shape[3] * shape[4]
            sd[key] = torch.randn(shape, generator=g, dtype=torch.float64) * math.sqrt(2.0 / fan_out)
        elif kind == "conv_b":
            sd[key] = torch.randn(shape, generator=g, dtype=torch.float64) * 0.05
        elif kind == "bn_w":
            sd[key] = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        elif kind == "bn_b":
            sd[key] = torch.randn(shape, generator=g, dtype=torch.float64) * 0.2
        elif kind == "bn_mean":
            sd[key] = torch.zeros(shape, dtype=torch.float64)
        elif kind == "bn_var":
            sd[key] = torch.ones(shape, dtype=torch.float64)
        else:
            sd[key] = torch.zeros(shape, dtype=torch.int64)
    # Residual branches get a small gain (gamma of the last BN of each block ~ U(0.15, 0.45)), in the
    # spirit of zero-init-residual training: with O(1) gains a random, untrained ResNet is chaotic (a
    # 1e-3 perturbation of an early activation grows to O(1) at the heads), which trained networks are not.
    kind, layers, head = M.ARCHS[arch]
    last_bn = "bn3" if kind == "bottleneck" else "bn2"
    for li, nb in enumerate(layers, start=1):
        for bi in range(nb):
            sd[f"layer{li}.{bi}.{last_bn}.weight"] *= 0.3
    # calibration pass (fp64): sets every BN's running statistics from the activations it sees
    img, _, _ = make_network_input(9000 + seed, calib_dims)
    x = img.double()[None, None]
    if device is not None:
        sd = {k: v.to(device) for k, v in sd.items()}
        x = x.to(device)
    stats = {}
    _, xup3 = M.features(sd, arch, x, bn=_calibrating_bn(sd, stats))
    # heads: O(1) logits on the calibration volume
    for k in (0, 1):
        w = sd[f"fcs.{k}.weight"]
        logits = torch.nn.functional.conv3d(xup3, w)
        s = logits.std(dim=(0, 2, 3, 4)).clamp_min(1e-12)
        sd[f"fcs.{k}.weight"] = w * (logit_std / s).view(-1, 1, 1, 1, 1)
        mean = (logits.mean(dim=(0, 2, 3, 4)) * (logit_std / s))
        if head == "reg":
            sd[f"fcs.{k}.bias"] = reg_bias - mean
        else:
            sd[f"fcs.{k}.bias"] = torch.randn(w.shape[0], generator=g, dtype=torch.float64).to(mean.device) * 0.5 - mean
    out = {}
    for key, v in sd.items():
        v = v.cpu()
        out[prefix + key] = v.to(torch.float32) if v.dtype == torch.float64 else v
    return out


def state_dict_checksum(sd):
    """Order-independent fp64 fingerprint of a state_dict (detects RNG / calibration drift)."""
    tot = 0.0
    for key in sorted(sd):
        v = sd[key].double()
        tot += float(v.sum()) + 0.5 * float((v * v).sum())
    return tot
