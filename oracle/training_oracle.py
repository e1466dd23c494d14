"""TEST INFRASTRUCTURE — CPU oracle of one training step of the seg-reg network.

Restates, in plain torch CPU fp32 ops with autograd, what `ScanRegLightningModule.shared_step(TRAIN)` does
(models.py:530-570) around the train-mode forward of med3d.py:369-388:
  * train-mode BatchNorm (batch statistics, biased variance for normalisation), shortcut type A without gradient
    (med3d.py:110 `out.data`);
  * `_generate_regression_labels` models.py:477-493, `_interval_regression_loss` models.py:495-506,
    `_segmentation_loss` models.py:508-518 with `BinaryDice(1e-7)` / `BinaryCrossEntropy` (metrics.py:4-47),
    total = loss_cle + loss_pse + 2 * mul_loss + seg_loss (models.py:565).
Pinned by tests/golden/train_step_med3ddram18.pt, produced by the unmodified reference network and the reference's
own loss code (oracle/make_golden.py gen_train_step).
"""
import contextlib

import torch
import torch.nn.functional as F

from . import med3d_oracle as M

BETA, GAMMA = 0.7338, 0.2578  # models.py:412-413


def batch_norm_train(sd, prefix, x):
    """nn.BatchNorm3d in training mode; running statistics are not part of the comparison."""
    return F.batch_norm(x, sd[prefix + ".running_mean"].clone(), sd[prefix + ".running_var"].clone(),
                        sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.1, M.BN_EPS)


@contextlib.contextmanager
def _detached_shortcut():
    old = M.SHORTCUT_DETACH
    M.SHORTCUT_DETACH = True
    try:
        yield
    finally:
        M.SHORTCUT_DETACH = old


def forward_train(sd, arch, x, lungs):
    """(dense_outs, reg_outs) of med3d.py:369-388 with train-mode BatchNorm."""
    B = x.shape[0]
    with _detached_shortcut():
        _, xup3 = M.features(sd, arch, x, bn=batch_norm_train)
    dense = [torch.sigmoid(F.conv3d(xup3, sd[f"fcs.{k}.weight"], sd[f"fcs.{k}.bias"])) for k in (0, 1)]
    m = F.interpolate(lungs, xup3.shape[-3:], mode="nearest")
    regs = [(d * m).view(B, -1).sum(dim=-1) / m.view(B, -1).sum(dim=-1) for d in dense]
    return dense, regs


def label_bands(labels, ratio_map, tightness=1.0):
    """models.py:477-493."""
    out = []
    for c in labels:
        lo, hi = ratio_map[int(c)]
        if lo < 1e-7:
            out.append((0.0, 0.0))
        else:
            mid, span = (lo + hi) / 2.0, (hi - lo) * tightness / 2.0
            out.append((mid - span, mid + span))
    return torch.tensor(out, dtype=torch.float32)


def interval_regression_loss(outs, bands, weights):
    """models.py:495-506."""
    d = torch.cat([outs.unsqueeze(1), bands], dim=1)
    d = BETA * d ** GAMMA
    K = (0.5 * (d[:, 2] - d[:, 1])) ** 2
    unhinged = (d[:, 0] - (d[:, 2] + d[:, 1]) / 2.0) ** 2 - K
    return (10.0 * F.leaky_relu(unhinged, negative_slope=0.0) * weights).sum()


def total_loss(dense, regs, lungs, ems, cle_labels, pse_labels, cle_bands, pse_bands, cle_w, pse_w):
    """models.py:547-565.  lungs / ems: float [B, 1, D, H, W]."""
    B = lungs.shape[0]
    loss_cle = interval_regression_loss(regs[0], cle_bands, cle_w)
    loss_pse = interval_regression_loss(regs[1], pse_bands, pse_w)
    binary = torch.logical_or(cle_labels > 0, pse_labels > 0).long()
    size = dense[0].shape[-3:]
    seg = F.interpolate(ems * binary.float().view(B, 1, 1, 1, 1), size, mode="nearest").detach()
    lung = F.interpolate(lungs, size=size, mode="nearest")
    a, b = dense[0] * lung, dense[1] * lung
    mul = (2.0 * (b.reshape(-1) * a.reshape(-1)).sum() + 1e-7) / (a.sum() + b.sum() + 1e-7)  # metrics.py:33-37
    p = torch.clamp(dense[0] + dense[1], min=0.0, max=1.0)
    t = seg.float()
    alpha = (1.0 - t.sum() / t.shape[0]).clamp(0.3, 0.7)  # metrics.py:18
    pt = p * t + (1.0 - p) * (1.0 - t)
    w = alpha * t + (1.0 - alpha) * (1.0 - t)
    logp = torch.log(pt.clamp(1e-6, 1.0 - 1e-6))
    nll = -1.0 * (0.85 * logp * w * lung + logp * w * (1.0 - lung))
    return loss_cle + loss_pse + 2.0 * mul + nll.sum() / w.sum()


def probe_loss(dense, regs):
    """A smooth, well-conditioned functional of the network outputs (fixed seeded linear probe of the dense maps plus
    the regression scores).  The reference's loss has a hinge and a 0.26-th power of scores near zero, which turns
    bf16-level forward noise into 10-40 % gradient noise (measured with a bf16-rounding copy of this oracle); the
    probe lets the gradient *plumbing* of every layer be compared tightly."""
    g = torch.Generator().manual_seed(77)
    total = 0.0
    for d in dense:
        r = torch.randn(tuple(d.shape), generator=g).to(d.device)
        total = total + (d * r).sum() / d[0].numel() ** 0.5
    return total + 3.0 * regs[0].sum() - 2.0 * regs[1].sum()


def train_step_grads(sd, arch, image, lungs, ems, cle_labels, pse_labels, cle_bands, pse_bands, cle_w, pse_w,
                     loss_fn=None):
    """image [B, D, H, W] fp32, lungs / ems [B, D, H, W] -> (loss, {parameter name: gradient}).
    `loss_fn(dense, regs)` replaces the reference's loss (see `probe_loss`)."""
    params = {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var"))}
    full = dict(sd)
    full.update(params)
    lungs5, ems5 = lungs.unsqueeze(1).float(), ems.unsqueeze(1).float()
    dense, regs = forward_train(full, arch, image.unsqueeze(1).float(), lungs5)
    if loss_fn is not None:
        loss = loss_fn(dense, regs)
    else:
        loss = total_loss(dense, regs, lungs5, ems5, cle_labels, pse_labels, cle_bands, pse_bands, cle_w, pse_w)
    names = list(params)
    grads = torch.autograd.grad(loss, [params[n] for n in names])
    return loss.detach(), dict(zip(names, grads)), [d.detach() for d in dense], [r.detach() for r in regs]


def train_case(batch=2, dims=(32, 32, 32), arch="med3ddram18", seed=2):
    """The seeded inputs of the golden training step (shared by make_golden, the CPU pin and the GPU test)."""
    from . import synthetic

    sd = synthetic.make_state_dict(arch, seed=seed, calib_dims=dims)
    xs, ls, es = zip(*[synthetic.make_network_input(40 + i, dims) for i in range(batch)])
    return {
        "arch": arch, "dims": dims, "batch": batch, "weight_seed": seed, "sd": sd,
        "image": torch.stack(xs), "lung_mask": torch.stack(ls).bool(), "em_mask": torch.stack(es).bool(),
        "cls_label": torch.tensor([3, 0][:batch]), "pse_label": torch.tensor([1, 2][:batch]),
        "cle_weights": torch.tensor([1.0, 2.0][:batch]), "pse_weights": torch.tensor([1.5, 0.5][:batch]),
    }


def grad_summary(grads, seed=123):
    """Per-parameter (L2 norm, projection on a seeded random direction): a compact fingerprint of a gradient set."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(grads):
        v = grads[name].detach().double().reshape(-1)
        r = torch.randn(v.numel(), generator=g, dtype=torch.float64)
        out[name] = (float(v.norm()), float((v * r).sum() / r.norm()))
    return out


def loss_closed_form(dense, lungs, ems, cle_labels, pse_labels, cle_bands, pse_bands, cle_w, pse_w):
    """The same loss as `total_loss` (with the lobe-masked means of med3d.py:387 computed inside) written the way the
    fused kernel K11 evaluates it: one pass of seven sums per scan, a scalar step, and the gradient of both maps in
    closed form — no autograd.  Returns (loss, [grad_cle, grad_pse], [reg_cle, reg_pse]).  test_oracle_cpu checks it
    against autograd of `total_loss`, which is pinned to the reference; the GPU test checks the kernel against
    `total_loss` directly, so this function is documentation of the derivation with a test attached.

    dense: 2 x [B,1,d,h,w] fp32 in (0,1); lungs / ems: float 0/1 [B,1,D,H,W]."""
    d0, d1 = dense[0].detach().double(), dense[1].detach().double()
    B, size = d0.shape[0], d0.shape[-3:]
    m = F.interpolate(lungs, size=size, mode="nearest").double()
    flag = torch.logical_or(cle_labels > 0, pse_labels > 0).double().view(B, 1, 1, 1, 1)
    t = F.interpolate(ems * flag.float(), size, mode="nearest").double()
    s32 = (dense[0].detach() + dense[1].detach())  # fp32 sum, as the network's outputs are added (models.py:513)
    both = s32.clamp(0.0, 1.0)
    pt32 = torch.where(t > 0, both, 1.0 - both)
    f = torch.where(m > 0, 0.85, 1.0).double()
    logp = torch.log(pt32.clamp(1e-6, 1.0 - 1e-6)).double() * f
    flat = lambda v: v.reshape(B, -1).sum(-1)  # noqa: E731
    S0, S1, M, I = flat(d0 * m), flat(d1 * m), flat(m), flat(d0 * m * d1 * m)
    T, A1, A0 = flat(t), flat(logp * (t > 0)), flat(logp * (t == 0))
    regs = [(S0 / M).float(), (S1 / M).float()]

    def interval(x, bands, w):  # models.py:495-506 and d/dx of it
        dx, dlo, dhi = (BETA * v.double() ** GAMMA for v in (x, bands[:, 0], bands[:, 1]))
        k, mid = 0.5 * (dhi - dlo), (dhi + dlo) / 2.0
        u = (dx - mid) ** 2 - k * k
        on = u > 0
        loss = (10.0 * u * w.double() * on).sum()
        grad = torch.where(on, 10.0 * w.double() * 2.0 * (dx - mid) * BETA * GAMMA * x.double() ** (GAMMA - 1.0), 0.0)
        return loss, grad

    loss_cle, g_cle = interval(regs[0], cle_bands, cle_w)
    loss_pse, g_pse = interval(regs[1], pse_bands, pse_w)
    num, den = 2.0 * I.sum() + 1e-7, S0.sum() + S1.sum() + 1e-7
    mul = num / den
    alpha = float((1.0 - T.sum() / B).clamp(0.3, 0.7))
    n_vox = float(B * size[0] * size[1] * size[2])
    wsum = alpha * T.sum() + (1.0 - alpha) * (n_vox - T.sum())
    seg = -(alpha * A1.sum() + (1.0 - alpha) * A0.sum()) / wsum
    loss = loss_cle + loss_pse + 2.0 * mul + seg
    c1, c2 = 4.0 / den, 2.0 * num / den ** 2
    # cross entropy through both clamps: ATen passes the gradient on the closed interval [min, max]
    open_ = (s32 >= 0) & (s32 <= 1) & (pt32 >= 1e-6) & (pt32 <= 1.0 - 1e-6)
    wv = torch.where(t > 0, alpha, 1.0 - alpha)
    dl = torch.where(open_, -(wv * f / wsum) / pt32.double() * torch.where(t > 0, 1.0, -1.0), 0.0)
    shape = (B, 1, 1, 1, 1)
    grads = [m * ((g_cle / M).view(shape) + c1 * d1 * m - c2) + dl,
             m * ((g_pse / M).view(shape) + c1 * d0 * m - c2) + dl]
    return loss.float(), [g.float() for g in grads], regs
