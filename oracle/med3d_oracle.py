"""CPU oracle of the Med3D seg-reg / seg-cls forward pass.  TEST INFRASTRUCTURE, not product code.

A functional restatement, in plain torch CPU ops on a `state_dict`, of
/root/reference/med3d.py (ResNetSegReg.forward 369-388, ResNetSegCls.forward 270-285 and the
blocks they call).  The arithmetic itself lives in PyTorch's ATen (the reference pins
torch==1.12.0, install_files/requirements.in:11; this image has 2.11) — conv3d, eval batch_norm,
max_pool3d, trilinear/nearest interpolate — so the restatement calls the same ATen ops in fp32
(or fp64 for weight calibration) and differs from the reference only in structure: no nn.Module,
no in-place ops, state read straight from the dict.

Pinned by tests/test_oracle_cpu.py against golden vectors produced by the unmodified reference
(oracle/make_golden.py -> tests/golden/), and against the live reference when /root/reference exists.
"""
import torch
import torch.nn.functional as F

# conf/<arch>.yaml -> (block kind, blocks per layer, head kind)      [conf/*.yaml, med3d.py:391-425]
ARCHS = {
    "med3d": ("basic", (3, 4, 6, 3), "cls"),
    "med3d18": ("basic", (2, 2, 2, 2), "cls"),
    "med3d50": ("bottleneck", (3, 4, 6, 3), "cls"),
    "med3ddram": ("basic", (3, 4, 6, 3), "reg"),
    "med3ddram18": ("basic", (2, 2, 2, 2), "reg"),
    "med3ddram50": ("bottleneck", (3, 4, 6, 3), "reg"),
}
# (planes, stride, dilation) of layer1..4                             [med3d.py:306-312]
LAYER_CFG = ((64, 1, 1), (128, 2, 1), (256, 1, 2), (512, 1, 4))
N_CLASSES = (6, 3)  # conf/med3d*.yaml n_classes
BN_EPS = 1e-5


def expansion(kind):
    return 4 if kind == "bottleneck" else 1  # med3d.py:116, 148


def batch_norm_eval(sd, prefix, x):
    """nn.BatchNorm3d in eval mode (running statistics), e.g. med3d.py:121, 303."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.0, BN_EPS)


# med3d.py:110 builds the shortcut from `out.data`: in training no gradient flows through shortcut type A.
# The training oracle switches this on; the forward value is the same either way.
SHORTCUT_DETACH = False


def shortcut_a(x, planes, stride):
    """med3d.py:103-112: avg_pool3d(kernel 1, stride) is a strided subsample; pad channels with zeros."""
    out = (x.detach() if SHORTCUT_DETACH else x)[:, :, ::stride, ::stride, ::stride]
    pad = planes - out.shape[1]
    if pad > 0:
        out = torch.cat([out, out.new_zeros((out.shape[0], pad) + tuple(out.shape[2:]))], dim=1)
    return out


def basic_block(sd, p, x, planes, stride, dilation, downsample, bn):
    """med3d.py:129-144."""
    out = F.conv3d(x, sd[p + ".conv1.weight"], None, stride, dilation, dilation)
    out = torch.relu(bn(sd, p + ".bn1", out))
    out = F.conv3d(out, sd[p + ".conv2.weight"], None, 1, dilation, dilation)
    out = bn(sd, p + ".bn2", out)
    res = shortcut_a(x, planes, stride) if downsample else x
    return torch.relu(out + res)


def bottleneck_block(sd, p, x, planes, stride, dilation, downsample, bn):
    """med3d.py:164-184."""
    out = torch.relu(bn(sd, p + ".bn1", F.conv3d(x, sd[p + ".conv1.weight"])))
    out = F.conv3d(out, sd[p + ".conv2.weight"], None, stride, dilation, dilation)
    out = torch.relu(bn(sd, p + ".bn2", out))
    out = bn(sd, p + ".bn3", F.conv3d(out, sd[p + ".conv3.weight"]))
    res = shortcut_a(x, planes * 4, stride) if downsample else x
    return torch.relu(out + res)


def centre_crop_like(skip, ref_shape):
    """crop_concat_5d's crop of the skip tensor (med3d.py:44-46): start = ceil((b-a)/2) per axis."""
    sl = [slice(None), slice(None)]
    for a, b in zip(ref_shape[2:], skip.shape[2:]):
        if a > b:
            raise ValueError("up-sampled tensor larger than the skip tensor (med3d.py:43)")
        start = -((a - b) // 2)  # ceil((b-a)/2)
        sl.append(slice(start, start + a))
    return skip[tuple(sl)]


def up_block(sd, p, x, skip, bn):
    """UpsampleConvBlock5d.forward, med3d.py:85-89 (+ conv blocks 74-80)."""
    up = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    y = torch.cat([up, centre_crop_like(skip, up.shape)], dim=1)
    for i in (0, 1):
        q = f"{p}.conv_blocks.{i}"
        y = F.conv3d(y, sd[q + ".0.weight"], sd[q + ".0.bias"], 1, 1)
        y = torch.relu(bn(sd, q + ".1", y))
    return y


def features(sd, arch, x, bn=batch_norm_eval, taps=None):
    """Everything up to `xup3` (med3d.py:371-381 / 272-282).  `taps`, if a dict, receives named
    intermediate activations for kernel bring-up."""
    kind, layers, _ = ARCHS[arch]
    block = bottleneck_block if kind == "bottleneck" else basic_block
    e = expansion(kind)
    x = torch.relu(bn(sd, "bn1", F.conv3d(x, sd["conv1.weight"], None, 2, 3)))
    xp = F.max_pool3d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["conv1"], taps["maxpool"] = x, xp
    inplanes = 64
    feats = []
    y = xp
    for li, ((planes, stride, dilation), nblocks) in enumerate(zip(LAYER_CFG, layers), start=1):
        for bi in range(nblocks):
            first = bi == 0
            ds = first and (stride != 1 or inplanes != planes * e)  # med3d.py:343
            y = block(sd, f"layer{li}.{bi}", y, planes, stride if first else 1, dilation, ds, bn)
            if first:
                inplanes = planes * e
        feats.append(y)
        if taps is not None:
            taps[f"layer{li}"] = y
    x1, x4 = feats[0], feats[3]
    xup1 = up_block(sd, "us1", x4, x1, bn)
    xup2 = up_block(sd, "us2", xup1, x, bn)
    xup3 = F.conv3d(xup2, sd["us3.0.weight"], sd["us3.0.bias"], 1, 1)
    xup3 = torch.relu(bn(sd, "us3.1", xup3))
    if taps is not None:
        taps["us1"], taps["us2"], taps["us3"] = xup1, xup2, xup3
    return x, xup3


def forward(sd, arch, x, lungs=None, taps=None):
    """(dense_outs, reg_outs | cls_outs) exactly as the reference module returns them.

    reg: med3d.py:382-388 — sigmoid heads, lungs resampled with legacy `nearest`, masked mean
         (lungs=None -> ones_like(stem output), i.e. the plain mean, quirk Q3).
    cls: med3d.py:283-285 — raw heads, global average pool; `lungs` ignored.
    """
    head = ARCHS[arch][2]
    B = x.shape[0]
    stem, xup3 = features(sd, arch, x, taps=taps)
    if head == "cls":
        dense = [F.conv3d(xup3, sd[f"fcs.{k}.weight"], sd[f"fcs.{k}.bias"]) for k in (0, 1)]
        return dense, [F.adaptive_avg_pool3d(d, 1).view(B, -1) for d in dense]
    dense = [torch.sigmoid(F.conv3d(xup3, sd[f"fcs.{k}.weight"], sd[f"fcs.{k}.bias"])) for k in (0, 1)]
    if lungs is None:
        m = torch.ones_like(stem)
    else:
        m = F.interpolate(lungs, xup3.shape[-3:], mode="nearest")
    regs = [(d * m).view(B, -1).sum(dim=-1) / m.view(B, -1).sum(dim=-1) for d in dense]
    return dense, regs


# ------------------------------------------------------------------------------------------
# state_dict layout — names, shapes and order of the reference modules
# ------------------------------------------------------------------------------------------
def state_layout(arch):
    """[(key, shape, kind)] in the reference's registration order; kind in
    {'conv_w', 'conv_b', 'bn_w', 'bn_b', 'bn_mean', 'bn_var', 'bn_count'}."""
    kind, layers, head = ARCHS[arch]
    e = expansion(kind)
    out = []

    def bn(p, c):
        out.extend([(p + ".weight", (c,), "bn_w"), (p + ".bias", (c,), "bn_b"),
                    (p + ".running_mean", (c,), "bn_mean"), (p + ".running_var", (c,), "bn_var"),
                    (p + ".num_batches_tracked", (), "bn_count")])

    out.append(("conv1.weight", (64, 1, 7, 7, 7), "conv_w"))
    bn("bn1", 64)
    inplanes = 64
    for li, ((planes, _, _), nblocks) in enumerate(zip(LAYER_CFG, layers), start=1):
        for bi in range(nblocks):
            p = f"layer{li}.{bi}"
            if kind == "basic":
                out.append((p + ".conv1.weight", (planes, inplanes, 3, 3, 3), "conv_w"))
                bn(p + ".bn1", planes)
                out.append((p + ".conv2.weight", (planes, planes, 3, 3, 3), "conv_w"))
                bn(p + ".bn2", planes)
            else:
                out.append((p + ".conv1.weight", (planes, inplanes, 1, 1, 1), "conv_w"))
                bn(p + ".bn1", planes)
                out.append((p + ".conv2.weight", (planes, planes, 3, 3, 3), "conv_w"))
                bn(p + ".bn2", planes)
                out.append((p + ".conv3.weight", (planes * 4, planes, 1, 1, 1), "conv_w"))
                bn(p + ".bn3", planes * 4)
            inplanes = planes * e
    for name, cin in (("us1", (512 + 64) * e), ("us2", 128)):
        for i, c in enumerate((cin, 64)):
            q = f"{name}.conv_blocks.{i}"
            out.append((q + ".0.weight", (64, c, 3, 3, 3), "conv_w"))
            out.append((q + ".0.bias", (64,), "conv_b"))
            bn(q + ".1", 64)
    out.append(("us3.0.weight", (32, 64, 3, 3, 3), "conv_w"))
    out.append(("us3.0.bias", (32,), "conv_b"))
    bn("us3.1", 32)
    for k, c in enumerate(N_CLASSES if head == "cls" else (1, 1)):
        out.append((f"fcs.{k}.weight", (c, 32, 1, 1, 1), "conv_w"))
        out.append((f"fcs.{k}.bias", (c,), "conv_b"))
    return out


def conv_flops(arch, dims, batch=1):
    """Algorithmic conv FLOPs (2*M*N*K over every nn.Conv3d call) for an input of `dims`."""
    def co(n, k, s, d, p):
        return (n + 2 * p - d * (k - 1) - 1) // s + 1

    kind, layers, head = ARCHS[arch]
    e = expansion(kind)
    total = 0
    d1 = tuple(co(n, 7, 2, 1, 3) for n in dims)
    vox = lambda t: t[0] * t[1] * t[2]  # noqa: E731
    total += 2 * vox(d1) * 64 * 343
    cur = tuple(co(n, 3, 2, 1, 1) for n in d1)
    d2 = cur
    inplanes = 64
    for (planes, stride, _), nblocks in zip(LAYER_CFG, layers):
        for bi in range(nblocks):
            s = stride if bi == 0 else 1
            nxt = tuple(co(n, 3, s, 1, 1) for n in cur)
            if kind == "basic":
                total += 2 * vox(nxt) * planes * inplanes * 27 + 2 * vox(nxt) * planes * planes * 27
            else:
                total += 2 * vox(cur) * planes * inplanes
                total += 2 * vox(nxt) * planes * planes * 27
                total += 2 * vox(nxt) * planes * 4 * planes
            inplanes = planes * e
            cur = nxt
    total += 2 * vox(d2) * 64 * ((512 + 64) * e) * 27 + 2 * vox(d2) * 64 * 64 * 27
    total += 2 * vox(d1) * 64 * 128 * 27 + 2 * vox(d1) * 64 * 64 * 27
    total += 2 * vox(d1) * 32 * 64 * 27
    total += 2 * vox(d1) * 32 * (sum(N_CLASSES) if head == "cls" else 2)
    return total * batch
