"""Training step of the Med3D seg-reg network on the backward kernels (BASELINE config 5).

What runs where, stated plainly:

* every convolution — forward, data gradient and weight gradient, >99.9 % of the FLOPs of a training step — runs on
  the hand-written sm_100a kernels (`ops.Conv3dPlan`, `ops.stem_conv7`, `backward.Conv3dDgradPlan`,
  `backward.Conv3dWgradPlan`) through `ConvFn` / `StemFn`, autograd Functions over NDHWC bf16 activations;
* train-mode BatchNorm with its ReLU and residual add (forward: statistics + running-stat update + apply; backward:
  reduction + apply, optional SyncBatchNorm exchange between the phases) runs on the K10 kernels (`bn_train.cu`,
  `backward.BatchNormTrainFn`);
* max-pool and x2 trilinear up-sampling run on K3 / K4 forward and their gather-style adjoints K3T / K4T
  (`train_glue.cu`, `backward.MaxPool3dFn` / `Upsample2xFn`);
* the two 1x1x1 sigmoid heads run on K5T (`backward.HeadsSigmoidFn`), conv bias gradients on K10's channel sums;
* the per-step re-packing of the fp32 master weights into the kernels' 16-bit operands is one kernel per operand
  (`backward.pack_weight_into`);
* lobe-masked pooling + the three losses, forward and gradient, are K11 (`backward.TrainLossFn`, train_loss.cu); Adam is
  K12, one launch over the flat parameter / gradient buffers (`backward.FlatAdam`); the convolutions' weight gradients
  are written by K9 straight into the flat gradient buffer (no AccumulateGrad pass);
* what is left on ATen (cuDNN disabled) is shortcut-A slicing/padding, the stem weight-gradient re-layout and dtype
  casts of per-channel vectors.  `TrainStep(loss="aten", optimizer="torch")` keeps the earlier ATen loss and
  torch.optim.Adam for A/B checks.  `bench.py --mode train` measures the step as an additional line and names the
  remainder in `config.glue`.

The module takes the drop-in network (`med3d.resnet{18,34,50}segreg()`, reference `state_dict` keys) and reproduces
what the reference does in `ScanRegLightningModule.shared_step(TRAIN)` (models.py:530-570): forward in train mode
(med3d.py:369-388), loss = interval regression x2 + 2 x dice + BCE (models.py:495-518, metrics.py:4-47), backward.
Data-parallel training (train.py:100 `strategy=ddp`) averages the gradients over the process group with one
all-reduce per bucket (`backward.GradBuckets`; NCCL over NVLink on the GPUs).
Quirk kept: shortcut type A carries no gradient (`out.data`, med3d.py:110).
"""
import torch
import torch.nn.functional as F

from . import ops
from .backward import (BatchNormTrainFn, Conv3dDgradPlan, Conv3dWgradPlan, FlatAdam, GradBuckets, HeadsSigmoidFn, MaxPool3dFn,
                       PeerExchange, TrainLossFn, Upsample2xFn, channel_sums, pack_weight_into)
from .engine import LAYER_CFG

ACT = torch.bfloat16  # activations and their gradients
# Test hook: a list here receives (layer name, [sources], dy, weight, [dx per source], dw) clones from every
# ConvFn.backward so that each layer's dgrad / wgrad can be checked in isolation on the tensors it really saw.
BACKWARD_TAP = None
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


def _ncdhw(x):
    """NDHWC tensor -> NCDHW-logical view (channels_last_3d strides) for ATen ops; no copy."""
    return x.permute(0, 4, 1, 2, 3)


def _ndhwc(v):
    """NCDHW-logical tensor -> contiguous NDHWC (no copy when it already is channels-last)."""
    return v.permute(0, 2, 3, 4, 1).contiguous()


class _PlanCache:
    """Plans bake device pointers into TMA tensor maps; tensors produced by ATen move between steps only while the
    caching allocator warms up, so a layer keeps the plan of the last pointer set it saw."""

    def __init__(self):
        self.key, self.plan = None, None

    def get(self, key, make):
        if self.key != key:
            self.plan = None  # release the old buffers first
            self.plan, self.key = make(), key
        return self.plan


class ConvLayer:
    """Per-convolution state: geometry + cached forward / dgrad / wgrad plans (index = source of a concatenation)."""

    def __init__(self, name, kernel, stride, dilation):
        self.name, self.k, self.s, self.dl = name, kernel, stride, dilation
        self.fwd, self.dgrad, self.wgrad = _PlanCache(), [_PlanCache(), _PlanCache()], [_PlanCache(), _PlanCache()]
        self.zero_bias = None
        self.flops = 0  # algorithmic forward FLOPs (2*M*N*K) of the last call
        self.dw = None  # fp32 [Cout, Cin_total, kd, kh, kw], shared by the wgrad plans of both sources
        # (flat-buffer view of weight.grad, callback): when set, K9 writes the weight gradient there directly and the
        # callback replaces autograd's post-accumulate hook (TrainStep)
        self.grad_sink = None


class ConvFn(torch.autograd.Function):
    """y = conv3d(cat([x1, x2]), weight) (+ bias) on NDHWC bf16; x2 optional (decoder concat, med3d.py:87)."""

    @staticmethod
    def forward(ctx, layer, weight, bias, x1, x2):
        srcs = [x1] if x2 is None else [x1, x2]
        cout = weight.shape[0]
        if layer.zero_bias is None or layer.zero_bias.device != x1.device:
            layer.zero_bias = torch.zeros(cout, dtype=torch.float32, device=x1.device)
        taps = weight.shape[2] * weight.shape[3] * weight.shape[4]

        def make():
            w_buf = torch.empty((cout, taps * weight.shape[1]), dtype=ACT, device=x1.device)
            b_buf = torch.zeros_like(layer.zero_bias)
            plan = ops.Conv3dPlan(x1, w_buf, b_buf, x2=x2, kernel=layer.k, stride=layer.s, dilation=layer.dl, relu=False)
            return plan, w_buf, b_buf

        plan, w_buf, b_buf = layer.fwd.get(tuple(t.data_ptr() for t in srcs) + tuple(x1.shape), make)
        pack_weight_into(weight, w_buf)
        if bias is not None:
            b_buf.copy_(bias.detach())
        layer.flops = plan.flops
        out = plan.run().detach()  # a fresh alias: the plan-owned buffer itself never carries autograd history
        ctx.layer, ctx.has_bias = layer, bias is not None
        ctx.save_for_backward(weight, *srcs)
        return out

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        weight, *srcs = ctx.saved_tensors
        dy = dy.contiguous()
        cin_total = weight.shape[1]
        if layer.grad_sink is not None:
            dw = layer.grad_sink[0]
        else:
            if layer.dw is None or layer.dw.device != dy.device:
                layer.dw = torch.zeros(tuple(weight.shape), dtype=torch.float32, device=dy.device)
            dw = layer.dw
        grads, off = [None, None], 0
        for i, x in enumerate(srcs):
            c = x.shape[4]
            wp = layer.wgrad[i].get((x.data_ptr(), dy.data_ptr(), dw.data_ptr()) + tuple(x.shape), lambda: Conv3dWgradPlan(
                x, dy, dw=dw, kernel=layer.k, stride=layer.s, dilation=layer.dl, cin_total=cin_total, cin_offset=off))
            wp.run()
            if ctx.needs_input_grad[3 + i]:
                dp = layer.dgrad[i].get((dy.data_ptr(),) + tuple(x.shape), lambda: Conv3dDgradPlan(
                    dy, weight, tuple(x.shape[1:4]), kernel=layer.k, stride=layer.s, dilation=layer.dl,
                    cin_range=(off, off + c)))
                pack_weight_into(weight, dp.packed, transpose=True, cin_range=(off, off + c))
                grads[i] = dp.run().detach()
            off += c
        db = channel_sums(dy) if ctx.has_bias else None
        if BACKWARD_TAP is not None:
            BACKWARD_TAP.append((layer.name, [x.clone() for x in srcs], dy.clone(), weight.detach().clone(),
                                 [None if g is None else g.clone() for g in grads[:len(srcs)]], dw.clone()))
        if layer.grad_sink is not None:  # weight.grad is complete in the flat buffer: tell the bucket, skip AccumulateGrad
            layer.grad_sink[1]()
            return None, None, db, grads[0], grads[1]
        # AccumulateGrad adds layer.dw into weight.grad right away (stream order), so the shared buffer can be reused
        return None, dw, db, grads[0], grads[1]


class StemFn(torch.autograd.Function):
    """conv1 (7^3, stride 2, 1 -> 64, med3d.py:296-304) on the fused stem kernel; its weight gradient goes through the
    (kh, kw)-unfolded image (K2a) and the streaming wgrad kernel as a 7x1x1 convolution over 64 pseudo-channels."""

    @staticmethod
    def forward(ctx, layer, weight, image):
        if layer.zero_bias is None or layer.zero_bias.device != image.device:
            layer.zero_bias = torch.zeros(64, dtype=torch.float32, device=image.device)
        packed = ops.pack_stem_weight_fused(weight, dtype=ACT)
        out = ops.stem_conv7(image, packed, layer.zero_bias, relu=False)
        layer.flops = 2 * out.numel() * 343
        ctx.layer = layer
        ctx.save_for_backward(image)
        return out

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        (image,) = ctx.saved_tensors
        dy = dy.contiguous()
        unfolded = ops.stem_expand(image, dtype=ACT)  # [N, D, H', W', kh*8 + kw]
        wp = layer.wgrad[0].get((unfolded.data_ptr(), dy.data_ptr()) + tuple(image.shape), lambda: Conv3dWgradPlan(
            unfolded, dy, kernel=(7, 1, 1), stride=(2, 1, 1), dilation=1, padding=(3, 0, 0)))
        dwp = wp.run()  # [64, 64 pseudo-channels, 7, 1, 1]
        dw = dwp.view(64, 8, 8, 7)[:, :7, :7, :].permute(0, 3, 1, 2).reshape(64, 1, 7, 7, 7).contiguous()
        return None, dw, None


class TrainableMed3D:
    """Functional train-mode forward of a drop-in seg-reg network (basic or bottleneck blocks) with gradients."""

    def __init__(self, model, glue="native", sync_bn=None):
        """glue: "native" = train-mode BatchNorm (+ReLU, +residual) on the K10 kernels; "aten" = the same through ATen
        (kept for A/B checks).  sync_bn: None/False = per-rank statistics, True or a process group = SyncBatchNorm
        (train.py:101)."""
        if getattr(model, "head_kind", "reg") != "reg":
            raise ValueError("TrainableMed3D: the regression networks (med3ddram*) are the trained ones (train.py:72)")
        if glue not in ("native", "aten"):
            raise ValueError(f"glue must be 'native' or 'aten', got {glue!r}")
        self.model = model
        self.layers = {}
        self.glue, self.sync_bn = glue, sync_bn
        self.bn_sinks = {}  # id(BatchNorm3d) -> (flat gradient view [dbeta | dgamma], callback); set by TrainStep
        self._counters = [m.num_batches_tracked for m in model.modules()
                          if isinstance(m, torch.nn.BatchNorm3d) and m.num_batches_tracked is not None]

    def training_flops(self):
        """Algorithmic FLOPs of one training step of the last forward: fprop + dgrad + wgrad of every convolution
        (2*M*N*K each), the stem without a data gradient."""
        total = 0
        for name, lay in self.layers.items():
            total += lay.flops * (2 if name == "conv1" else 3)
        return total

    def _layer(self, name, conv):
        lay = self.layers.get(name)
        if lay is None:
            lay = ConvLayer(name, tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.dilation))
            self.layers[name] = lay
        return lay

    def _conv(self, name, conv, x1, x2=None):
        return ConvFn.apply(self._layer(name, conv), conv.weight, conv.bias, x1, x2)

    def _bn(self, bn, x, relu=True, res=None):
        """relu(bn(x) (+ res)) in train mode."""
        if self.glue == "native":
            return BatchNormTrainFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, res, relu, bn.eps,
                                          BN_MOMENTUM, self.sync_bn, self.bn_sinks.get(id(bn)))
        y = F.batch_norm(_ncdhw(x), bn.running_mean, bn.running_var, bn.weight, bn.bias, True, BN_MOMENTUM, bn.eps)
        y = _ndhwc(y)
        if res is not None:
            y = y + res
        return torch.relu(y) if relu else y

    def _block(self, name, blk, x, planes_out):
        if hasattr(blk, "conv3"):  # Bottleneck, med3d.py:164-184
            out = self._bn(blk.bn1, self._conv(name + ".conv1", blk.conv1, x))
            out = self._bn(blk.bn2, self._conv(name + ".conv2", blk.conv2, out))
            last_bn, last = blk.bn3, self._conv(name + ".conv3", blk.conv3, out)
            stride = blk.conv2.stride[0]
        else:  # BasicBlock, med3d.py:129-144
            out = self._bn(blk.bn1, self._conv(name + ".conv1", blk.conv1, x))
            last_bn, last = blk.bn2, self._conv(name + ".conv2", blk.conv2, out)
            stride = blk.conv1.stride[0]
        res = x
        if stride != 1 or x.shape[4] != planes_out:  # shortcut type A (med3d.py:103-112): no gradient (out.data)
            res = x.detach()[:, ::stride, ::stride, ::stride, :]
            res = F.pad(res, (0, planes_out - res.shape[4])).contiguous()
        return self._bn(last_bn, last, relu=True, res=res)  # relu(bn(conv) + residual), fused

    def _up(self, name, us, x, skip):
        if self.glue == "native":
            up = Upsample2xFn.apply(x)
        else:
            up = _ndhwc(F.interpolate(_ncdhw(x).float(), scale_factor=2, mode="trilinear", align_corners=True).to(ACT))
        if tuple(up.shape[1:4]) != tuple(skip.shape[1:4]):
            raise ValueError(f"{name}: up-sampled {tuple(up.shape)} vs skip {tuple(skip.shape)} (input dims must be 8k)")
        y = self._bn(us.conv_blocks[0][1], self._conv(name + ".conv_blocks.0.0", us.conv_blocks[0][0], up, skip))
        return self._bn(us.conv_blocks[1][1], self._conv(name + ".conv_blocks.1.0", us.conv_blocks[1][0], y))

    def forward(self, image, lungs=None, with_regs=True):
        """image [B, D, H, W] fp32 (CUDA), lungs [B, D, H, W] float 0/1 or None ->
        (dense_outs: 2 x fp32 [B, 1, D/2, H/2, W/2], reg_outs: 2 x fp32 [B]) as med3d.py:369-388.
        `with_regs=False` skips the lobe-masked means (K11 computes them together with the loss) and returns None."""
        m = self.model
        B = image.shape[0]
        if self._counters:  # every BatchNorm of the network runs exactly once per forward: one launch for all counters
            torch._foreach_add_(self._counters, 1)
        with torch.backends.cudnn.flags(enabled=False):
            c1 = StemFn.apply(self._layer("conv1", m.conv1), m.conv1.weight, image.float().contiguous())
            x = self._bn(m.bn1, c1)
            if self.glue == "native":
                y = MaxPool3dFn.apply(x)
            else:
                y = _ndhwc(F.max_pool3d(_ncdhw(x), kernel_size=3, stride=2, padding=1))
            feats = []
            for li, stage in enumerate((m.layer1, m.layer2, m.layer3, m.layer4), start=1):
                planes_out = LAYER_CFG[li - 1][0] * m.expansion
                for bi, blk in enumerate(stage):
                    y = self._block(f"layer{li}.{bi}", blk, y, planes_out)
                feats.append(y)
            xup1 = self._up("us1", m.us1, feats[3], feats[0])
            xup2 = self._up("us2", m.us2, xup1, x)
            xup3 = self._bn(m.us3[1], self._conv("us3.0", m.us3[0], xup2))
            if self.glue == "native" and len(m.fcs) == 2 and all(fc.weight.shape[0] == 1 for fc in m.fcs):
                dense = list(HeadsSigmoidFn.apply(xup3, m.fcs[0].weight, m.fcs[0].bias, m.fcs[1].weight, m.fcs[1].bias))
            else:
                v = xup3.float()  # [B, D, H, W, 32]; the 1x1x1 heads (med3d.py:329-332, 382) are a 32-long dot per voxel
                dense = [torch.sigmoid(v @ fc.weight.view(fc.weight.shape[0], 32).t() + fc.bias).permute(0, 4, 1, 2, 3)
                         for fc in m.fcs]
            if not with_regs:
                return dense, None
            if lungs is None:
                mask = torch.ones((B, 1) + tuple(dense[0].shape[2:]), device=image.device)
            else:
                mask = F.interpolate(lungs.view(B, 1, *lungs.shape[-3:]).float(), dense[0].shape[-3:], mode="nearest")
            regs = [(d * mask).view(B, -1).sum(dim=-1) / mask.view(B, -1).sum(dim=-1) for d in dense]
        return dense, regs


# ------------------------------------------------------------------------------------------
# losses of ScanRegLightningModule (models.py:495-518, metrics.py)
# ------------------------------------------------------------------------------------------
BETA, GAMMA = 0.7338, 0.2578  # models.py:412-413


def regression_label_bands(labels, ratio_mapping, tightness=1.0):
    """models.py:477-493: class index -> (lo, hi) band of the lesion ratio; class 0 collapses to (0, 0)."""
    bands = []
    for c in labels:
        lo, hi = ratio_mapping[int(c)]
        if lo < 1e-7:
            bands.append((0.0, 0.0))
        else:
            mid, span = (lo + hi) / 2.0, (hi - lo) * tightness / 2.0
            bands.append((mid - span, mid + span))
    return torch.tensor(bands, dtype=torch.float32)


def interval_regression_loss(outs, bands, weights):
    """models.py:495-506."""
    d = torch.cat([outs.unsqueeze(1), bands], dim=1)
    d = BETA * d ** GAMMA
    K = (0.5 * (d[:, 2] - d[:, 1])) ** 2
    unhinged = (d[:, 0] - (d[:, 2] + d[:, 1]) / 2.0) ** 2 - K
    return (10.0 * F.leaky_relu(unhinged, negative_slope=0.0) * weights).sum()


def dice_coef(y, y_hat, smooth=1e-7):
    """metrics.py:33-37."""
    inter = (y_hat.reshape(-1) * y.reshape(-1)).sum()
    return (2.0 * inter + smooth) / (y.sum() + y_hat.sum() + smooth)


def masked_bce(y, y_hat, mask, smoothness=0.85, eps=1e-6):
    """metrics.py:10-30 (BinaryCrossEntropy.__call__ with a mask)."""
    t = y.float()
    alpha = (1.0 - t.sum() / t.shape[0]).clamp(0.3, 0.7)
    pt = y_hat * t + (1.0 - y_hat) * (1.0 - t)
    w = alpha * t + (1.0 - alpha) * (1.0 - t)
    logp = torch.log(pt.clamp(eps, 1.0 - eps))
    nll = -1.0 * (smoothness * logp * w * mask + logp * w * (1.0 - mask))
    return nll.sum() / w.sum()


def training_loss(dense, regs, lungs, ems, cle_labels, pse_labels, cle_bands, pse_bands, cle_weights, pse_weights):
    """models.py:547-565: loss_cle + loss_pse + 2 * mul_loss + seg_loss.  lungs/ems [B, D, H, W] float 0/1."""
    B = lungs.shape[0]
    loss_cle = interval_regression_loss(regs[0], cle_bands, cle_weights)
    loss_pse = interval_regression_loss(regs[1], pse_bands, pse_weights)
    binary = torch.logical_or(cle_labels > 0, pse_labels > 0).float().view(B, 1, 1, 1, 1)
    size = dense[0].shape[-3:]
    seg_labels = F.interpolate(ems.view(B, 1, *ems.shape[-3:]).float() * binary, size, mode="nearest").detach()
    lung_labels = F.interpolate(lungs.view(B, 1, *lungs.shape[-3:]).float(), size, mode="nearest")
    mul_loss = dice_coef(dense[0] * lung_labels, dense[1] * lung_labels)
    both = torch.clamp(dense[0] + dense[1], min=0.0, max=1.0)
    seg_loss = masked_bce(seg_labels, both, lung_labels, smoothness=0.85)
    return loss_cle + loss_pse + 2.0 * mul_loss + seg_loss


class TrainStep:
    """forward -> loss -> backward -> gradient average over the process group -> Adam (models.py:685-698, lr from args).

    Parameters and gradients live in two flat fp32 buffers with the same layout (`GradBuckets`, parameters in reverse
    registration order, i.e. roughly the order backward produces them); every `nn.Parameter` of the network becomes a
    view into the parameter buffer, so `state_dict()` / `load_state_dict()` keep working.  K9 writes each convolution's
    weight gradient straight into its slice; the remaining gradients arrive through autograd's AccumulateGrad.  A
    bucket's (asynchronous, NCCL) all-reduce is launched the moment its last gradient lands, so the exchange overlaps
    the rest of the backward pass.  `sync_bn=True` (or "nccl") all-reduces the BatchNorm statistics as well
    (train.py:101) with NCCL; `sync_bn="peer"` does that exchange with K10x, one kernel over NVLink peer memory per
    BatchNorm pass (`backward.PeerExchange`; one node, <= 8 ranks; `self.peer.check()` reports a lost peer).

    loss = "native" (K11, `TrainLossFn`) | "aten" (`training_loss`); optimizer = "native" (K12, `FlatAdam`) | "torch"
    (torch.optim.Adam).  `last` holds the loss terms and the lobe-masked means of the last step (device tensors).
    """

    def __init__(self, model, lr=1e-4, bucket_bytes=64 << 20, group=None, sync_bn=False, loss="native", optimizer="native"):
        if loss not in ("native", "aten") or optimizer not in ("native", "torch"):
            raise ValueError(f"TrainStep: loss={loss!r} / optimizer={optimizer!r} (native|aten, native|torch)")
        if sync_bn not in (False, True, None, "nccl", "peer"):
            raise ValueError(f"TrainStep: sync_bn={sync_bn!r} (False | True/'nccl' | 'peer')")
        self.peer = PeerExchange(group) if sync_bn == "peer" else None  # collective: every rank builds it here
        if self.peer is not None:
            bn_group = self.peer
        else:
            bn_group = (group if group is not None else True) if sync_bn else None
        self.net = TrainableMed3D(model, sync_bn=bn_group)
        self.params = [(n, p) for n, p in model.named_parameters()]
        dev = self.params[0][1].device
        self.buckets = GradBuckets([(n, tuple(p.shape)) for n, p in reversed(self.params)], dev, bucket_bytes)
        self.flat_param = torch.empty_like(self.buckets.flat)
        self._bucket_size = [0] * self.buckets.num_buckets
        # convolutions that run through ConvFn: everything but the stem (StemFn re-lays its gradient out with ATen) and
        # the two 1x1x1 heads (HeadsSigmoidFn)
        conv_weights = {name + ".weight": conv for name, conv in model.named_modules()
                        if isinstance(conv, torch.nn.Conv3d) and name != "conv1" and not name.startswith("fcs.")}
        # BatchNorm: bias and weight are neighbours in the flat layout (reverse registration order), and K10's backward
        # reduction yields [dbeta | dgamma] in one vector: one copy fills both gradients
        bn_pairs = {}
        for name, mod in model.named_modules():
            if isinstance(mod, torch.nn.BatchNorm3d) and mod.weight is not None and mod.bias is not None:
                ob, nb, _ = self.buckets.slices[name + ".bias"]
                ow, nw, _ = self.buckets.slices[name + ".weight"]
                if ob + nb == ow and nb == nw:
                    bn_pairs[name + ".bias"] = bn_pairs[name + ".weight"] = (mod, ob, nb + nw)
        for n, p in self.params:
            off, numel, shape = self.buckets.slices[n]
            view = self.flat_param[off:off + numel].view(shape)
            view.copy_(p.data)
            p.data = view  # the parameter now lives in the flat buffer (same values, same shape)
            p.grad = self.buckets.view(n)
            b = self.buckets.bucket_of[n]
            self._bucket_size[b] += 1
            conv = conv_weights.get(n)
            if conv is not None:
                self.net._layer(n[:-len(".weight")], conv).grad_sink = (p.grad, self._make_ready(b))
            elif n in bn_pairs:
                mod, start, length = bn_pairs[n]
                if n.endswith(".weight"):  # registered once per pair; the callback counts both parameters down
                    pair = (self.buckets.bucket_of[n[:-len("weight")] + "bias"], b)
                    self.net.bn_sinks[id(mod)] = (self.buckets.flat[start:start + length], self._make_ready_many(pair))
            else:
                p.register_post_accumulate_grad_hook(self._make_hook(n, b))
        self._pending = list(self._bucket_size)
        self._launched = [False] * self.buckets.num_buckets
        self.loss_kind = loss
        if optimizer == "native":
            self.opt = FlatAdam(self.flat_param, self.buckets.flat, lr=lr)
        else:
            self.opt = torch.optim.Adam([p for _, p in self.params], lr=lr)
        self.group = group
        self.last = {}

    def _grad_ready(self, bucket):
        self._pending[bucket] -= 1
        if self._pending[bucket] == 0:
            self._launched[bucket] = True
            self.buckets.reduce_bucket(bucket, self.group)

    def _make_ready(self, bucket):
        return lambda: self._grad_ready(bucket)

    def _make_ready_many(self, buckets):
        def ready():
            for b in buckets:
                self._grad_ready(b)
        return ready

    def _make_hook(self, name, bucket):
        def hook(param):
            view = self.buckets.view(name)
            if param.grad is not None and param.grad.data_ptr() != view.data_ptr():  # autograd replaced the tensor
                view.copy_(param.grad)
                param.grad = view
            self._grad_ready(bucket)
        return hook

    def zero_grad(self):
        self.buckets.flat.zero_()
        self._pending = list(self._bucket_size)
        self._launched = [False] * self.buckets.num_buckets

    def decay_lr(self, gamma=0.95):
        """ExponentialLR(gamma=0.95), stepped once per epoch by Lightning (models.py:694-697)."""
        if isinstance(self.opt, FlatAdam):
            self.opt.decay_lr(gamma)
        else:
            for g in self.opt.param_groups:
                g["lr"] *= gamma

    def step(self, batch, cle_bands, pse_bands, cle_weights, pse_weights):
        last_name, last_param = self.params[-1]
        if last_param.data_ptr() != self.flat_param.data_ptr():
            raise RuntimeError("TrainStep: the network's parameters no longer live in this step's flat buffers (the model "
                               "was moved or re-typed after TrainStep was built); build a new TrainStep")
        self.zero_grad()
        if self.loss_kind == "native":
            dense, _ = self.net.forward(batch["image"], None, with_regs=False)
            loss, terms, regs = TrainLossFn.apply(dense[0], dense[1], batch["lung_mask"], batch["em_mask"], batch["cls_label"],
                                                  batch["pse_label"], cle_bands, pse_bands, cle_weights, pse_weights,
                                                  BETA, GAMMA)
            self.last = {"loss_cle": terms[0], "loss_pse": terms[1], "mul_loss": terms[2], "seg_loss": terms[3],
                         "reg_outs": [regs[:, 0], regs[:, 1]]}
        else:
            lungs = batch["lung_mask"].float()
            dense, regs = self.net.forward(batch["image"], lungs)
            loss = training_loss(dense, regs, lungs, batch["em_mask"].float(), batch["cls_label"], batch["pse_label"],
                                 cle_bands, pse_bands, cle_weights, pse_weights)
            self.last = {"reg_outs": [r.detach() for r in regs]}
        loss.backward()
        for i in range(self.buckets.num_buckets):  # buckets holding a parameter that received no gradient
            if not self._launched[i]:
                self.buckets.reduce_bucket(i, self.group)
        self.buckets.wait()
        self.opt.step()
        self.steps_done = getattr(self, "steps_done", 0) + 1
        if self.peer is not None and self.steps_done % 256 == 0:
            self.peer.check()  # one 4-byte read-back every 256 steps: ranks must not drift apart on local statistics
        # the kernels changed the parameters behind autograd's back: bump the version counters, and tell the
        # inference engines of the module to re-pack (med3d._Med3DSegNet.weights_epoch)
        torch.autograd.graph.increment_version([p for _, p in self.params])
        mark = getattr(self.net.model if hasattr(self.net, "model") else self.net, "mark_weights_changed", None)
        if mark is not None:
            mark()
        return loss.detach()
