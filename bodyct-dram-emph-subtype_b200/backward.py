"""Training-path wrappers: convolution backward, train-mode BatchNorm, pooling / up-sampling adjoints,
heads, loss (K11), Adam (K12), the SyncBatchNorm peer exchange (K10x) and the gradient buckets.

The reference has no backward code: `train.py` lets Lightning's automatic optimisation call
`loss.backward()` (models.py:495-582), so autograd differentiates every `nn.Conv3d` of med3d.py
(conv3x3x3 :93-100, the decoder convs :67/:76, the bottleneck convs :152-157).  These wrappers produce
what autograd leaves behind for one convolution:

* `Conv3dWgradPlan`  — `weight.grad` (K9, `dram_conv3d_wgrad_*`: tcgen05 GEMM with the voxels as K);
* `Conv3dDgradPlan`  — the gradient of the input: a transposed convolution, which for stride 1 *is* a K1
  convolution of `dy` with the spatially flipped, channel-transposed weights (`pack_dgrad_weight`), so it
  runs on the forward kernels (plane-ring kernel for Cin <= 64, tile kernel otherwise); stride 2 first
  scatters `dy` onto the stride lattice of a zeroed buffer (one convolution of the network, layer2.0.conv1);
* `GradBuckets` — flat fp32 gradient storage that `Conv3dWgradPlan` writes into, all-reduced per bucket
  over the process group (NCCL over NVLink on the GPUs; data-parallel training, train.py:101 `ddp`).

Activations and their gradients are NDHWC 16-bit (bf16 recommended for gradients); weight gradients fp32.
There is no CPU path.
"""
import ctypes as C

import torch

from . import _capi
from ._capi import ConvDesc, check
from . import ops
from .ops import ACT_DTYPES, Conv3dPlan, _need, _need16, _p, _stream, _triple, pack_conv_weight


def _geometry(kernel, stride, dilation, padding):
    k, s, dl = _triple(kernel), _triple(stride), _triple(dilation)
    pad = tuple(dl[i] * (k[i] - 1) // 2 for i in range(3)) if padding is None else _triple(padding)
    return k, s, dl, pad


def conv_out_size(size, k, s, dl, pad):
    return tuple((size[i] + 2 * pad[i] - dl[i] * (k[i] - 1) - 1) // s[i] + 1 for i in range(3))


class Conv3dWgradPlan:
    """`dw[Cout, cin_total, kd, kh, kw]` fp32 (PyTorch layout) from x [N,D,H,W,Cin] and dy [N,Do,Ho,Wo,Cout].

    `dw` may be a view into a flat gradient bucket (contiguous); `cin_offset`/`cin_total` select the channel
    range this source fills, so the concatenated decoder input (med3d.py:87) is two plans on one `dw`.
    Two kernels sit behind the entry point: the plane kernel (3x3x3, stride 1, dilation 1, Cout <= 64: layer1 and the
    decoder) and the streaming kernel (everything else, or `algo="tiles"`).
    """

    def __init__(self, x, dy, *, dw=None, kernel=3, stride=1, dilation=1, padding=None, cin_total=None,
                 cin_offset=0, algo="auto"):
        lib = _capi.load()
        _need16(x, "conv3d_wgrad x", 5)
        _need(dy, x.dtype, "conv3d_wgrad dy", 5)
        n, di, hi, wi, cin = x.shape
        cout = dy.shape[4]
        k, s, dl, pad = _geometry(kernel, stride, dilation, padding)
        want = (n,) + conv_out_size((di, hi, wi), k, s, dl, pad) + (cout,)
        if tuple(dy.shape) != want:
            raise ValueError(f"conv3d_wgrad: dy shape {tuple(dy.shape)} != {want}")
        cin_total = cin if cin_total is None else cin_total
        shape = (cout, cin_total) + k
        if dw is None:
            dw = torch.zeros(shape, dtype=torch.float32, device=x.device)
        else:
            _need(dw, torch.float32, "conv3d_wgrad dw", 5)
            if tuple(dw.shape) != shape:
                raise ValueError(f"conv3d_wgrad: dw shape {tuple(dw.shape)} != {shape}")
        d = ConvDesc()
        d.n, d.di, d.hi, d.wi, d.c1, d.c2, d.cout = n, di, hi, wi, cin, 0, cout
        d.kd, d.kh, d.kw = k
        d.sd, d.sh, d.sw = s
        d.dd, d.dh, d.dw = dl
        d.pd, d.ph, d.pw = pad
        d.dtype = ACT_DTYPES[x.dtype]
        d.algo = _capi.CONV_ALGO[algo]  # "tiles" forces the streaming kernel where the plane kernel would be picked
        nbytes = lib.dram_conv3d_wgrad_workspace_bytes(C.byref(d))
        if nbytes < 0:
            raise _capi.DramError(f"dram_conv3d_wgrad_workspace_bytes: {_capi.last_error()}")
        self.workspace = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=x.device)
        handle = C.c_void_p()
        check(lib.dram_conv3d_wgrad_plan_create(C.byref(d), _p(x), _p(dy), _p(dw), cin_total, cin_offset,
                                                _p(self.workspace), int(nbytes), C.byref(handle)),
              "dram_conv3d_wgrad_plan_create")
        self._handle, self._lib = handle, lib
        self._keep = (x, dy, dw)
        self.dw = dw
        flops, items, ks, bn = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.dram_conv3d_wgrad_plan_info(handle, C.byref(flops), C.byref(items), C.byref(ks), C.byref(bn)),
              "dram_conv3d_wgrad_plan_info")
        self.flops, self.items, self.kslices, self.block_n = flops.value, items.value, ks.value, abs(bn.value)
        self.algo = "planes" if bn.value < 0 else "stream"

    def run(self, accumulate=False, max_ctas=0):
        check(self._lib.dram_conv3d_wgrad_run(self._handle, 1 if accumulate else 0, max_ctas, _stream()),
              "dram_conv3d_wgrad_run")
        return self.dw

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                self._lib.dram_conv3d_wgrad_plan_destroy(h)
            except Exception:
                pass
            self._handle = None


def pack_dgrad_weight(weight, dtype=torch.bfloat16, cin_range=None, pad_cout_to=64):
    """[Cout, Cin, kd, kh, kw] -> the packed K1 weight of the transposed convolution:
    rows = input channels (optionally the slice `cin_range` of a concatenated input), K = flipped taps x Cout
    (Cout zero-padded to a multiple of `pad_cout_to`, the channel granularity of the forward kernels)."""
    w = weight.detach().to(torch.float32)
    if cin_range is not None:
        w = w[:, cin_range[0]:cin_range[1]]
    cout = w.shape[0]
    padc = (-cout) % pad_cout_to
    if padc:
        w = torch.cat([w, w.new_zeros((padc,) + tuple(w.shape[1:]))], dim=0)
    wt = w.flip(2, 3, 4).permute(1, 0, 2, 3, 4).contiguous()  # [Cin, Cout, kd, kh, kw], taps reversed
    return pack_conv_weight(wt, dtype=dtype)


def pack_weight_into(weight, out, *, transpose=False, cin_range=None):
    """Device-side re-packing of an fp32 Conv3d weight into an existing 16-bit operand buffer (no temporaries):
    `transpose=False` -> K1's [Cout, taps*Cin] (= ops.pack_conv_weight), `transpose=True` -> the data-gradient operand
    [Cin, taps*Cout_pad] (= pack_dgrad_weight).  Used by the training step, whose weights change every iteration."""
    w = weight.detach()
    _need(w, torch.float32, "pack_weight_into weight", 5)
    _need16(out, "pack_weight_into out", 2)
    cout, cin_total = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    c0, c1 = (0, cin_total) if cin_range is None else cin_range
    if transpose:
        cout_pad = out.shape[1] // taps
        want = (c1 - c0, taps * cout_pad)
    else:
        cout_pad = cout
        want = (cout, taps * (c1 - c0))
    if tuple(out.shape) != want or (transpose and cout_pad < cout):
        raise ValueError(f"pack_weight_into: out {tuple(out.shape)} does not fit weight {tuple(w.shape)} (want {want})")
    check(_capi.load().dram_pack_conv_weight(_p(w), _p(out), cout, cin_total, taps, c0, c1 - c0, 1 if transpose else 0,
                                             cout_pad, ACT_DTYPES[out.dtype], _stream()), "dram_pack_conv_weight")
    return out


class Conv3dDgradPlan:
    """dx [N, D, H, W, Cin] (16-bit NDHWC) = conv_transpose3d(dy, weight) for the forward geometry given.

    `weight` is the forward convolution's [Cout, Cin_total, kd, kh, kw] parameter (fp32); `cin_range` picks the
    channels of one source of a concatenated input.  The transposed convolution runs on the K1 kernels.
    """

    def __init__(self, dy, weight, in_size, *, kernel=3, stride=1, dilation=1, padding=None, cin_range=None,
                 out=None, algo="auto"):
        _need16(dy, "conv3d_dgrad dy", 5)
        k, s, dl, pad = _geometry(kernel, stride, dilation, padding)
        n, do, ho, wo, cout = dy.shape
        in_size = tuple(in_size)
        if (do, ho, wo) != conv_out_size(in_size, k, s, dl, pad):
            raise ValueError(f"conv3d_dgrad: dy {tuple(dy.shape)} is not the output of a conv over {in_size}")
        if weight.shape[0] != cout or tuple(weight.shape[2:]) != k:
            raise ValueError(f"conv3d_dgrad: weight {tuple(weight.shape)} does not match dy / kernel")
        self.packed = pack_dgrad_weight(weight, dtype=dy.dtype, cin_range=cin_range).to(dy.device)
        cin = self.packed.shape[0]
        self.dy = dy
        cpad = (-cout) % 64
        # the K1 convolution reads `src`: dy itself, or dy scattered onto the stride lattice / channel-padded
        lattice = tuple(in_size[i] + 2 * pad[i] - dl[i] * (k[i] - 1) for i in range(3))
        if s != (1, 1, 1) or cpad:
            self.src = torch.zeros((n,) + lattice + (cout + cpad,), dtype=dy.dtype, device=dy.device)
            self._scatter = self.src[:, ::s[0], ::s[1], ::s[2], :cout][:, :do, :ho, :wo]
        else:
            self.src, self._scatter = dy, None
        tpad = tuple(dl[i] * (k[i] - 1) - pad[i] for i in range(3))
        if min(tpad) < 0:
            raise ValueError("conv3d_dgrad: padding larger than dilation*(k-1) is not supported")
        self.bias = torch.zeros(cin, dtype=torch.float32, device=dy.device)
        self.plan = Conv3dPlan(self.src, self.packed, self.bias, kernel=k, stride=1, dilation=dl, padding=tpad,
                               relu=False, out=out, algo=algo)
        if tuple(self.plan.out_shape[1:4]) != in_size:
            raise AssertionError(f"conv3d_dgrad: internal size mismatch {self.plan.out_shape} vs {in_size}")
        self.out = self.plan.out
        # algorithmic FLOPs of the transposed convolution (not of the zero-inserted one that runs)
        self.flops = 2 * n * do * ho * wo * cout * cin * k[0] * k[1] * k[2]

    def run(self, max_ctas=0):
        if self._scatter is not None:
            self._scatter.copy_(self.dy)
        return self.plan.run(max_ctas)


class Upsample2xFn(torch.autograd.Function):
    """x2 trilinear up-sampling, align_corners=True (med3d.py:83, 86) on NDHWC 16-bit: K4 forward, K4T backward."""

    @staticmethod
    def forward(ctx, x):
        ctx.in_shape = tuple(x.shape)
        return ops.upsample2x(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        n, d, h, w, c = ctx.in_shape
        dy = dy.contiguous()
        dx = torch.empty(ctx.in_shape, dtype=dy.dtype, device=dy.device)
        check(_capi.load().dram_upsample2x_backward(_p(dy), _p(dx), n, d, h, w, c, ACT_DTYPES[dy.dtype], _stream()),
              "dram_upsample2x_backward")
        return dx


class MaxPool3dFn(torch.autograd.Function):
    """MaxPool3d(3, stride 2, padding 1) (med3d.py:305) on NDHWC 16-bit: K3 forward, K3T backward (first maximum of a
    window takes the gradient, as ATen)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.maxpool3d(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, d, h, w, c = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        lib = _capi.load()
        ws = torch.empty(int(lib.dram_maxpool3d_backward_workspace_bytes(n, d, h, w, c)), dtype=torch.uint8, device=x.device)
        check(lib.dram_maxpool3d_backward(_p(x), _p(dy), _p(dx), _p(ws), n, d, h, w, c, ACT_DTYPES[x.dtype], _stream()),
              "dram_maxpool3d_backward")
        return dx


class HeadsSigmoidFn(torch.autograd.Function):
    """The two 1x1x1 regression heads + sigmoid (med3d.py:329-332, 382) on the 32-channel NDHWC feature map:
    (x [B,D,H,W,32], w0 [1,32,1,1,1], b0 [1], w1, b1) -> two fp32 maps [B,1,D,H,W] (K5T)."""

    @staticmethod
    def forward(ctx, x, w0, b0, w1, b1):
        lib = _capi.load()
        _need16(x, "heads x", 5)
        if x.shape[4] != 32 or w0.numel() != 32 or w1.numel() != 32:
            raise ValueError(f"heads: expected 32 input channels and two 1-channel heads, got x {tuple(x.shape)}")
        B, D, H, W, _ = x.shape
        m = B * D * H * W
        w = torch.cat([w0.detach().reshape(1, 32), w1.detach().reshape(1, 32)]).float().contiguous()
        b = torch.cat([b0.detach().reshape(1), b1.detach().reshape(1)]).float().contiguous()
        out0 = torch.empty((B, 1, D, H, W), dtype=torch.float32, device=x.device)
        out1 = torch.empty_like(out0)
        check(lib.dram_heads_sigmoid_forward(_p(x), _p(w), _p(b), _p(out0), _p(out1), m, ACT_DTYPES[x.dtype], _stream()),
              "dram_heads_sigmoid_forward")
        ctx.save_for_backward(x, w, out0, out1)
        ctx.shapes = (tuple(w0.shape), tuple(b0.shape), tuple(w1.shape), tuple(b1.shape))
        return out0, out1

    @staticmethod
    def backward(ctx, g0, g1):
        lib = _capi.load()
        x, w, s0, s1 = ctx.saved_tensors
        m = s0.numel()
        g0 = (torch.zeros_like(s0) if g0 is None else g0).float().contiguous()
        g1 = (torch.zeros_like(s1) if g1 is None else g1).float().contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty((2, 32), dtype=torch.float32, device=x.device)
        db = torch.empty(2, dtype=torch.float32, device=x.device)
        ws = torch.empty(int(lib.dram_heads_workspace_bytes()), dtype=torch.uint8, device=x.device)
        check(lib.dram_heads_sigmoid_backward(_p(x), _p(w), _p(s0), _p(s1), _p(g0), _p(g1), _p(dx), _p(dw), _p(db), _p(ws), m,
                                              ACT_DTYPES[x.dtype], _stream()), "dram_heads_sigmoid_backward")
        sw0, sb0, sw1, sb1 = ctx.shapes
        return dx, dw[0].reshape(sw0), db[0:1].reshape(sb0), dw[1].reshape(sw1), db[1:2].reshape(sb1)


class TrainLossFn(torch.autograd.Function):
    """K11: loss of `ScanRegLightningModule.shared_step(TRAIN)` (models.py:547-565) and its gradient on the two dense
    maps, fused (`dram_train_loss_forward/backward`).

    (cle_map, pse_map [B,1,D2,H2,W2] fp32; lungs, ems [B,D,H,W] 0/1 bytes or bool; cle_labels, pse_labels int64 [B];
    cle_bands, pse_bands fp32 [B,2]; cle_weights, pse_weights fp32 [B]) ->
    (loss scalar, terms fp32 [4] = loss_cle, loss_pse, mul_loss, seg_loss, regs fp32 [B,2] = the lobe-masked means of
    med3d.py:387).  Only `loss` carries a gradient."""

    @staticmethod
    def forward(ctx, cle_map, pse_map, lungs, ems, cle_labels, pse_labels, cle_bands, pse_bands, cle_weights, pse_weights,
                beta=0.7338, gamma=0.2578):
        lib = _capi.load()
        _need(cle_map, torch.float32, "train_loss cle_map", 5)
        _need(pse_map, torch.float32, "train_loss pse_map", 5)
        if cle_map.shape != pse_map.shape or cle_map.shape[1] != 1:
            raise ValueError(f"train_loss: maps {tuple(cle_map.shape)} / {tuple(pse_map.shape)} must both be [B,1,D2,H2,W2]")
        dev = cle_map.device
        lungs, ems = _mask_u8(lungs, "train_loss lungs"), _mask_u8(ems, "train_loss ems")
        B, _, d2, h2, w2 = cle_map.shape
        if lungs.shape != ems.shape or lungs.shape[0] != B:
            raise ValueError(f"train_loss: masks {tuple(lungs.shape)} / {tuple(ems.shape)} vs batch {B}")
        d, h, w = lungs.shape[1:]

        def vec(t, dtype, shape, what):
            t = t.to(device=dev, dtype=dtype).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"train_loss: {what} has shape {tuple(t.shape)}, expected {shape}")
            return t

        cl, pl = vec(cle_labels, torch.int64, (B,), "cle_labels"), vec(pse_labels, torch.int64, (B,), "pse_labels")
        cb, pb = vec(cle_bands, torch.float32, (B, 2), "cle_bands"), vec(pse_bands, torch.float32, (B, 2), "pse_bands")
        cw, pw = vec(cle_weights, torch.float32, (B,), "cle_weights"), vec(pse_weights, torch.float32, (B,), "pse_weights")
        coef = torch.empty(_capi.LOSS_COEF_HEAD + 4 * B, dtype=torch.float32, device=dev)
        ws = torch.empty(int(lib.dram_train_loss_workspace_bytes(B)), dtype=torch.uint8, device=dev)
        check(lib.dram_train_loss_forward(_p(cle_map), _p(pse_map), _p(lungs), _p(ems), _p(cl), _p(pl), _p(cb), _p(pb), _p(cw),
                                          _p(pw), B, d, h, w, d2, h2, w2, beta, gamma, _p(coef), _p(ws), _stream()),
              "dram_train_loss_forward")
        ctx.save_for_backward(cle_map, pse_map, lungs, ems, cl, pl, coef)
        ctx.geom = (B, d, h, w, d2, h2, w2)
        loss, terms = coef[0].clone(), coef[1:5].clone()
        regs = coef[_capi.LOSS_COEF_HEAD:].view(B, 4)[:, :2].clone()
        ctx.mark_non_differentiable(terms, regs)
        return loss, terms, regs

    @staticmethod
    def backward(ctx, gloss, _gterms, _gregs):
        cle_map, pse_map, lungs, ems, cl, pl, coef = ctx.saved_tensors
        g0, g1 = torch.empty_like(cle_map), torch.empty_like(pse_map)
        gl = gloss.to(torch.float32).contiguous()
        check(_capi.load().dram_train_loss_backward(_p(cle_map), _p(pse_map), _p(lungs), _p(ems), _p(cl), _p(pl), _p(coef),
                                                    _p(gl), *ctx.geom, _p(g0), _p(g1), _stream()),
              "dram_train_loss_backward")
        return (g0, g1) + (None,) * 10


def _mask_u8(mask, what):
    """bool / uint8 0-1 CUDA mask [B,D,H,W] -> contiguous uint8 view (no copy for bool)."""
    if not mask.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (there is no CPU path)")
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    elif mask.dtype != torch.uint8:
        mask = (mask != 0).contiguous().view(torch.uint8)
    if mask.dim() == 5 and mask.shape[1] == 1:
        mask = mask[:, 0]
    if mask.dim() != 4:
        raise ValueError(f"{what}: expected [B,D,H,W], got {tuple(mask.shape)}")
    return mask.contiguous()


class FlatAdam:
    """K12: torch.optim.Adam(params, lr) of `configure_optimizers` (models.py:685-698) as one launch per step over flat
    fp32 buffers.  `flat_param` / `flat_grad` hold every parameter / gradient back to back (the parameters of the
    network are views into `flat_param`); exp_avg / exp_avg_sq live here.  `lr` may be changed between steps (the
    reference's ExponentialLR(gamma=0.95) steps once per epoch: `decay_lr()`)."""

    def __init__(self, flat_param, flat_grad, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        _need(flat_param, torch.float32, "FlatAdam flat_param", 1)
        _need(flat_grad, torch.float32, "FlatAdam flat_grad", 1)
        if flat_param.shape != flat_grad.shape:
            raise ValueError(f"FlatAdam: {tuple(flat_param.shape)} parameters vs {tuple(flat_grad.shape)} gradients")
        self.param, self.grad = flat_param, flat_grad
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat_param), torch.zeros_like(flat_param)
        self.lr, self.betas, self.eps, self.steps = float(lr), (float(betas[0]), float(betas[1])), float(eps), 0

    def step(self, grad_scale=1.0):
        self.steps += 1
        check(_capi.load().dram_adam_step(_p(self.param), _p(self.grad), _p(self.exp_avg), _p(self.exp_avg_sq),
                                          self.param.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.steps,
                                          grad_scale, _stream()), "dram_adam_step")

    def decay_lr(self, gamma=0.95):
        self.lr *= gamma

    def state_dict(self):
        return {"step": self.steps, "lr": self.lr, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}

    def load_state_dict(self, state):
        self.steps, self.lr = int(state["step"]), float(state["lr"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])


class PeerExchange:
    """K10x: the SyncBatchNorm statistics exchange (train.py:101) as one kernel over NVLink peer memory instead of a
    library collective (`dram_peer_allreduce_f64`, peer_exchange.cu).  One process per GPU on one node, at most 8 ranks.
    Construction is collective over `group`: every rank allocates its exchange buffer, the CUDA IPC handles travel
    through `all_gather_object`, every rank maps every peer's buffer, then a barrier.  `all_reduce(sums)` must be called
    in the same order on every rank (it is: the BatchNorm layers run in the same order everywhere)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        lib = _capi.load()
        self._lib, self.group = lib, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError(f"PeerExchange: {self.world} ranks; the peer-memory exchange covers one node (<= 8 GPUs)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        own, handle = C.c_void_p(), (C.c_ubyte * _capi.PEER_HANDLE_BYTES)()
        check(lib.dram_peer_alloc(lib.dram_peer_exchange_bytes(), C.byref(own), handle), "dram_peer_alloc")
        self._own, self._opened = own.value, []
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._ptrs = (C.c_void_p * 8)()
        for r, h in enumerate(handles):
            if r == self.rank:
                self._ptrs[r] = self._own
                continue
            ptr, buf = C.c_void_p(), (C.c_ubyte * _capi.PEER_HANDLE_BYTES).from_buffer_copy(h)
            check(lib.dram_peer_open(buf, C.byref(ptr)), f"dram_peer_open (rank {r})")
            self._ptrs[r] = ptr.value
            self._opened.append(ptr.value)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.seq = 0
        dist.barrier(group)  # every buffer is zeroed and mapped everywhere before the first flag is raised

    def all_reduce(self, sums, out=None):
        """fp64 [n <= 4096] -> the sum over the ranks, the same bits on every rank (`out` may be `sums`)."""
        _need(sums, torch.float64, "PeerExchange sums", 1)
        out = torch.empty_like(sums) if out is None else _need(out, torch.float64, "PeerExchange out", 1)
        self.seq += 1
        check(self._lib.dram_peer_allreduce_f64(_p(sums), _p(out), sums.numel(), self._ptrs, self.world, self.rank, self.seq,
                                                _p(self.status), _stream()), "dram_peer_allreduce_f64")
        return out

    def check(self):
        """Raises if an exchange gave up waiting for a peer (synchronises the device; call between steps)."""
        code = int(self.status.item())
        if code:
            raise RuntimeError(f"PeerExchange: rank {self.rank} waited > 60 s for rank {code - 1} at some exchange <= {self.seq}")

    def close(self):
        """Collective: every rank unmaps its peers' buffers, then (after a barrier — CUDA IPC wants the importers gone
        before the exporter frees) releases its own."""
        import torch.distributed as dist

        torch.cuda.synchronize(self.device)
        for ptr in self._opened:
            self._lib.dram_peer_close(C.c_void_p(ptr))
        self._opened = []
        dist.barrier(self.group)
        if self._own:
            self._lib.dram_peer_free(C.c_void_p(self._own))
            self._own = None


_BN_WS = {}


def _bn_workspace(c, device):
    key = (device.index, c)
    ws = _BN_WS.get(key)
    if ws is None:
        nbytes = _capi.load().dram_bn_workspace_bytes(c)
        ws = _BN_WS[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
    return ws


class BatchNormTrainFn(torch.autograd.Function):
    """y = act(batch_norm_train(x) (+ res)) on NDHWC 16-bit rows (K10, `dram_bn_*`): batch statistics, running-stat
    update (in place on the given buffers), ReLU and residual add fused; backward returns dx, dgamma, dbeta, dres.
    `group` (a torch.distributed group, True for the default group, or a `PeerExchange`) all-reduces the per-channel
    sums between the two phases of both reductions: SyncBatchNorm as train.py:101 asks for."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, res, relu, eps, momentum, group, sink=None):
        """sink: None, or (fp32 view [2c] of a flat gradient buffer laid out [dbeta | dgamma], callback): backward then
        writes both parameter gradients there with one copy and calls the callback instead of returning them."""
        lib = _capi.load()
        _need16(x, "bn_train x")
        c = x.shape[-1]
        m = x.numel() // c
        dt = ACT_DTYPES[x.dtype]
        if res is not None:
            _need(res, x.dtype, "bn_train res")
            if res.shape != x.shape:
                raise ValueError(f"bn_train: residual {tuple(res.shape)} vs x {tuple(x.shape)}")
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        ws = _bn_workspace(c, x.device)
        sums = torch.empty(2 * c, dtype=torch.float64, device=x.device)
        check(lib.dram_bn_stats(_p(x), m, c, dt, _p(sums), _p(ws), _stream()), "dram_bn_stats")
        count = float(m)
        world = _sync_world(group)
        if world > 1:
            _sync_all_reduce(sums, group)
            count *= world
        scale, shift, mean, rstd = (torch.empty(c, dtype=torch.float32, device=x.device) for _ in range(4))
        check(lib.dram_bn_finalize(_p(sums), count, _p(g32), _p(b32), eps, momentum, _p(running_mean), _p(running_var),
                                   _p(scale), _p(shift), _p(mean), _p(rstd), c, _stream()), "dram_bn_finalize")
        out = torch.empty_like(x)
        check(lib.dram_bn_apply(_p(x), _p(scale), _p(shift), _p(res), 1 if relu else 0, _p(out), m, c, dt, _stream()),
              "dram_bn_apply")
        ctx.save_for_backward(x, out if relu else None, mean, rstd, g32)
        ctx.meta = (m, c, dt, group, res is not None, sink)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _capi.load()
        x, y, mean, rstd, g32 = ctx.saved_tensors
        m, c, dt, group, has_res, sink = ctx.meta
        dy = dy.contiguous()
        ws = _bn_workspace(c, x.device)
        sums = torch.empty(2 * c, dtype=torch.float64, device=x.device)
        check(lib.dram_bn_backward_reduce(_p(dy), _p(x), _p(y), _p(mean), _p(rstd), m, c, dt, _p(sums), _p(ws), _stream()),
              "dram_bn_backward_reduce")
        # local sums (the gradient exchange averages them later): {sum dz, sum dz*xhat} = {dbeta, dgamma}
        if sink is not None:
            sink[0].copy_(sums)
            dbeta = dgamma = None
        else:
            local = sums.float()
            dbeta, dgamma = local[:c], local[c:]
        count = float(m)
        world = _sync_world(group)
        if world > 1:
            sums = _sync_all_reduce(sums, group, in_place=False)
            count *= world
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res and ctx.needs_input_grad[5] else None
        check(lib.dram_bn_backward_apply(_p(dy), _p(x), _p(y), _p(mean), _p(rstd), _p(g32), _p(sums), count, _p(dx), _p(dres),
                                         m, c, dt, _stream()), "dram_bn_backward_apply")
        if sink is not None:
            sink[1]()
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None, None


def channel_sums(x):
    """Per-channel sum over all rows of an NDHWC 16-bit tensor, fp32 [C] (phase 1+2 of K10's statistics): the bias
    gradient of a convolution is this of its dy."""
    lib = _capi.load()
    _need16(x, "channel_sums x")
    c = x.shape[-1]
    sums = torch.empty(2 * c, dtype=torch.float64, device=x.device)
    check(lib.dram_bn_stats(_p(x), x.numel() // c, c, ACT_DTYPES[x.dtype], _p(sums), _p(_bn_workspace(c, x.device)), _stream()),
          "dram_bn_stats")
    return sums[:c].float()


def _sync_all_reduce(sums, group, in_place=True):
    """The statistics exchange of SyncBatchNorm: K10x over peer memory when `group` is a PeerExchange, else NCCL."""
    if isinstance(group, PeerExchange):
        return group.all_reduce(sums, out=sums if in_place else None)
    import torch.distributed as dist

    if not in_place:
        sums = sums.clone()
    dist.all_reduce(sums, group=None if group is True else group)
    return sums


def _sync_world(group):
    if group is None or group is False:
        return 1
    if isinstance(group, PeerExchange):
        return group.world
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(None if group is True else group)


class GradBuckets:
    """Flat fp32 gradient storage, split into buckets that are all-reduced independently.

    `named_shapes` = ordered [(name, shape)] in the order gradients become ready (reverse forward order);
    a bucket closes once it holds `bucket_bytes`.  `view(name)` is the contiguous fp32 tensor a wgrad plan
    writes into; `reduce_bucket(i)` launches the (asynchronous) average over the process group as soon as the
    bucket's last gradient has been produced, so the exchange overlaps the rest of the backward pass.
    """

    def __init__(self, named_shapes, device, bucket_bytes=64 << 20):
        self.slices, self.bucket_of, bounds = {}, {}, [0]
        off = 0
        for name, shape in named_shapes:
            numel = 1
            for v in shape:
                numel *= int(v)
            self.slices[name] = (off, numel, tuple(shape))
            self.bucket_of[name] = len(bounds) - 1
            off += numel
            if (off - bounds[-1]) * 4 >= bucket_bytes:
                bounds.append(off)
        if bounds[-1] != off:
            bounds.append(off)
        self.bounds = bounds
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self._work = [None] * self.num_buckets

    @property
    def num_buckets(self):
        return len(self.bounds) - 1

    def view(self, name):
        off, numel, shape = self.slices[name]
        return self.flat[off:off + numel].view(shape)

    def bucket(self, i):
        return self.flat[self.bounds[i]:self.bounds[i + 1]]

    def reduce_bucket(self, i, group=None):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        b = self.bucket(i)
        b.div_(dist.get_world_size(group))
        self._work[i] = dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group, async_op=True)
        return self._work[i]

    def wait(self):
        for i, w in enumerate(self._work):
            if w is not None:
                w.wait()
                self._work[i] = None
