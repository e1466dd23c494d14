"""TEST INFRASTRUCTURE — CPU oracle of the convolution backward pass.

The reference has no backward code of its own: `train.py:88-106` builds a Lightning `Trainer` whose automatic
optimisation calls `loss.backward()` on the losses of `models.py:495-582`, and autograd differentiates each
`nn.Conv3d` that med3d.py constructs (`conv3x3x3` med3d.py:93-100, decoder convs med3d.py:67/76, bottleneck
convs med3d.py:152-157, stem med3d.py:296-304).  The arithmetic therefore lives in PyTorch (pinned by the
reference at 1.12.0, 2.11 here); this oracle runs exactly that — `F.conv3d` on CPU in fp32 and
`torch.autograd.grad` — on the operands the kernels read.  Parity is pinned to the reference only through these
call sites (the reference ships no gradient fixtures): "parity unpinned" for this row.
"""
import torch
import torch.nn.functional as F


def conv3d_grads(x, weight, dy, stride=1, dilation=1, padding=None):
    """x [N,Cin,D,H,W], weight [Cout,Cin,kd,kh,kw], dy [N,Cout,Do,Ho,Wo] (all fp32, CPU) ->
    (dx, dw) as autograd produces them for `y = F.conv3d(x, weight, None, stride, padding, dilation)`."""
    k = weight.shape[2:]
    dl = (dilation,) * 3 if isinstance(dilation, int) else tuple(dilation)
    if padding is None:
        padding = tuple(dl[i] * (k[i] - 1) // 2 for i in range(3))
    x = x.detach().clone().requires_grad_(True)
    w = weight.detach().clone().requires_grad_(True)
    y = F.conv3d(x, w, None, stride=stride, padding=padding, dilation=dilation)
    assert y.shape == dy.shape, (y.shape, dy.shape)
    dx, dw = torch.autograd.grad(y, (x, w), dy)
    return dx, dw
