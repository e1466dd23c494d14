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


def bn_train(x, gamma, beta, res, relu, dy, eps=1e-5, momentum=0.1):
    """Train-mode nn.BatchNorm3d (+ residual add, + ReLU) as the blocks of med3d.py:129-144 / 164-184 compose it, on
    CPU fp32 with autograd.  x, res, dy: [N, C, D, H, W].  Returns y, running stats after one update from (0, 1), and
    the gradients (dx, dgamma, dbeta, dres)."""
    x = x.detach().clone().requires_grad_(True)
    g = gamma.detach().clone().requires_grad_(True)
    b = beta.detach().clone().requires_grad_(True)
    r = None if res is None else res.detach().clone().requires_grad_(True)
    rm, rv = torch.zeros_like(gamma), torch.ones_like(gamma)
    y = F.batch_norm(x, rm, rv, g, b, True, momentum, eps)
    if r is not None:
        y = y + r
    if relu:
        y = torch.relu(y)
    ins = (x, g, b) + (() if r is None else (r,))
    grads = torch.autograd.grad(y, ins, dy)
    return y.detach(), rm, rv, grads


def upsample2x_grad(x, dy):
    """nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True) (med3d.py:83): output and input gradient."""
    x = x.detach().clone().requires_grad_(True)
    y = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    (dx,) = torch.autograd.grad(y, x, dy)
    return y.detach(), dx


def maxpool3d_grad(x, dy):
    """nn.MaxPool3d(kernel_size=3, stride=2, padding=1) (med3d.py:305): output and input gradient."""
    x = x.detach().clone().requires_grad_(True)
    y = F.max_pool3d(x, kernel_size=3, stride=2, padding=1)
    (dx,) = torch.autograd.grad(y, x, dy)
    return y.detach(), dx
