"""Drop-in `med3d` module: the reference's six factories and two network classes, with the
forward pass executed by hand-written sm_100a kernels (libdram_b200.so) instead of ATen/cuDNN.

Interface kept from /root/reference/med3d.py:
  * factories `resnet{18,34,50}seg{cls,reg}(**kwargs)` (med3d.py:391-425), kwargs `shortcut_type`
    (default 'A') and, for the cls nets, `n_classes` (default [6, 3]);
  * classes `ResNetSegCls` / `ResNetSegReg` with `.forward(x, lungs=None) -> (dense_outs, outs)`
    (med3d.py:270-285, 369-388) and `.get_target_layer()` (-> `us3`);
  * `state_dict()` keys, shapes, dtypes and order identical to the reference,
    so any reference checkpoint loads — the sub-modules below are ordinary nn.Conv3d /
    nn.BatchNorm3d objects used purely as parameter containers;
  * initialisation as med3d.py:334-339 (kaiming-normal fan_out convolutions, BN weight 1 / bias 0).

What differs: `forward()` is eval-mode inference on a CUDA (B200) device.  Calling it in training mode, on CPU
tensors, or with an input size the up-sampling path cannot match raises — there is deliberately no PyTorch
fallback.  The training step on the same parameters lives in `training.py` (`TrainableMed3D`, `TrainStep`).
"""
import torch
import torch.nn as nn

from .engine import LAYER_CFG, Med3DEngine


class _Block(nn.Module):
    """Parameter container for one residual block (BasicBlock med3d.py:115-127 or Bottleneck 147-162)."""

    def __init__(self, kind, inplanes, planes, stride, dilation, downsample, shortcut_type):
        super().__init__()
        if kind == "basic":
            self.conv1 = nn.Conv3d(inplanes, planes, 3, stride=stride, dilation=dilation, padding=dilation, bias=False)
            self.bn1 = nn.BatchNorm3d(planes)
            self.relu = nn.ReLU(inplace=True)
            self.conv2 = nn.Conv3d(planes, planes, 3, dilation=dilation, padding=dilation, bias=False)
            self.bn2 = nn.BatchNorm3d(planes)
        else:
            self.conv1 = nn.Conv3d(inplanes, planes, 1, bias=False)
            self.bn1 = nn.BatchNorm3d(planes)
            self.conv2 = nn.Conv3d(planes, planes, 3, stride=stride, dilation=dilation, padding=dilation, bias=False)
            self.bn2 = nn.BatchNorm3d(planes)
            self.conv3 = nn.Conv3d(planes, planes * 4, 1, bias=False)
            self.bn3 = nn.BatchNorm3d(planes * 4)
            self.relu = nn.ReLU(inplace=True)
        # type A has no parameters (the reference stores a functools.partial); a truthy marker keeps
        # `block.downsample is not None` meaningful.  Type B owns a 1x1x1 conv + BN.
        if downsample and shortcut_type != "A":
            out = planes * (4 if kind == "bottleneck" else 1)
            self.downsample = nn.Sequential(nn.Conv3d(inplanes, out, 1, stride=stride, bias=False), nn.BatchNorm3d(out))
        else:
            self.downsample = "A" if downsample else None
        self.stride, self.dilation, self.shortcut_type = stride, dilation, shortcut_type

    def forward(self, x):
        raise RuntimeError("blocks are parameter containers; run the whole network through its forward()")


class BasicBlock(_Block):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, shortcut_type="A"):
        super().__init__("basic", inplanes, planes, stride, dilation, downsample, shortcut_type)


class Bottleneck(_Block):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, shortcut_type="A"):
        super().__init__("bottleneck", inplanes, planes, stride, dilation, downsample, shortcut_type)


class UpsampleConvBlock5d(nn.Module):
    """Parameter container of a decoder stage: `conv_blocks.{0,1}.{0: conv3^3 bias=True, 1: BN, 2: ReLU}`
    (med3d.py:50-89).  The x2 trilinear up-sampling and the skip concat are kernels K4 / K1."""

    def __init__(self, in_chs, base_chs):
        super().__init__()
        self.conv_blocks = nn.Sequential(*[
            nn.Sequential(nn.Conv3d(i, o, kernel_size=3, padding=1, bias=True), nn.BatchNorm3d(o), nn.ReLU(inplace=True))
            for i, o in zip(in_chs, base_chs)])

    def forward(self, inputs, cats, args=None):
        raise RuntimeError("decoder stages are parameter containers; run the whole network through its forward()")


class _Med3DSegNet(nn.Module):
    head_kind = None  # 'cls' | 'reg'

    def __init__(self, block, layers, shortcut_type="A", head_channels=(1, 1)):
        super().__init__()
        self.block_kind = "bottleneck" if block is Bottleneck else "basic"
        self.expansion = block.expansion
        self.shortcut_type = shortcut_type
        self.inplanes = 64
        self.conv1 = nn.Conv3d(1, 64, kernel_size=7, stride=(2, 2, 2), padding=(3, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(kernel_size=(3, 3, 3), stride=2, padding=1)
        stages = []
        for (planes, stride, dilation), n in zip(LAYER_CFG, layers):
            blocks = []
            for i in range(n):
                first = i == 0
                ds = first and (stride != 1 or self.inplanes != planes * block.expansion)
                blocks.append(block(self.inplanes, planes, stride=stride if first else 1, dilation=dilation,
                                    downsample=ds, shortcut_type=shortcut_type))
                self.inplanes = planes * block.expansion
            stages.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        self.us1 = UpsampleConvBlock5d([(512 + 64) * block.expansion, 64], [64, 64])
        self.us2 = UpsampleConvBlock5d([64 + 64, 64], [64, 64])
        self.us3 = nn.Sequential(nn.Conv3d(64, 32, kernel_size=3, padding=1), nn.BatchNorm3d(32), nn.ReLU(inplace=True))
        self.fcs = nn.ModuleList([nn.Conv3d(32, c, kernel_size=1, padding=0, stride=1, bias=True) for c in head_channels])
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self._engines = {}
        # 16-bit storage type of activations/weights: None -> ops.default_act_dtype() (env DRAM_B200_DTYPE)
        self.act_dtype = None
        # Bumped whenever the parameters may have changed; engines re-fold BatchNorm and re-pack their operands when
        # it differs from the value they packed at.  Covered automatically: load_state_dict (post hook), .to()/.half()
        # (_apply rebuilds the engines), train() <-> eval() transitions (a training step edits parameters in place).
        # NOT seen: in-place edits of parameters while the module stays in eval mode (`p.data.mul_(...)`,
        # `p.copy_(...)`): call mark_weights_changed() after those.
        self.weights_epoch = 0
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.mark_weights_changed())

    def mark_weights_changed(self):
        """Tell the cached engines that parameters/buffers were edited in place (see weights_epoch)."""
        self.weights_epoch += 1

    def train(self, mode=True):
        if mode != self.training:
            self.weights_epoch += 1
        return super().train(mode)

    def get_target_layer(self):
        return self.us3

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engines"] = {}  # plans hold device pointers and C handles: never pickled or deep-copied
        return state

    # ------------------------------------------------------------------ engine cache
    def engine(self, batch, dims, device):
        """The static plan for this (batch, D, H, W) on `device` (built on first use)."""
        key = (batch, tuple(dims), device.index if device.index is not None else torch.cuda.current_device(),
               self.act_dtype)
        eng = self._engines.get(key)
        if eng is None:
            eng = Med3DEngine(self, batch, dims, device, self.act_dtype)
            self._engines[key] = eng
        return eng

    def _apply(self, fn, *args, **kwargs):
        # .to()/.cuda()/.half() move or retype the parameters the plans were packed from
        self._engines = {}
        return super()._apply(fn, *args, **kwargs)

    def forward(self, x, lungs=None):
        if self.training:
            raise RuntimeError("this module's forward() is the eval-mode inference path (call .eval()); a training step "
                               "on the same parameters runs through dram_b200.training.TrainableMed3D / TrainStep")
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise RuntimeError("dram_b200.med3d needs CUDA tensors on a B200; there is no CPU fallback")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape [B, 1, D, H, W], got {tuple(x.shape)}")
        if next(self.parameters()).device != x.device:
            raise RuntimeError("model parameters and input are on different devices")
        B = x.shape[0]
        eng = self.engine(B, tuple(x.shape[2:]), x.device)
        with torch.no_grad():
            dense, pooled = eng.run(x.float().contiguous(), lungs)
            # the engine owns its output buffers; hand out copies so a later forward cannot alias them
            return [d.clone() for d in dense], [p.clone() for p in pooled]


class ResNetSegCls(_Med3DSegNet):
    """med3d.py:187-285: raw class maps (6 CLE, 3 PSE channels) + global average pooled logits."""
    head_kind = "cls"

    def __init__(self, block, layers, shortcut_type="A", n_classes=[6, 3]):
        self.n_classes = list(n_classes)
        super().__init__(block, layers, shortcut_type, head_channels=tuple(self.n_classes))


class ResNetSegReg(_Med3DSegNet):
    """med3d.py:288-388: sigmoid maps (1 CLE, 1 PSE channel) + lobe-masked means."""
    head_kind = "reg"

    def __init__(self, block, layers, shortcut_type="A"):
        super().__init__(block, layers, shortcut_type, head_channels=(1, 1))


def resnet18segcls(**kwargs):
    return ResNetSegCls(BasicBlock, [2, 2, 2, 2], **kwargs)


def resnet34segcls(**kwargs):
    return ResNetSegCls(BasicBlock, [3, 4, 6, 3], **kwargs)


def resnet50segcls(**kwargs):
    return ResNetSegCls(Bottleneck, [3, 4, 6, 3], **kwargs)


def resnet18segreg(**kwargs):
    return ResNetSegReg(BasicBlock, [2, 2, 2, 2], **kwargs)


def resnet34segreg(**kwargs):
    return ResNetSegReg(BasicBlock, [3, 4, 6, 3], **kwargs)


def resnet50segreg(**kwargs):
    return ResNetSegReg(Bottleneck, [3, 4, 6, 3], **kwargs)
