"""DRN-C-26 feature extractor (PyTorch / cuDNN) -- the INPUT producer of the hot path, not
part of the product.  Architecture after Yu et al., "Dilated Residual Networks" as used by
the reference (models/drn.py:230-285): 7x7 stem, eight stages, output stride 8, 512 channels.
Weights are random-init (no checkpoint can be fetched here); stage 8 (``layer8``) is the map
the reference clusters (``--use_feature_maps 7``)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _conv_bn(cin, cout, k, stride=1, dilation=1):
    pad = dilation * (k // 2)
    return [nn.Conv2d(cin, cout, k, stride, pad, dilation=dilation, bias=False),
            nn.BatchNorm2d(cout)]


class _Block(nn.Module):
    """Two 3x3 convs; optional identity / projection shortcut."""

    def __init__(self, cin, cout, stride, dil1, dil2, residual):
        super().__init__()
        self.body = nn.Sequential(*_conv_bn(cin, cout, 3, stride, dil1), nn.ReLU(inplace=True),
                                  *_conv_bn(cout, cout, 3, 1, dil2))
        self.residual = residual
        self.proj = None
        if residual and (stride != 1 or cin != cout):
            self.proj = nn.Sequential(*_conv_bn(cin, cout, 1, stride))

    def forward(self, x):
        y = self.body(x)
        if self.residual:
            y = y + (x if self.proj is None else self.proj(x))
        return torch.relu(y)


class DRNC26(nn.Module):
    #            channels, blocks, stride, dilation, new_level, residual
    STAGES = [(16, 1, 1, 1, True, True), (32, 1, 2, 1, True, True), (64, 2, 2, 1, True, True),
              (128, 2, 2, 1, True, True), (256, 2, 1, 2, False, True),
              (512, 2, 1, 4, False, True), (512, 1, 1, 2, False, False),
              (512, 1, 1, 1, False, False)]

    def __init__(self):
        super().__init__()
        self.stem = nn.Sequential(*_conv_bn(3, 16, 7), nn.ReLU(inplace=True))
        stages, cin = [], 16
        for (c, n, stride, dil, new_level, residual) in self.STAGES:
            blocks = []
            for i in range(n):
                if i == 0:
                    d1 = 1 if dil == 1 else (dil // 2 if new_level else dil)
                    blocks.append(_Block(cin, c, stride, d1, dil, residual))
                else:
                    blocks.append(_Block(c, c, 1, dil, dil, residual))
                cin = c
            stages.append(nn.Sequential(*blocks))
        self.stages = nn.ModuleList(stages)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan))

    def forward(self, x, out_middle=False):
        x = self.stem(x)
        maps = []
        for st in self.stages:
            x = st(x)
            maps.append(x)
        return (x, maps) if out_middle else x


def fold_batchnorm(model):
    """Inference-time folding of every Conv2d -> BatchNorm2d pair into one convolution with bias
    (same function in exact arithmetic; removes one full pass over each activation map)."""
    def fold_seq(seq):
        mods = list(seq.children())
        out, i = [], 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d):
                bn = mods[i + 1]
                scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                conv = nn.Conv2d(m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding,
                                 dilation=m.dilation, bias=True)
                conv.weight.data = m.weight.data * scale.view(-1, 1, 1, 1)
                b0 = m.bias.data if m.bias is not None else torch.zeros_like(bn.running_mean)
                conv.bias.data = (b0 - bn.running_mean) * scale + bn.bias
                out.append(conv)
                i += 2
            else:
                out.append(m)
                i += 1
        return nn.Sequential(*out)

    def walk(mod):
        for name, child in list(mod.named_children()):
            if isinstance(child, nn.Sequential) and any(isinstance(c, nn.BatchNorm2d) for c in child.children()):
                setattr(mod, name, fold_seq(child))
            else:
                walk(child)
    with torch.no_grad():
        walk(model)
    return model


class FusedDRN(nn.Module):
    """Inference form of a BatchNorm-folded DRNC26 for CUDA: every conv + bias + ReLU is one cuDNN
    call (``torch.cudnn_convolution_relu``) and the tail of a residual block -- conv + bias +
    shortcut + ReLU -- is one ``torch.cudnn_convolution_add_relu``.  Same function as the module it
    wraps (up to the convolution algorithm's rounding); removes the separate element-wise passes
    over the full-resolution activation maps."""

    def __init__(self, folded: DRNC26):
        super().__init__()
        self.m = folded

    @staticmethod
    def _cr(x, conv):
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding,
                                            conv.dilation, conv.groups)

    @staticmethod
    def _car(x, conv, z):
        return torch.cudnn_convolution_add_relu(x, conv.weight, z, 1.0, conv.bias, conv.stride,
                                                conv.padding, conv.dilation, conv.groups)

    def forward(self, x, out_middle=False):
        x = self._cr(x, self.m.stem[0])
        maps = []
        for st in self.m.stages:
            for blk in st:
                c1, c2 = blk.body[0], blk.body[2]
                y = self._cr(x, c1)
                if blk.residual:
                    z = x if blk.proj is None else torch.nn.functional.conv2d(
                        x, blk.proj[0].weight, blk.proj[0].bias, blk.proj[0].stride,
                        blk.proj[0].padding, blk.proj[0].dilation)
                    x = self._car(y, c2, z)
                else:
                    x = self._cr(y, c2)
            maps.append(x)
        return (x, maps) if out_middle else x


def drn_c_26(seed=1111, device='cuda', channels_last=True, fold_bn=False, fused=False):
    torch.manual_seed(seed)
    m = DRNC26().eval()
    if fold_bn:
        m = fold_batchnorm(m)
    m = m.to(device)
    if channels_last:
        m = m.to(memory_format=torch.channels_last)
    for p in m.parameters():
        p.requires_grad_(False)
    if fused:
        assert fold_bn, 'the fused inference form needs the BatchNorms folded'
        m = FusedDRN(m)
    return m
