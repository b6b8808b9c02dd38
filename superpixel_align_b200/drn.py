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


def drn_c_26(seed=1111, device='cuda', channels_last=True):
    torch.manual_seed(seed)
    m = DRNC26().eval().to(device)
    if channels_last:
        m = m.to(memory_format=torch.channels_last)
    for p in m.parameters():
        p.requires_grad_(False)
    return m
