"""Feature / context encoders (not on the named hot path, kept as
cuDNN modules in channels_last; parameter names and shapes equal the reference's
ptlflow/models/raft/extractor.py:122-267 so checkpoints load strictly)."""
from __future__ import annotations

import torch
import torch.nn as nn


def _norm(kind: str, channels: int, groups: int = 8) -> nn.Module:
    if kind == "batch":
        return nn.BatchNorm2d(channels)
    if kind == "instance":
        return nn.InstanceNorm2d(channels)
    if kind == "group":
        return nn.GroupNorm(num_groups=groups, num_channels=channels)
    if kind == "none":
        return nn.Identity()
    raise ValueError(f"unknown norm_fn {kind!r}")


class ResidualBlock(nn.Module):
    """3x3 -> 3x3 with identity (or strided 1x1) shortcut."""

    def __init__(self, in_planes: int, planes: int, norm_fn: str = "group", stride: int = 1) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride=stride, padding=1)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        g = planes // 8
        self.norm1, self.norm2 = _norm(norm_fn, planes, g), _norm(norm_fn, planes, g)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(norm_fn, planes, g)
            # the norm is shared: state_dict carries it as both norm3.* and downsample.1.* (as upstream)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class BottleneckBlock(nn.Module):
    """1x1 -> 3x3 -> 1x1 bottleneck used by the small encoder."""

    def __init__(self, in_planes: int, planes: int, norm_fn: str = "group", stride: int = 1) -> None:
        super().__init__()
        q = planes // 4
        self.conv1 = nn.Conv2d(in_planes, q, 1)
        self.conv2 = nn.Conv2d(q, q, 3, stride=stride, padding=1)
        self.conv3 = nn.Conv2d(q, planes, 1)
        self.relu = nn.ReLU(inplace=True)
        g = planes // 8
        self.norm1, self.norm2, self.norm3 = _norm(norm_fn, q, g), _norm(norm_fn, q, g), _norm(norm_fn, planes, g)
        self.downsample = None
        if stride != 1:
            self.norm4 = _norm(norm_fn, planes, g)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride=stride), self.norm4)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        y = self.relu(self.norm3(self.conv3(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class _Encoder(nn.Module):
    block = ResidualBlock
    widths = (64, 64, 96, 128)

    def __init__(self, output_dim: int = 128, norm_fn: str = "batch", dropout: float = 0.0) -> None:
        super().__init__()
        self.norm_fn = norm_fn
        w0, w1, w2, w3 = self.widths
        self.norm1 = _norm(norm_fn, w0, 8)
        self.conv1 = nn.Conv2d(3, w0, 7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        self.layer1 = self._stage(w0, w1, 1)
        self.layer2 = self._stage(w1, w2, 2)
        self.layer3 = self._stage(w2, w3, 2)
        self.conv2 = nn.Conv2d(w3, output_dim, 1)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)) and m.weight is not None:
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _stage(self, cin: int, cout: int, stride: int) -> nn.Sequential:
        return nn.Sequential(self.block(cin, cout, self.norm_fn, stride=stride), self.block(cout, cout, self.norm_fn, stride=1))

    def forward(self, x):
        """Accepts one tensor or a list/tuple of two (processed as one batch: instance norm is
        per-sample, so this is exact -- extractor.py:173-176)."""
        pair = isinstance(x, (tuple, list))
        if pair:
            n = x[0].shape[0]
            x = torch.cat(list(x), dim=0)
        x = self.relu1(self.norm1(self.conv1(x)))
        x = self.layer3(self.layer2(self.layer1(x)))
        x = self.conv2(x)
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return (x[:n], x[n:]) if pair else x


class BasicEncoder(_Encoder):
    block = ResidualBlock
    widths = (64, 64, 96, 128)


class SmallEncoder(_Encoder):
    block = BottleneckBlock
    widths = (32, 32, 64, 96)
