"""Parameter containers for the update blocks.

These modules only *hold* the weights under the reference's names and shapes
(ptlflow/models/raft/update.py:6-153: ``encoder.convc1 ... gru.convz1 ... flow_head.conv2 ...
mask.0 / mask.2``) so that ``restore_model``'s strict ``load_state_dict`` and ``count_parameters``
work unchanged.  The arithmetic runs in libptlflow_b200 (ptlflow_b200/engine.py packs these
tensors once); calling ``forward`` here is an error by design -- there is no PyTorch fallback.
"""
from __future__ import annotations

import torch.nn as nn


def _no_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is a parameter container: the update block runs in libptlflow_b200 (see ptlflow_b200/engine.py)")


class FlowHead(nn.Module):
    def __init__(self, input_dim: int = 128, hidden_dim: int = 256) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)

    forward = _no_forward


class ConvGRU(nn.Module):
    def __init__(self, hidden_dim: int = 128, input_dim: int = 192 + 128) -> None:
        super().__init__()
        for name in ("convz", "convr", "convq"):
            setattr(self, name, nn.Conv2d(hidden_dim + input_dim, hidden_dim, 3, padding=1))

    forward = _no_forward


class SepConvGRU(nn.Module):
    def __init__(self, hidden_dim: int = 128, input_dim: int = 192 + 128) -> None:
        super().__init__()
        cin = hidden_dim + input_dim
        for sfx, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for name in ("convz", "convr", "convq"):
                setattr(self, name + sfx, nn.Conv2d(cin, hidden_dim, k, padding=p))

    forward = _no_forward


class SmallMotionEncoder(nn.Module):
    def __init__(self, corr_levels: int, corr_radius: int) -> None:
        super().__init__()
        planes = corr_levels * (2 * corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(planes, 96, 1)
        self.convf1 = nn.Conv2d(2, 64, 7, padding=3)
        self.convf2 = nn.Conv2d(64, 32, 3, padding=1)
        self.conv = nn.Conv2d(128, 80, 3, padding=1)

    forward = _no_forward


class BasicMotionEncoder(nn.Module):
    def __init__(self, corr_levels: int, corr_radius: int) -> None:
        super().__init__()
        planes = corr_levels * (2 * corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(planes, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    forward = _no_forward


class SmallUpdateBlock(nn.Module):
    def __init__(self, corr_levels: int, corr_radius: int, hidden_dim: int = 96) -> None:
        super().__init__()
        self.encoder = SmallMotionEncoder(corr_levels, corr_radius)
        self.gru = ConvGRU(hidden_dim=hidden_dim, input_dim=82 + 64)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=128)

    forward = _no_forward


class BasicUpdateBlock(nn.Module):
    def __init__(self, corr_levels: int, corr_radius: int, hidden_dim: int = 128, input_dim: int = 128) -> None:
        super().__init__()
        self.encoder = BasicMotionEncoder(corr_levels, corr_radius)
        self.gru = SepConvGRU(hidden_dim=hidden_dim, input_dim=128 + hidden_dim)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 64 * 9, 1))

    forward = _no_forward
