"""Corr-block protocol (the intra-model seam of the reference, ptlflow/models/raft/corr.py:104-118):

    corr_fn = get_corr_block(fmap1, fmap2, num_levels=4, radius=4, alternate_corr=False)
    corr = corr_fn(coords)        # coords [B,2,H,W] -> [B, L*(2r+1)^2, H, W], contiguous, coords' dtype

Same names, argument meaning and output layout; the arithmetic is libptlflow_b200's
(pfb_corr_volume_build / pfb_corr_lookup / pfb_corr_lookup_onthefly).  Any RAFT-family model that
vendors its own copy of this file can bind to these classes unchanged.
"""
from __future__ import annotations

import torch

from ... import ops


class CorrBlock:
    """All-pairs volume + pooled pyramid, built once; ``__call__`` = radius-r multi-scale lookup."""

    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, radius: int = 4, impl: int = 0):
        self.num_levels, self.radius = num_levels, radius
        f1, f2 = ops.to_pixel_major(fmap1), ops.to_pixel_major(fmap2)
        self.grid_hw = tuple(f1.shape[1:3])
        self.corr_pyramid = ops.corr_volume_build(f1, f2, num_levels, impl=impl)

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        c = ops.coords_to_pixel_major(coords)
        return ops.corr_lookup(self.corr_pyramid, c, self.radius, self.grid_hw, nchw=True, out_dtype=coords.dtype)


class AlternateCorrBlock:
    """On-the-fly variant: never materialises the 4D volume (corr.py:67-101)."""

    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, radius: int = 4):
        self.num_levels, self.radius = num_levels, radius
        self.fmap1 = ops.to_pixel_major(fmap1)
        self.pyramid = ops.feature_pyramid(ops.to_pixel_major(fmap2), num_levels)

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        c = ops.coords_to_pixel_major(coords)
        return ops.corr_lookup_onthefly(self.fmap1, self.pyramid, c, self.radius, nchw=True, out_dtype=coords.dtype)


def get_corr_block(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, radius: int = 4,
                   alternate_corr: bool = False):
    cls = AlternateCorrBlock if alternate_corr else CorrBlock
    return cls(fmap1=fmap1, fmap2=fmap2, num_levels=num_levels, radius=radius)
