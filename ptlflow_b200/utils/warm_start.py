"""Warm start: forward-warp the previous low-resolution flow (RAFT's ``forward_interpolate``).

The reference does this on the CPU with scipy ``griddata(method="nearest")``
(ptlflow/utils/external/raft.py:155-185, wrapper ptlflow/utils/utils.py:454-478), on the critical path of
``infer.py`` / ``validate.py`` when ``warm_start=True``.  CUDA tensors go through
``pfb_forward_interpolate`` (exact nearest neighbour in fp64 on the device, no host round trip); the scipy form
below is the restatement of the reference for host tensors and the checker of the device kernel in the tests.
"""
from __future__ import annotations

import numpy as np
import torch


def _forward_interpolate(flow: torch.Tensor) -> torch.Tensor:
    from scipy import interpolate

    f = flow.detach().float().cpu().numpy()
    dx, dy = f[0], f[1]
    ht, wd = dx.shape
    x0, y0 = np.meshgrid(np.arange(wd), np.arange(ht))
    x1, y1 = (x0 + dx).reshape(-1), (y0 + dy).reshape(-1)
    dxf, dyf = dx.reshape(-1), dy.reshape(-1)
    keep = (x1 > 0) & (x1 < wd) & (y1 > 0) & (y1 < ht)
    if keep.sum() == 0:
        return torch.zeros_like(flow)
    pts = (x1[keep], y1[keep])
    fx = interpolate.griddata(pts, dxf[keep], (x0, y0), method="nearest", fill_value=0)
    fy = interpolate.griddata(pts, dyf[keep], (x0, y0), method="nearest", fill_value=0)
    return torch.from_numpy(np.stack([fx, fy], axis=0)).to(dtype=flow.dtype)


def forward_interpolate_batch(prev_flow: torch.Tensor) -> torch.Tensor:
    if prev_flow.is_cuda:
        from .. import ops

        return ops.forward_interpolate(prev_flow).to(dtype=prev_flow.dtype)
    out = torch.stack([_forward_interpolate(prev_flow[i]) for i in range(prev_flow.shape[0])], dim=0)
    return out.to(dtype=prev_flow.dtype, device=prev_flow.device)
