"""Drop-in for the reference's native plugin module ``alt_cuda_corr``.

    import alt_cuda_corr                       # reference: pybind11 module, correlation.cpp:50-53
    (corr,) = alt_cuda_corr.forward(fmap1, fmap2, coords, radius)

becomes ``from ptlflow_b200 import alt_cuda_corr`` with the identical call.  Tensor contract as
in ptlflow/utils/external/alt_cuda_corr/correlation_kernel.cu:260-286: fmap1 [B,H1,W1,C],
fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2] (N == 1 at every call site), result [B,N,(2r+1)^2,H1,W1],
unscaled; inputs must be CUDA + contiguous else RuntimeError (CHECK_INPUT, correlation.cpp:19-21).
Unlike the reference (fp32 only) half/bfloat16 features are accepted directly.
"""
from __future__ import annotations

import torch

from ._lib import check, dtype_code, load, require_cuda, stream_ptr


def forward(fmap1: torch.Tensor, fmap2: torch.Tensor, coords: torch.Tensor, radius: int):
    for name, t in (("fmap1", fmap1), ("fmap2", fmap2), ("coords", coords)):
        require_cuda(t, name)
    if fmap1.dim() != 4 or fmap2.dim() != 4 or coords.dim() != 5 or coords.shape[-1] != 2:
        raise RuntimeError("alt_cuda_corr.forward: expected fmap [B,H,W,C] and coords [B,N,H1,W1,2]")
    if coords.shape[1] != 1:
        raise RuntimeError("alt_cuda_corr.forward: N != 1 is not used by any caller and not supported")
    B, H1, W1, C = fmap1.shape
    H2, W2 = fmap2.shape[1:3]
    c32 = coords if coords.dtype == torch.float32 else coords.float()
    rd = 2 * radius + 1
    out = torch.empty((B, 1, rd * rd, H1, W1), dtype=fmap1.dtype, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        check(load().pfb_alt_corr_forward(fmap1.data_ptr(), fmap2.data_ptr(), c32.data_ptr(), out.data_ptr(), B, H1, W1, H2, W2, C,
                                          radius, dtype_code(fmap1.dtype), dtype_code(out.dtype), stream_ptr(fmap1.device)),
              "alt_cuda_corr.forward")
    return [out]


def backward(fmap1, fmap2, coords, corr_grad, radius):
    raise NotImplementedError("alt_cuda_corr.backward is training-only and outside the inference hot path (SURVEY.md 2.2)")
