"""Input resizers used by BaseModel.preprocess_images / postprocess_predictions.

Same behaviour as ptlflow/utils/utils.py:34-213 (InputPadder / InputScaler) and
ptlflow/utils/external/raft.py:43-86, written against the needs of the RAFT path.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


class InputPadder:
    """Pad [..., H, W] so H and W become multiples of ``stride`` (or equal to ``size``)."""

    def __init__(self, dims: Sequence[int], stride: Optional[int], size: Optional[Tuple[int, int]] = None,
                 two_side_pad: bool = True, pad_mode: str = "replicate", pad_value: float = 0.0) -> None:
        h, w = int(dims[-2]), int(dims[-1])
        if size is None:
            th, tw = int(math.ceil(h / stride)) * stride, int(math.ceil(w / stride)) * stride
        else:
            th, tw = int(size[0]), int(size[1])
        ph, pw = th - h, tw - w
        self.tgt_size = (th, tw)
        self.pad_mode, self.pad_value = pad_mode, pad_value
        if two_side_pad:
            self._pad = [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]  # left, right, top, bottom
        else:
            self._pad = [pw // 2, pw - pw // 2, 0, ph]

    @property
    def pad_top_left(self) -> Tuple[int, int]:
        return self._pad[2], self._pad[0]

    def fill(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-3]
        y = x.reshape(-1, *x.shape[-3:]) if x.dim() > 4 else x
        kw = {"value": self.pad_value} if self.pad_mode == "constant" else {}
        y = F.pad(y, self._pad, mode=self.pad_mode, **kw)
        return y.reshape(*lead, *y.shape[-3:]) if x.dim() > 4 else y

    def unfill(self, x: torch.Tensor) -> torch.Tensor:
        if tuple(x.shape[-2:]) != self.tgt_size:
            return x  # already at the original size (e.g. written un-padded by the fused upsample)
        l, r, t, b = self._pad
        return x[..., t : x.shape[-2] - b, l : x.shape[-1] - r]


class InputScaler:
    """Resize [..., H, W] to a stride multiple / fixed size and back (flow values rescaled)."""

    def __init__(self, orig_shape: Sequence[int], stride: Optional[int] = None, size: Optional[Tuple[int, int]] = None,
                 scale_factor: Optional[float] = 1.0, interpolation_mode: str = "bilinear",
                 interpolation_align_corners: bool = False) -> None:
        self.orig_height, self.orig_width = int(orig_shape[-2]), int(orig_shape[-1])
        if stride is not None:
            assert size is None, "stride and size are mutually exclusive"
            self.tgt_height = int(math.ceil(self.orig_height / stride)) * stride
            self.tgt_width = int(math.ceil(self.orig_width / stride)) * stride
        elif size is not None:
            self.tgt_height, self.tgt_width = int(size[0]), int(size[1])
        else:
            self.tgt_height = int(self.orig_height * scale_factor)
            self.tgt_width = int(self.orig_width * scale_factor)
        self.interpolation_mode = interpolation_mode
        self.interpolation_align_corners = interpolation_align_corners

    def _resize(self, x, h, w, is_flow):
        lead = x.shape[:-3]
        y = x.reshape(-1, *x.shape[-3:])
        src_h, src_w = y.shape[-2:]
        y = F.interpolate(y, size=(h, w), mode=self.interpolation_mode, align_corners=self.interpolation_align_corners)
        if is_flow:
            scale = torch.tensor([w / src_w, h / src_h], dtype=y.dtype, device=y.device).view(1, 2, 1, 1)
            y = y * scale
        return y.reshape(*lead, *y.shape[-3:])

    def fill(self, x, is_flow: bool = False):
        return self._resize(x, self.tgt_height, self.tgt_width, is_flow)

    def unfill(self, x, is_flow: bool = False):
        return self._resize(x, self.orig_height, self.orig_width, is_flow)


def bgr_val_as_tensor(val, reference: torch.Tensor, position: int = -3) -> torch.Tensor:
    """Scalar / 3-vector -> tensor broadcastable against ``reference`` with BGR at ``position``."""
    if not isinstance(val, torch.Tensor):
        if isinstance(val, (int, float)):
            val = [float(val)] * 3
        val = torch.as_tensor(list(val) if not hasattr(val, "shape") else val)
    val = val.to(dtype=reference.dtype, device=reference.device)
    if val.dim() == 1 and val.numel() == 3:
        shape = [1] * reference.dim()
        shape[position] = 3
        val = val.reshape(shape)
    return val


def count_parameters(model: torch.nn.Module) -> int:
    """Trainable parameters (ptlflow/utils/utils.py:262-277; 5 257 536 for raft, the model_benchmark.py column)."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def tensor_dict_to_numpy(tensor_dict, padder: Optional[InputPadder] = None):
    """Every tensor of the dict -> numpy HWC of its first sample / first frame (ptlflow/utils/utils.py:331-361)."""
    out = {}
    for k, v in tensor_dict.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().float().cpu()
            if padder is not None:
                v = padder.unfill(v)
            while v.dim() > 3:
                v = v[0]
            v = v.permute(1, 2, 0).numpy()
        out[k] = v
    return out
