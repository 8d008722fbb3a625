"""IOAdapter: numpy frames -> the model's input dict and back, with the call shape of the reference's
ptlflow/utils/io_adapter.py:29-200 as infer.py:141-196 / validate.py use it:

    io_adapter = IOAdapter(model.output_stride, img.shape[:2], target_size=None, cuda=True, fp16=True)
    inputs = io_adapter.prepare_inputs([img1, img2])       # {"images": [1, 2, 3, H, W]}  BGR, [0, 1]
    preds = io_adapter.unscale(model(inputs))

Host-side plumbing only: HWC arrays become CHW float tensors in [0, 1], lists are
stacked, a leading batch axis is added until the tensors are 5-D, optional rescaling goes through ``InputScaler``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from .utils import InputScaler


def _to_tensor(v) -> torch.Tensor:
    """HWC array (or a list of them) -> float CHW tensor (stacked); uint8 images are divided by 255."""
    if isinstance(v, (list, tuple)):
        return torch.stack([_to_tensor(x) for x in v], dim=0)
    a = np.asarray(v)
    if a.ndim == 2:
        a = a[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))
    return t.float() / 255.0 if a.dtype == np.uint8 else t.float()


class IOAdapter:
    def __init__(self, output_stride: int, input_size: Tuple[int, int], target_size: Optional[Tuple[int, int]] = None,
                 target_scale_factor: Optional[float] = None, interpolation_mode: str = "bilinear",
                 interpolation_align_corners: bool = False, cuda: bool = False, fp16: bool = False) -> None:
        self.output_stride = output_stride
        self.target_size, self.target_scale_factor = target_size, target_scale_factor
        self.cuda, self.fp16 = cuda, fp16
        self.scaler = None
        if (target_size is not None and min(target_size) > 0) or (target_scale_factor is not None and target_scale_factor > 0):
            self.scaler = InputScaler(orig_shape=input_size, size=target_size, scale_factor=target_scale_factor,
                                      interpolation_mode=interpolation_mode, interpolation_align_corners=interpolation_align_corners)

    def prepare_inputs(self, images: Optional[Union[np.ndarray, List[np.ndarray]]] = None,
                       flows: Optional[Union[np.ndarray, List[np.ndarray]]] = None, inputs: Optional[Dict[str, Any]] = None,
                       image_only: bool = False, **kwargs) -> Dict[str, torch.Tensor]:
        if inputs is None:
            raw = {"images": images, "flows": flows}
            raw.update(kwargs)
            inputs = {k: _to_tensor(v) for k, v in raw.items() if v is not None and len(v) > 0}
        if self.cuda and torch.cuda.is_available():
            inputs = {k: v.cuda() if isinstance(v, torch.Tensor) else v for k, v in inputs.items()}
            if self.fp16:
                inputs = {k: v.half() if isinstance(v, torch.Tensor) and ("flow" in k or "image" in k) else v for k, v in inputs.items()}
        for k, v in list(inputs.items()):
            if (image_only and k != "images") or not isinstance(v, torch.Tensor):
                continue
            while v.dim() < 5:
                v = v.unsqueeze(0)
            if self.scaler is not None:
                v = self.scaler.fill(v, is_flow=k.startswith("flow"))
            inputs[k] = v
        return inputs

    def unscale(self, outputs: Dict[str, Any], image_only: bool = False) -> Dict[str, Any]:
        for k, v in list(outputs.items()):
            if (image_only and k != "images") or not isinstance(v, torch.Tensor):
                continue
            if self.scaler is not None and v.dim() >= 4:
                v = self.scaler.unfill(v, is_flow=k.startswith("flow"))
            outputs[k] = v
        return outputs
