"""BaseModel: the model contract of the reference, kept at the boundary.

Surface mirrored from ptlflow/models/base_model/base_model.py:62-320:
constructor arguments (-> hparams), ``preprocess_images``, ``postprocess_predictions``,
``forward(dict) -> dict`` with ``flows`` [B,N,2,H,W], and the attributes the scripts read
(``output_stride``, ``train_size``, ``warm_start``, ``metric_interpolate_pred_to_target_size``,
``val_metrics``).  Training (``training_step``/``configure_optimizers``) is outside the hot path
 and raises.
"""
from __future__ import annotations

from abc import abstractmethod
from typing import Any, Callable, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from ...utils.lightning_compat import LightningModule
from ...utils.utils import InputPadder, InputScaler, bgr_val_as_tensor


class BaseModel(LightningModule):
    def __init__(
        self,
        output_stride: int,
        loss_fn: Optional[Callable] = None,
        lr: Optional[float] = None,
        wdecay: Optional[float] = None,
        warm_start: bool = False,
        metric_interpolate_pred_to_target_size: bool = False,
    ) -> None:
        super().__init__()
        self.output_stride = output_stride
        self.loss_fn = loss_fn
        self.lr = lr
        self.wdecay = wdecay
        self.warm_start = warm_start
        self.metric_interpolate_pred_to_target_size = metric_interpolate_pred_to_target_size
        self._train_size = None
        self.train_avg_length = None
        self.extra_params = None
        self.val_metrics = nn.ModuleList()
        self.val_dataset_names = []
        self.prev_preds = None
        self.has_trained_on_ptlflow = False
        self.save_hyperparameters(ignore=["loss_fn"])

    # -- attributes the reference scripts touch -------------------------------------------
    @property
    def train_size(self):
        return self._train_size

    @train_size.setter
    def train_size(self, value):
        if value is not None:
            if not (isinstance(value, (tuple, list)) and len(value) == 2 and all(isinstance(v, int) for v in value)):
                raise AssertionError("train_size must be a pair of ints")
        self._train_size = value

    def add_extra_param(self, name, value):
        if self.extra_params is None:
            self.extra_params = {}
        self.extra_params[name] = value

    # -- pre / post processing ---------------------------------------------------------------
    def preprocess_images(
        self,
        images: torch.Tensor,
        stride: Optional[int] = None,
        bgr_add=0,
        bgr_mult=1,
        bgr_to_rgb: bool = False,
        image_resizer: Optional[Union[InputPadder, InputScaler]] = None,
        resize_mode: str = "pad",
        target_size: Optional[Tuple[int, int]] = None,
        pad_mode: str = "replicate",
        pad_value: float = 0.0,
        pad_two_side: bool = True,
        interpolation_mode: str = "bilinear",
        interpolation_align_corners: bool = True,
    ):
        """(images + bgr_add) * bgr_mult, optional BGR->RGB, then pad/resize to the stride.
        The caller's tensor is never modified (base_model.py:206-214 works on a copy too)."""
        x = (images + bgr_val_as_tensor(bgr_add, images)) * bgr_val_as_tensor(bgr_mult, images)
        if bgr_to_rgb:
            x = torch.flip(x, [-3])
        stride = self.output_stride if stride is None else stride
        if target_size is not None:
            stride = None
        if image_resizer is None:
            if resize_mode == "pad":
                image_resizer = InputPadder(x.shape, stride=stride, size=target_size, pad_mode=pad_mode,
                                            two_side_pad=pad_two_side, pad_value=pad_value)
            elif resize_mode == "interpolation":
                image_resizer = InputScaler(x.shape, stride=stride, size=target_size, interpolation_mode=interpolation_mode,
                                            interpolation_align_corners=interpolation_align_corners)
            else:
                raise ValueError(f"resize_mode must be one of (pad, interpolation). Found: {resize_mode}.")
        x = image_resizer.fill(x).contiguous()
        return x, image_resizer

    def postprocess_predictions(self, prediction: torch.Tensor, image_resizer, is_flow: bool) -> torch.Tensor:
        if isinstance(image_resizer, InputScaler):
            return image_resizer.unfill(prediction, is_flow=is_flow)
        return image_resizer.unfill(prediction)

    @abstractmethod
    def forward(self, *args: Any, **kwargs: Any) -> Dict[str, torch.Tensor]:
        ...

    # -- evaluation plumbing (validate.py / test.py call these) ------------------------------
    def validation_step(self, batch: Dict[str, Any], batch_idx: int = 0, dataloader_idx: int = 0) -> Dict[str, Any]:
        """forward + warm-start bookkeeping (base_model.py:366-430) + end-point error when ground truth is there."""
        # ordering of base_model.py:396-430: prev_preds goes into the forward unconditionally; AFTER the forward it is
        # dropped when this batch starts a sequence and replaced by this batch's (detached) predictions otherwise
        if self.warm_start:
            batch = dict(batch, prev_preds=self.prev_preds)
        preds = self(batch)
        if self.warm_start:
            meta = batch.get("meta") or {}
            if "is_seq_start" in meta and meta["is_seq_start"][0]:
                self.prev_preds = None
            else:
                self.prev_preds = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in preds.items()}
        out = {"preds": preds}
        if "flows" in batch:
            gt = batch["flows"].to(preds["flows"].device, torch.float32)
            epe = torch.linalg.norm(preds["flows"].float() - gt, dim=2)
            if "valids" in batch:
                v = batch["valids"].to(epe.device)[:, :, 0] >= 0.5
                out["metrics"] = {"val/epe": (epe * v).sum() / v.sum().clamp(min=1)}
            else:
                out["metrics"] = {"val/epe": epe.mean()}
        return out

    def test_step(self, batch: Dict[str, Any], batch_idx: int = 0) -> Dict[str, Any]:
        return self.validation_step(batch, batch_idx)["preds"]

    def training_step(self, *a, **k):
        raise NotImplementedError("ptlflow_b200 covers the inference hot path; training/backward kernels are out of scope (SURVEY.md section 8f)")

    def configure_optimizers(self):
        raise NotImplementedError("ptlflow_b200 covers the inference hot path; training is out of scope (SURVEY.md section 8f)")
