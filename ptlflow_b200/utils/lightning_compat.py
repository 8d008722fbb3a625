"""``LightningModule`` if lightning is installed, else a minimal stand-in.

The reference's BaseModel is a ``pl.LightningModule`` (ptlflow/models/base_model/base_model.py:62)
but the inference path uses only three things from it: ``save_hyperparameters`` (hparams end up in
checkpoints and are read by ``restore_model``), ``log`` and ``log_dict``.  Lightning is not part of
this image, so the stand-in provides exactly those on top of ``nn.Module``.
"""
from __future__ import annotations

import inspect
from argparse import Namespace

import torch.nn as nn

try:  # pragma: no cover - depends on the environment
    from lightning.pytorch import LightningModule as _PL

    HAVE_LIGHTNING = True
except Exception:  # lightning absent (this image) or broken
    _PL = None
    HAVE_LIGHTNING = False


class _MiniLightningModule(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self._hparams = Namespace()

    @property
    def hparams(self) -> Namespace:
        return self._hparams

    def save_hyperparameters(self, *names, ignore=None, **_):
        """Collect the constructor arguments of the outermost ``__init__`` on the call stack that
        belongs to this object (what lightning's ``collect_init_args`` does)."""
        ignore = set([ignore] if isinstance(ignore, str) else (ignore or []))
        found = {}
        frame = inspect.currentframe().f_back
        while frame is not None:
            loc = frame.f_locals
            if frame.f_code.co_name == "__init__" and loc.get("self") is self:
                info = inspect.getargvalues(frame)
                for a in info.args[1:]:
                    found.setdefault(a, loc[a])
                if info.keywords:
                    for k, v in loc[info.keywords].items():
                        found.setdefault(k, v)
            frame = frame.f_back
        for k in list(found):
            if k in ignore or k.startswith("_") or k == "__class__":
                found.pop(k)
        if names:
            found = {k: v for k, v in found.items() if k in names}
        self._hparams = Namespace(**found)

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


LightningModule = _PL if HAVE_LIGHTNING else _MiniLightningModule
