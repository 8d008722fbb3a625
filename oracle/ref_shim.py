"""Load the REAL reference hot-path modules from /root/reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``import ptlflow`` fails in this image
(jsonargparse / lightning / torchmetrics / timm are absent), but the
files on the hot path only need torch + einops + scipy once three tiny stand-ins exist:
``lightning.pytorch.LightningModule`` (used as an nn.Module), ``torchmetrics.Metric``
(``add_state``) and namespace packages that bypass ``ptlflow/__init__.py`` and
``ptlflow/models/__init__.py``.  Recipe verified in SURVEY.md appendix C.

/root/reference does not exist on the GPU box: nothing under tests -m gpu, smoke() or
bench.py may call this module.  It is used by oracle/make_golden.py and by CPU tests that
skip when the reference checkout is absent.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PTLFLOW_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ptlflow", "models", "raft"))


def install() -> None:
    """Pre-seed sys.modules so ``import ptlflow.models.raft.raft`` resolves to the reference files."""
    if "ptlflow.models.raft" in sys.modules and getattr(sys.modules["ptlflow"], "_b200_shim", False):
        return
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    import torch.nn as nn

    lightning, lp = types.ModuleType("lightning"), types.ModuleType("lightning.pytorch")
    lightning.__path__, lp.__path__ = [], []

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    lp.LightningModule = LightningModule
    lightning.pytorch = lp
    sys.modules.setdefault("lightning", lightning)
    sys.modules.setdefault("lightning.pytorch", lp)

    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class Metric(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

            def add_state(self, name, default, dist_reduce_fx=None):
                self.register_buffer(name, default, persistent=False)

        tm.Metric = Metric
        sys.modules["torchmetrics"] = tm

    root = os.path.join(REFERENCE_ROOT, "ptlflow")
    for name, sub in [
        ("ptlflow", ""),
        ("ptlflow.utils", "utils"),
        ("ptlflow.utils.external", "utils/external"),
        ("ptlflow.models", "models"),
        ("ptlflow.models.base_model", "models/base_model"),
        ("ptlflow.models.raft", "models/raft"),
        ("ptlflow.models.gma", "models/gma"),
    ]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(root, sub)]
        sys.modules[name] = m
    sys.modules["ptlflow"]._b200_shim = True


def load_raft():
    """-> the reference module ptlflow.models.raft.raft (classes ``raft``, ``raft_small``)."""
    install()
    import ptlflow.models.raft.raft as ref_raft  # noqa: WPS433  (reference code, imported not copied)

    return ref_raft


def load_raft_corr():
    install()
    import ptlflow.models.raft.corr as ref_corr

    return ref_corr


def load_gma():
    install()
    import ptlflow.models.gma.gma as ref_gma

    return ref_gma


def build_reference_model(variant: str, seed: int = 0, **kwargs):
    """Reference model in eval mode holding oracle.synth weights (fnet./cnet./update_block. keys)."""
    import torch

    from . import synth

    mod = load_gma() if variant == "gma" else load_raft()
    model = getattr(mod, variant)(**kwargs).eval()
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k.split(".")[0] in ("fnet", "cnet", "update_block", "att"):
            new[k] = torch.from_numpy(synth.synth_tensor(k, tuple(v.shape), seed)).to(v.dtype).reshape(v.shape)
        else:
            new[k] = v
    model.load_state_dict(new)
    return model
