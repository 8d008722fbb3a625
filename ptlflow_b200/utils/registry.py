"""Model registry with the reference's decorator surface (ptlflow/utils/registry.py:32-56).

``@register_model`` records the class under its name; ``get_model_reference`` /
``get_model_names`` in the package root read from here.  Lightning is optional in this image,
so ``RegisteredModel`` derives from whatever ``BaseModel`` derives from.
"""
from __future__ import annotations

import sys
from typing import Dict, List

from .lightning_compat import LightningModule

_models_dict: Dict[str, type] = {}
_trainable_models: List[str] = []
_ptlflow_trained_models: List[str] = []


class RegisteredModel(LightningModule):
    """Marker base so CLI code can ask ``issubclass(cls, RegisteredModel)``."""


def register_model(model_class):
    name = model_class.__name__
    package = sys.modules.get(model_class.__module__.rpartition(".")[0])
    if package is not None:
        exported = getattr(package, "__all__", None)
        if exported is None:
            package.__all__ = [name]
        elif name not in exported:
            exported.append(name)
    _models_dict[name] = model_class
    registered = type(name, (model_class, RegisteredModel), {"__module__": model_class.__module__})
    return registered


def trainable(model_class):
    _trainable_models.append(model_class.__name__)
    return model_class


def ptlflow_trained(model_class):
    _ptlflow_trained_models.append(model_class.__name__)
    return model_class
