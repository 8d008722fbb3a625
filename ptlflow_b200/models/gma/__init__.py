from .gma import gma  # noqa: F401
