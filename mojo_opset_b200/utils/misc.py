"""Small helpers shared by the op classes (reference: ``mojo_opset/utils/misc.py:1-33``)."""

import os

_FACTORY_KEYS = ("device", "dtype", "layout", "requires_grad", "pin_memory", "memory_format")


def get_bool_env(key: str, default: bool = True) -> bool:
    raw = os.environ.get(key)
    if raw is None:
        return default
    raw = raw.lower()
    if raw in ("1", "yes", "true"):
        return True
    if raw in ("0", "no", "false"):
        return False
    return default


def get_tensor_factory_kwargs(**kwargs):
    """Keep only the ``torch.empty``-style keyword arguments that were actually given."""
    return {k: v for k, v in kwargs.items() if v is not None and k in _FACTORY_KEYS}
