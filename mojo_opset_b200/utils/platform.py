"""Platform detection for the B200 backend.

Mirrors the role of the reference's ``mojo_opset/utils/platform.py:16-75`` (``get_platform``,
``get_torch_device``, ``get_dist_backend``) but knows exactly one accelerator: an NVIDIA
Blackwell part with compute capability 10.x ("b200").  Everything else is "meta_device",
on which no B200 op registers (the ops fail loudly instead of falling back).
"""

import functools
import os

import torch

PLATFORM_B200 = "b200"
PLATFORM_NONE = "meta_device"


@functools.lru_cache
def get_platform() -> str:
    forced = os.environ.get("MOJO_PLATFORM")
    if forced:
        return forced.strip().lower()
    try:
        if torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10:
            return PLATFORM_B200
    except Exception:  # pragma: no cover - driver hiccup == no accelerator
        pass
    return PLATFORM_NONE


@functools.lru_cache
def get_torch_device() -> str:
    return "cuda" if get_platform() == PLATFORM_B200 else "meta"


@functools.lru_cache
def get_dist_backend() -> str:
    """nccl over NVLink5/NVSwitch on B200, gloo anywhere else (CPU tests)."""
    return "nccl" if get_platform() == PLATFORM_B200 else "gloo"
