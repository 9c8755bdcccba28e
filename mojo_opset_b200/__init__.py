"""mojo_opset_b200 - a B200 (sm_100a) backend for Mojo Opset's paged-attention decoder hot path.

Layout
------
``core/``       the operator registry + the op interfaces the path needs (same dispatch contract, ctor and
                ``forward`` signatures as the reference's ``mojo_opset/core``); selected by ``MOJO_BACKEND``.
``backends/b200/operators``  the ``B200*`` op classes: argument checks, output allocation, stream plumbing.
``functional``  tensor-level entry points that marshal ``data_ptr()``/strides into the C ABI.
``csrc/``       hand-written CUDA for sm_100a behind ``extern "C"`` entry points (``include/mojo_b200.h``).
``plugin``      registers the same kernels as backend ``b200`` inside an installed upstream ``mojo_opset``.

There is no CPU fallback anywhere: without an sm_100 GPU and the built ``libmojo_b200.so`` the ops raise.
"""

from . import core
from .core import *  # noqa: F401,F403
from . import backends  # noqa: F401  (registers the B200* classes on an sm_100 platform)
from ._lib import check_device_errors  # noqa: F401  (data errors the kernels flagged on the device -> ValueError)

__version__ = "0.1.0"
