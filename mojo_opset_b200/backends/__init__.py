"""Backend packages.  Only ``b200`` exists; its classes are importable anywhere but register with the
op registry only on an sm_100 platform (reference gating: ``mojo_opset/backends/__init__.py:19-33``)."""

from .b200 import *  # noqa: F401,F403
