"""Register the b200 kernels as backend ``b200`` inside an installed UPSTREAM ``mojo_opset``.

The reference loads out-of-tree backends through the entry-point group ``mojo_opset.plugins``
(``mojo_opset/__init__.py:19-45``; example plugin ``mojo_opset/tests/base/plugin/pyproject.toml:11-12``): a
callable that runs at ``import mojo_opset``.  ``register()`` is that callable::

    [project.entry-points."mojo_opset.plugins"]
    b200 = "mojo_opset_b200.plugin:register"

What it does, all through the reference's own registration rules (``core/backend_registry.py:48-91``):

* upstream knows no NVIDIA platform (``utils/platform.py:16-41`` returns ``"meta_device"`` on a B200 box), so
  ``"b200"`` is put at the head of that platform's backend priority list - in place, because
  ``BACKEND_PRIORITY_LIST`` aliases the same list object (``backend_registry.py:13-21``);
* for every hot-path op it creates ``B200<Op>`` as a DIRECT subclass of upstream's ``Mojo<Op>`` (so the TP wrapper's
  ``type(module).__base__`` lookup, ``distributed/parallel/mojo_parallel.py:63-71``, keeps working) whose ``forward``
  is the one of this package's ``B200<Op>`` class; ``__init_subclass__`` registers it (``core/operator.py:22-36``);
* constructor signatures, attributes and error conventions are upstream's own - only ``forward`` is replaced.

After that ``MOJO_BACKEND=b200`` (or no variable at all: b200 has the highest priority) makes
``mojo_opset.MojoPagedDecodeGQA(...)`` etc. instantiate the CUDA-backed classes.  Nothing falls back: on a box without
an sm_100 GPU the ops raise at call time.
"""

from typing import Dict

OPS = (
    "PagedDecodeGQA",
    "PagedPrefillGQA",
    "PagedPrefillSWA",
    "PagedDecodeSWA",
    "SWA",
    "Sdpa",
    "StorePagedKVCache",
    "ResidualAddRMSNorm",
    "RMSNorm",
    "ApplyRoPE",
    "RotaryEmbedding",
    "SwiGLU",
    "Silu",
    "GemmAllReduce",
    "Gelu",
    "LayerNorm",
    "GridRoPE",
)


def register(upstream=None) -> Dict[str, type]:
    """Create and register ``B200<Op>`` classes on ``upstream`` (default: ``import mojo_opset``).

    Returns {op name: new class}.  Idempotent: ops that already have a ``b200`` implementation are skipped.
    """
    if upstream is None:
        import mojo_opset as upstream  # the reference package

    from mojo_opset.core import backend_registry as up_registry
    from mojo_opset.utils.platform import get_platform as up_get_platform

    from mojo_opset_b200.backends import b200 as ours

    platform = up_get_platform()
    priority = up_registry.PLATFORM_BACKEND_PRIORITY.setdefault(platform, ["torch"])
    if "b200" not in priority:
        priority.insert(0, "b200")
    if "b200" not in up_registry.BACKEND_PRIORITY_LIST:  # not the same list object (other platform key)
        up_registry.BACKEND_PRIORITY_LIST.insert(0, "b200")

    created = {}
    for op in OPS:
        core_cls = getattr(upstream, "Mojo" + op, None)
        if core_cls is None:  # ops the reference still keeps under mojo_opset.experimental (e.g. MojoGridRoPE)
            try:
                import importlib

                core_cls = getattr(importlib.import_module(upstream.__name__ + ".experimental"), "Mojo" + op, None)
            except ImportError:
                core_cls = None
        our_cls = getattr(ours, "B200" + op, None)
        if core_cls is None or our_cls is None:
            continue
        if "b200" in core_cls.get_registered_backends():
            continue
        namespace = {
            "__module__": __name__,
            "__doc__": our_cls.__doc__,
            "supported_platforms_list": [platform],
            "forward": our_cls.forward,
        }
        # private helpers of our backend class itself (lazy state, workspaces) ...
        for name, attr in vars(our_cls).items():
            if name.startswith("_") and not name.startswith("__") and callable(attr):
                namespace.setdefault(name, attr)
        # ... and argument-check helpers our forward() calls on self (defined on this package's mirror of the core op)
        for base in our_cls.__mro__[1:]:
            for name, attr in vars(base).items():
                if name.startswith("_check_") and callable(attr) and not hasattr(core_cls, name):
                    namespace.setdefault(name, attr)
        # class creation == registration (MojoOperator.__init_subclass__)
        created[op] = type("B200" + op, (core_cls,), namespace)
    return created
