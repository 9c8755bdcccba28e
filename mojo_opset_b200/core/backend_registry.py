"""Per-operator backend registry.

Behavioural mirror of the reference's ``mojo_opset/core/backend_registry.py:13-124``:

* one registry per core op (``MojoXxx``); the op name is the class name minus ``Mojo``;
* an implementation class is named ``<Prefix><OpName>``; ``prefix.lower()`` is its backend key
  (``B200PagedDecodeGQA`` -> ``"b200"``), reference ``backend_registry.py:48-54``;
* the key must be a known backend for the platform and the platform must be listed in the
  class's ``supported_platforms_list`` (reference ``:65-91``), otherwise it is not registered;
* ``get(name)`` falls back to the highest-priority registered class when ``name`` is ``None``
  or unknown, unless ``strict=True`` (reference ``:93-118``).

What differs on purpose: the only accelerator backend is ``b200`` and there is NO built-in
``torch`` implementation in this package - the torch-native golden lives under ``oracle/`` and
registers itself as backend ``"torch"`` only when a test imports it.
"""

from typing import Dict
from typing import Optional

from mojo_opset_b200.utils.platform import get_platform

PLATFORM_BACKEND_PRIORITY = {
    "b200": ["b200", "torch"],
    "meta_device": ["b200", "torch"],
}


def _priority_list():
    return PLATFORM_BACKEND_PRIORITY.get(get_platform(), ["b200", "torch"])


def _normalize_backend_name(name: Optional[str]) -> Optional[str]:
    if name is None:
        return None
    return name.strip().lower()


class MojoBackendRegistry:
    def __init__(self, core_op_cls):
        assert core_op_cls.__name__.startswith("Mojo"), (
            f"Operator {core_op_cls.__name__} who is a subclass of MojoOperator, class name must start with Mojo."
        )
        self._core_op_cls = core_op_cls
        self._operator_name = core_op_cls.__name__[len("Mojo"):]
        self._registry: Dict[str, type] = {}

    def get_core_op_cls(self):
        return self._core_op_cls

    def register(self, cls) -> None:
        at = cls.__name__.find(self._operator_name)
        assert at != -1, (
            f"Operator {cls.__name__} who be a subclass of {self._core_op_cls.__name__} must "
            f"contain {self._operator_name} in its name."
        )
        backend = _normalize_backend_name(cls.__name__[:at])
        assert backend != "mojo", "should not register base backend"

        known = _priority_list()
        if backend not in known:
            for candidate in known:
                if backend.startswith(candidate):
                    raise NameError(
                        f"Operator {cls.__name__} backend[{backend}] is not supported, "
                        f"are you wish to named {candidate.upper()}{self._operator_name} ?"
                    )
            raise AssertionError(
                f"Operator {cls.__name__} backend[{backend}] is not supported for platform[{get_platform()}], "
                f"please choose from {known}."
            )

        if get_platform() not in getattr(cls, "supported_platforms_list", ()):
            # Not an error: e.g. B200* classes imported on a CPU-only box simply do not register.
            return

        if backend in self._registry:
            raise ValueError(f"Operator {self._core_op_cls.__name__} backend[{backend}] has been registered")

        self._registry[backend] = cls
        cls._backend = backend
        order = {name: i for i, name in enumerate(known)}
        self._registry = dict(sorted(self._registry.items(), key=lambda kv: order.get(kv[0], len(order))))

    def get(self, backend_name: Optional[str] = None, *, strict: bool = False):
        backend_name = _normalize_backend_name(backend_name)
        if backend_name is None or backend_name not in self._registry:
            if strict and backend_name is not None:
                raise KeyError(
                    f"{self._operator_name} backend {backend_name!r} is not registered; "
                    f"available: {list(self._registry)}"
                )
            if not self._registry:
                raise NotImplementedError(
                    f"Mojo{self._operator_name} has no implementation on platform '{get_platform()}': "
                    "the b200 backend needs an sm_100 GPU and there is no CPU fallback."
                )
            return next(iter(self._registry.values()))
        return self._registry[backend_name]

    def registered_backends(self):
        return tuple(self._registry)
