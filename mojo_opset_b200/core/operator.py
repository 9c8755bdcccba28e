"""``MojoOperator``: the drop-in boundary.

Same dispatch contract as the reference's ``mojo_opset/core/operator.py:18-135``:

* a class that directly subclasses ``MojoOperator`` is a *core op*: it owns a registry and carries
  the constructor / ``forward`` signature every backend must keep;
* a class that subclasses a core op is a *backend implementation* and registers itself by name;
* ``MojoXxx(...)`` resolves ``os.environ["MOJO_BACKEND"]`` on every instantiation
  (reference ``operator.py:38-51``) and returns an instance of the selected backend class;
* ``forward_diff_with`` is the A/B parity harness the accuracy tests use (reference ``:81-129``).

Unlike the reference, a core op here has no torch-native body: ``forward`` on the core class raises.
The golden lives in ``oracle/`` (test infrastructure) and the product path is CUDA only.
"""

import os

from typing import Optional

import torch

from mojo_opset_b200.utils.acc import check_tol_diff
from mojo_opset_b200.utils.misc import get_tensor_factory_kwargs


class MojoOperator(torch.nn.Module):
    supported_platforms_list = ["b200", "meta_device"]
    _backend = None

    def __init_subclass__(cls, **kwargs):
        kwargs.pop("default_priority", None)
        super().__init_subclass__(**kwargs)
        if MojoOperator in cls.__bases__:
            from mojo_opset_b200.core.backend_registry import MojoBackendRegistry

            cls._registry = MojoBackendRegistry(cls)
        else:
            cls._registry.register(cls)

    def __new__(cls, *args, **kwargs):
        if MojoOperator in cls.__bases__:
            target = cls._registry.get(os.environ.get("MOJO_BACKEND"))
            return target.__new__(target, *args, **kwargs)
        return super().__new__(cls)

    @classmethod
    def get_registry(cls):
        if getattr(cls, "_registry", None) is None:
            raise NotImplementedError(f"No {cls.__name__} implementation found, please register at least one.")
        return cls._registry

    @classmethod
    def get_backend_impl(cls, backend_name: Optional[str] = None, *, strict: bool = False):
        return cls.get_registry().get(backend_name, strict=strict)

    @classmethod
    def get_registered_backends(cls):
        return cls.get_registry().registered_backends()

    def __init__(self, **kwargs):
        torch.nn.Module.__init__(self)
        self.tensor_factory_kwargs = get_tensor_factory_kwargs(**kwargs)

    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__}.forward: the core op carries the interface only; instantiate it through "
            "MOJO_BACKEND=b200 on an sm_100 GPU (the torch-native golden is test infrastructure under oracle/)."
        )

    def forward_diff_with(
        self,
        other_op,
        *args,
        atol: float = 1e-2,
        rtol: float = 1e-2,
        ptol: float = 1.0,
        random_seed: int = 42,
        mixed_tol: bool = False,
        **kwargs,
    ):
        if type(self) is type(other_op):
            raise NotImplementedError(
                f"No dedicated backend for {type(self).__name__}; both operands resolve to the same implementation."
            )

        def _cloned(values):
            return [v.clone() if isinstance(v, torch.Tensor) else v for v in values]

        os.environ["PYTHONHASHSEED"] = str(random_seed)
        torch.manual_seed(random_seed)
        mine = self.forward(*_cloned(args), **dict(zip(kwargs, _cloned(kwargs.values()))))
        torch.manual_seed(random_seed)
        theirs = other_op.forward(*_cloned(args), **dict(zip(kwargs, _cloned(kwargs.values()))))

        assert mine is not None, "forward should return a non-None value."
        assert theirs is not None, "comparison operator should return a non-None value."
        check_tol_diff(mine, theirs, atol, rtol, ptol, mixed_tol)
        return mine

    def extra_repr(self) -> str:
        return ""
