"""RMSNorm ops on the path (reference ``mojo_opset/core/operators/normalization.py:71-111,308-362``)."""

import torch

from ..operator import MojoOperator


class MojoLayerNorm(MojoOperator):
    """``F.layer_norm`` over the last dim with an optional elementwise affine (reference
    ``normalization.py:19-66``)."""

    def __init__(self, norm_size: int, eps: float = 1e-5, elementwise_affine: bool = True, **kwargs):
        super().__init__(**kwargs)
        self.norm_size = norm_size
        self.elementwise_affine = elementwise_affine
        if elementwise_affine:
            self.weight = torch.nn.Parameter(torch.empty(norm_size, **self.tensor_factory_kwargs))
            self.bias = torch.nn.Parameter(torch.empty(norm_size, **self.tensor_factory_kwargs))
        else:
            self.weight = None
            self.bias = None
        self.variance_epsilon = eps

    def forward(self, hidden_state: torch.Tensor) -> torch.Tensor:
        return MojoOperator.forward(self)

    def extra_repr(self) -> str:
        return (f"norm_size={self.norm_size!r}, variance_epsilon={self.variance_epsilon!r}, "
                f"elementwise_affine={self.elementwise_affine!r}")


class MojoRMSNorm(MojoOperator):
    """``y = x * rsqrt(mean(x^2) + eps) * weight`` over the last dim, fp32 math, one rounding."""

    def __init__(self, norm_size: int, eps: float = 1e-5, **kwargs):
        super().__init__(**kwargs)
        self.norm_size = norm_size
        self.weight = torch.nn.Parameter(torch.empty(norm_size, **self.tensor_factory_kwargs))
        self.variance_epsilon = eps

    def forward(self, hidden_state: torch.Tensor) -> torch.Tensor:
        return MojoOperator.forward(self)

    def extra_repr(self) -> str:
        return f"norm_size={self.norm_size!r}, variance_epsilon={self.variance_epsilon!r}"


class MojoResidualAddRMSNorm(MojoOperator):
    """Fused residual add + RMSNorm.

    ``pre``:  ``r = x + res`` (rounded to the input dtype); ``y = rms_norm(r) * w``; returns ``(y, r)``.
    ``post``: ``h = x + res``; ``y = rms_norm(h) * w``; returns ``(y, y)`` (the same tensor twice).
    """

    def __init__(self, norm_size: int, eps: float = 1e-05, norm_pos: str = "pre", **kwargs):
        super().__init__(**kwargs)
        if norm_pos not in ("pre", "post"):
            raise ValueError("norm_pos should be 'pre' or 'post'")
        self.norm_size = norm_size
        self.variance_epsilon = float(eps)
        self.weight = torch.nn.Parameter(torch.empty(norm_size, **self.tensor_factory_kwargs))
        self.norm_pos = norm_pos

    def forward(self, hidden_state: torch.Tensor, residual: torch.Tensor):
        return MojoOperator.forward(self)

    def extra_repr(self) -> str:
        return (
            f"norm_size={self.norm_size!r}, variance_epsilon={self.variance_epsilon!r}, norm_pos={self.norm_pos!r}"
        )
