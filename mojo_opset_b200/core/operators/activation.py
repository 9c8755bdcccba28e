"""Activation ops on the path (reference ``mojo_opset/core/operators/activation.py:20-66``)."""

import torch

from ..operator import MojoOperator


class MojoGelu(MojoOperator):
    """Exact (erf) GELU element-wise, dtype preserved (reference ``activation.py:6-17``)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return MojoOperator.forward(self)


class MojoSilu(MojoOperator):
    """``x * sigmoid(x)`` element-wise, dtype preserved."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return MojoOperator.forward(self)


class MojoSwiGLU(MojoOperator):
    """``silu(gate) * up`` with the silu rounded to the input dtype before the product.

    ``swiglu_limit > 0`` first clamps ``up`` to ``[-limit, limit]`` and ``gate`` to ``<= limit``.
    """

    def __init__(self, swiglu_limit: float = 0.0, **kwargs):
        super().__init__(**kwargs)
        self.swiglu_limit = swiglu_limit

    def forward(self, gate_out: torch.Tensor, up_out: torch.Tensor) -> torch.Tensor:
        return MojoOperator.forward(self)

    def extra_repr(self) -> str:
        return f"swiglu_limit={self.swiglu_limit!r}"
