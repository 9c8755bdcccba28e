"""Fused compute + collective op on the path: the row-parallel ``o_proj`` of the tensor-parallel attention block
(reference ``mojo_opset/core/operators/compute_with_comm.py:57-117``)."""

from typing import Optional

import torch
import torch.distributed as dist

from ..operator import MojoOperator


class MojoGemmAllReduce(MojoOperator):
    """``output = all_reduce_sum(input @ weight [+ bias])`` over ``process_group``.

    Each rank holds a column shard of the input features and the matching shard of the weight:
    ``trans_weight=False`` -> ``weight[out_features, in_features_local]``, ``True`` -> ``[in_features_local,
    out_features]``.  The bias is added on every rank before the reduction, as in the reference.  Without an
    initialised process group the all-reduce is the identity."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, trans_weight: bool = False,
                 process_group: Optional[dist.ProcessGroup] = None):
        super().__init__()
        if not isinstance(trans_weight, bool):
            raise TypeError("trans_weight must be bool.")
        self.weight = weight
        self.bias = bias
        self.trans_weight = trans_weight
        self.process_group = process_group

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return MojoOperator.forward(self)

    def extra_repr(self) -> str:
        weight_shape = tuple(self.weight.shape) if isinstance(self.weight, torch.Tensor) else None
        has_bias = self.bias is not None
        return f"{weight_shape=}, {has_bias=}, {self.trans_weight=}".replace("self.", "")
