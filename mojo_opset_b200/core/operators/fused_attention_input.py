"""Fused pre-attention ops: the reference lists the shells ``MojoRoPEStoreKV`` and ``MojoNormRoPEStoreKV``
(``README.md:128-130``) without a core definition; the signatures below are the composition its in-tree Qwen3 block
performs with four ops (``modeling/qwen3/mojo_qwen3_dense.py:229-234``: ``q_norm``, ``k_norm``, ``rope``; and
``PagedDummyCache.update`` ``:99-109``: ``MojoStorePagedKVCache``), on token-major ``[T, heads, D]`` tensors."""

from typing import Optional

import torch

from ..operator import MojoOperator


class MojoRoPEStoreKV(MojoOperator):
    """``q_rot, k_rot = rope(q, k)``; ``store_paged_kv(k_rot, v)``; returns ``q_rot`` (caches updated in place).

    ``forward(q[T,Hq,D], k[T,Hkv,D], v[T,Hkv,D], cos[T,d], sin[T,d], key_cache, value_cache, block_table,
    cu_q_lens | None, context_kv_lens)`` - the last three as in ``MojoStorePagedKVCache`` (``None`` = decode)."""

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
                key_cache: torch.Tensor, value_cache: torch.Tensor, block_table: torch.Tensor,
                cu_q_lens: Optional[torch.Tensor], context_kv_lens: torch.Tensor) -> torch.Tensor:
        return MojoOperator.forward(self)


class MojoNormRoPEStoreKV(MojoOperator):
    """As ``MojoRoPEStoreKV`` with a per-head RMSNorm (``q_weight`` / ``k_weight`` of size ``head_dim``) in front
    of the rotation - Qwen3's ``q_norm`` / ``k_norm``."""

    def __init__(self, head_dim: int, eps: float = 1e-6, **kwargs):
        super().__init__(**kwargs)
        self.head_dim = head_dim
        self.variance_epsilon = float(eps)
        self.q_weight = torch.nn.Parameter(torch.empty(head_dim, **self.tensor_factory_kwargs))
        self.k_weight = torch.nn.Parameter(torch.empty(head_dim, **self.tensor_factory_kwargs))

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
                key_cache: torch.Tensor, value_cache: torch.Tensor, block_table: torch.Tensor,
                cu_q_lens: Optional[torch.Tensor], context_kv_lens: torch.Tensor) -> torch.Tensor:
        return MojoOperator.forward(self)

    def extra_repr(self) -> str:
        return f"head_dim={self.head_dim!r}, variance_epsilon={self.variance_epsilon!r}"
