from typing import Optional
from typing import Tuple

import torch

from mojo_opset_b200 import functional as F
from mojo_opset_b200.core import MojoStorePagedKVCache


class B200StorePagedKVCache(MojoStorePagedKVCache):
    """Chunk-plan path: one CTA per plan row.  Legacy ``(block_table, cu_q_lens, context_kv_lens)`` path: the
    plan is never materialised (no boolean-mask compaction, no host sync) - each new token resolves its slot
    on the device, which keeps the op CUDA-graph capturable."""

    supported_platforms_list = ["b200"]

    def forward(
        self,
        key_states: torch.Tensor,
        value_states: torch.Tensor,
        key_cache: torch.Tensor,
        value_cache: torch.Tensor,
        block_table: Optional[torch.Tensor] = None,
        cu_q_lens: Optional[torch.Tensor] = None,
        context_kv_lens: Optional[torch.Tensor] = None,
        *,
        chunk_metadata: Optional[torch.Tensor] = None,
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        self._check_store_args(key_states, value_states, block_table, cu_q_lens, context_kv_lens, chunk_metadata)
        if cu_q_lens is not None and context_kv_lens is not None:
            assert cu_q_lens.shape[0] == context_kv_lens.shape[0] + 1
        return F.store_paged_kv(key_states, value_states, key_cache, value_cache, chunk_metadata=chunk_metadata,
                                block_table=block_table, cu_q_lens=cu_q_lens, context_kv_lens=context_kv_lens)
