"""``MojoStorePagedKVCache``: interface, contracts and the chunk-plan builder.

Follows the reference's ``mojo_opset/core/operators/kv_cache.py`` (contracts :9-30, plan builder :33-101,
op :104-171).  A *chunk* is one row ``(src_token_start, dst_block_id, dst_block_offset, chunk_len)``:
``chunk_len`` consecutive new tokens that land in one physical block.
"""

from typing import Optional
from typing import Tuple

import torch

from ..operator import MojoOperator


def assert_paged_kv_store_contract(chunk_metadata: torch.Tensor) -> None:
    assert chunk_metadata.dtype == torch.int32
    assert chunk_metadata.dim() == 2
    assert chunk_metadata.shape[1] == 4


def assert_paged_kv_layout_contract(block_table, cu_q_lens, context_kv_lens) -> None:
    assert block_table.dtype == torch.int32
    assert block_table.dim() == 2
    if cu_q_lens is not None:
        assert cu_q_lens.dtype == torch.int32
        assert cu_q_lens.dim() == 1
    if context_kv_lens is not None:
        assert context_kv_lens.dtype == torch.int32
        assert context_kv_lens.dim() == 1
        assert block_table.shape[0] == context_kv_lens.shape[0]


def build_paged_kv_chunk_metadata(
    block_table: torch.Tensor,
    cu_q_lens: Optional[torch.Tensor],
    context_kv_lens: torch.Tensor,
    block_size: int,
) -> torch.Tensor:
    """Store plan ``[num_chunks, 4] int32`` for the new tokens of every sequence.

    Decode mode (``cu_q_lens is None``): one token per sequence at position ``context``.
    Prefill mode: sequence ``i`` appends ``q_i`` tokens at positions ``context_i .. context_i+q_i-1``;
    one chunk per physical block touched.  Rows with a negative context, a negative block id, an
    out-of-table logical block or ``q_len == 0`` are dropped.  Output order: by sequence, then block.
    (Reference ``kv_cache.py:33-101``.)  Note the compaction makes the shape data dependent: call it
    once per step outside CUDA graphs, or use the table-driven store path which needs no plan.
    """
    assert_paged_kv_layout_contract(block_table, cu_q_lens, context_kv_lens)
    num_seqs = context_kv_lens.shape[0]
    if cu_q_lens is not None:
        assert cu_q_lens.shape[0] == num_seqs + 1
    dev = block_table.device
    width = block_table.shape[1]
    if num_seqs == 0 or width == 0:
        return torch.empty((0, 4), dtype=torch.int32, device=dev)

    ctx = context_kv_lens.to(torch.int32)
    if cu_q_lens is None:
        seq = torch.arange(num_seqs, dtype=torch.int32, device=dev)
        pos = ctx.clamp_min(0)
        logical = torch.div(pos, block_size, rounding_mode="floor")
        physical = block_table[seq.long(), logical.clamp(0, width - 1).long()]
        keep = (ctx >= 0) & (logical < width) & (physical >= 0)
        plan = torch.stack((seq, physical, pos % block_size, torch.ones_like(seq)), dim=-1)
        return plan[keep]

    first_tok = cu_q_lens[:-1].to(torch.int32)
    q_lens = (cu_q_lens[1:] - cu_q_lens[:-1]).to(torch.int32)
    blk_lo = torch.arange(width, dtype=torch.int32, device=dev).unsqueeze(0) * block_size
    lo = torch.maximum(ctx.unsqueeze(1), blk_lo)
    hi = torch.minimum((ctx + q_lens).unsqueeze(1), blk_lo + block_size)
    length = (hi - lo).clamp_min(0)
    keep = (q_lens > 0).unsqueeze(1) & (ctx >= 0).unsqueeze(1) & (length > 0) & (block_table >= 0)
    plan = torch.stack(
        (first_tok.unsqueeze(1) + (lo - ctx.unsqueeze(1)), block_table, lo - blk_lo, length), dim=-1
    )
    return plan[keep]


class MojoStorePagedKVCache(MojoOperator):
    """Scatter ``key_states/value_states[T,Hkv,D]`` into ``key_cache/value_cache[NB,Hkv,bs,D]`` in place.

    Either ``chunk_metadata[C,4] int32`` (keyword only) or the legacy triple
    ``(block_table[B,MB], cu_q_lens[B+1] | None, context_kv_lens[B])``; never both.
    Returns the same two cache tensors.
    """

    def __init__(self):
        super().__init__()

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
        return MojoOperator.forward(self)

    @staticmethod
    def _check_store_args(key_states, value_states, block_table, cu_q_lens, context_kv_lens, chunk_metadata):
        """Argument contract shared by every backend (reference ``kv_cache.py:139-156``)."""
        assert key_states.dim() == 3 and value_states.dim() == 3 and key_states.shape == value_states.shape, (
            "key/value states must be (token_num, kv_head_num, head_dim), please check."
        )
        if chunk_metadata is None:
            assert block_table is not None, "block_table is required when chunk_metadata is not provided."
            assert context_kv_lens is not None, "context_kv_lens is required when chunk_metadata is not provided."
            assert_paged_kv_layout_contract(block_table, cu_q_lens, context_kv_lens)
        else:
            assert block_table is None and cu_q_lens is None and context_kv_lens is None, (
                "chunk_metadata path should not be mixed with block_table/cu_q_lens/context_kv_lens."
            )
            assert_paged_kv_store_contract(chunk_metadata)
