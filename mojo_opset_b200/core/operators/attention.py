"""Attention ops on the paged-decoder hot path: interface + contracts.

Signatures, constructor arguments, attribute names and contract checks follow the reference's
``mojo_opset/core/operators/attention.py`` (``MojoPagedDecodeGQA`` :113-232, ``MojoPagedPrefillGQA``
:315-451, ``MojoSdpa`` :456-504, contracts :12-37).  The torch-native bodies are NOT reproduced here:
see ``oracle/golden.py`` for the restatement used by the tests.
"""

from typing import Optional

import torch

from ..operator import MojoOperator

_GQA_LAYOUTS = ("ABAB", "AABB")


def assert_paged_prefill_contract(cu_q_lens, block_tables, cu_total_seq_lens) -> None:
    """Reference ``attention.py:12-28``."""
    assert isinstance(cu_q_lens, torch.Tensor)
    assert isinstance(block_tables, torch.Tensor)
    assert cu_q_lens.dtype == torch.int32
    assert block_tables.dtype == torch.int32
    num_seqs = cu_q_lens.shape[0] - 1
    if cu_total_seq_lens is not None:
        assert isinstance(cu_total_seq_lens, torch.Tensor)
        assert cu_total_seq_lens.dtype == torch.int32
        assert cu_total_seq_lens.dim() == 1
        assert cu_total_seq_lens.shape[0] == num_seqs + 1
    assert block_tables.shape[0] == num_seqs
    assert block_tables.dim() == 2


def assert_paged_decode_contract(block_tables, total_seq_lens) -> None:
    """Reference ``attention.py:31-37``."""
    assert isinstance(block_tables, torch.Tensor)
    assert isinstance(total_seq_lens, torch.Tensor)
    assert total_seq_lens.dtype == torch.int32
    assert block_tables.dtype == torch.int32
    assert block_tables.shape[0] == total_seq_lens.shape[0]
    assert block_tables.dim() == 2


class _PagedGQABase:
    def _init_paged(self, is_causal: bool, gqa_layout: str) -> None:
        if gqa_layout not in _GQA_LAYOUTS:
            raise ValueError(f"gqa_layout must be one of ['ABAB', 'AABB'], got {gqa_layout}")
        self.is_causal = is_causal
        self.gqa_layout = gqa_layout

    def extra_repr(self) -> str:
        return f"is_causal={self.is_causal!r}, gqa_layout={self.gqa_layout!r}"


class MojoPagedDecodeGQA(_PagedGQABase, MojoOperator):
    """One query token per sequence against a paged KV cache (``query[B,Hq,D]`` -> ``[B,Hq,D]``).

    ``key_cache/value_cache[N_blocks,Hkv,block_size,D]``, ``total_seq_lens[B] int32``,
    ``block_tables[B,MB] int32`` (-1 = unused).  Rows with ``seq_len <= 0`` produce zeros.
    Head mapping: AABB ``kv = h // G``; ABAB ``kv = h % Hkv``.
    """

    def __init__(self, is_causal: bool = True, gqa_layout: str = "AABB"):
        super().__init__()
        self._init_paged(is_causal, gqa_layout)

    def forward(
        self,
        query: torch.Tensor,
        key_cache: torch.Tensor,
        value_cache: torch.Tensor,
        total_seq_lens: torch.Tensor,
        block_tables: torch.Tensor,
        softmax_scale: Optional[float] = None,
        mask: Optional[torch.Tensor] = None,
        *,
        max_total_seq_len: Optional[int] = None,
    ):
        return MojoOperator.forward(self)


class MojoPagedPrefillGQA(_PagedGQABase, MojoOperator):
    """Var-len causal attention of a query chunk against paged KV holding the chunk (+ cached prefix).

    ``query[T,Hq,D]``, ``cu_q_lens[B+1] int32``, ``cu_total_seq_lens[B+1] int32 | None`` (None: kv_len =
    q_len).  Query row ``t`` of a sequence sees keys ``0 .. kv_len - q_len + t``.
    """

    def __init__(self, is_causal: bool = True, gqa_layout: str = "AABB"):
        super().__init__()
        self._init_paged(is_causal, gqa_layout)

    def forward(
        self,
        query: torch.Tensor,
        key_cache: torch.Tensor,
        value_cache: torch.Tensor,
        cu_q_lens: torch.Tensor,
        block_tables: torch.Tensor,
        softmax_scale: Optional[float] = None,
        cu_total_seq_lens: Optional[torch.Tensor] = None,
        mask: Optional[torch.Tensor] = None,
        max_q_len: Optional[int] = None,
        max_total_seq_len: Optional[int] = None,
    ):
        return MojoOperator.forward(self)


class _PagedSWABase(_PagedGQABase):
    def _init_swa(self, is_causal, gqa_layout, global_window_size, local_window_size) -> None:
        self._init_paged(is_causal, gqa_layout)
        self.gqa_interleave = gqa_layout == "ABAB"
        self.global_window_size = global_window_size
        self.local_window_size = local_window_size

    def extra_repr(self) -> str:
        return (f"is_causal={self.is_causal}, gqa_layout={self.gqa_layout}, "
                f"global_window_size={self.global_window_size}, local_window_size={self.local_window_size}")


class MojoPagedPrefillSWA(_PagedSWABase, MojoOperator):
    """``MojoPagedPrefillGQA`` with a sliding window (reference ``attention.py:533-643``): on top of the causal limit a
    key is visible iff ``key + local_window_size >= position`` or ``key < global_window_size``
    (``_generate_window_mask``, ``:507-531``); a window left ``None`` contributes nothing, both ``None`` = causal."""

    def __init__(self, is_causal: bool = True, gqa_layout: str = "AABB", global_window_size: Optional[int] = None,
                 local_window_size: Optional[int] = None):
        super().__init__()
        self._init_swa(is_causal, gqa_layout, global_window_size, local_window_size)

    def forward(
        self,
        query: torch.Tensor,
        key_cache: torch.Tensor,
        value_cache: torch.Tensor,
        cu_q_lens: torch.Tensor,
        block_table: torch.Tensor,
        softmax_scale: Optional[float] = None,
        cu_total_seq_lens: Optional[torch.Tensor] = None,
        *,
        max_q_len: Optional[int] = None,
        max_total_seq_len: Optional[int] = None,
    ):
        return MojoOperator.forward(self)


class MojoPagedDecodeSWA(_PagedSWABase, MojoOperator):
    """``MojoPagedDecodeGQA`` with the same window rule for the single query token at position ``seq_len - 1``
    (reference ``attention.py:645-745``)."""

    def __init__(self, is_causal: bool = True, gqa_layout: str = "AABB", global_window_size: Optional[int] = None,
                 local_window_size: Optional[int] = None):
        super().__init__()
        self._init_swa(is_causal, gqa_layout, global_window_size, local_window_size)

    def forward(
        self,
        query: torch.Tensor,
        key_cache: torch.Tensor,
        value_cache: torch.Tensor,
        total_seq_lens: torch.Tensor,
        block_table: torch.Tensor,
        softmax_scale: Optional[float] = None,
        *,
        max_total_seq_len: Optional[int] = None,
    ):
        return MojoOperator.forward(self)


class MojoSWA(_PagedSWABase, MojoOperator):
    """Non-paged sliding-window attention over PACKED var-len tensors (reference ``attention.py:747-838``):
    ``query[Tq,Hq,D]``, ``key/value[Tk,Hkv,D]``; sequence ``b`` owns query rows ``cu_q_lens[b]:cu_q_lens[b+1]`` and key
    rows ``cu_total_seq_lens[b]:cu_total_seq_lens[b+1]``; causal with offset ``kv_len - q_len`` plus the window rule of
    the paged SWA ops."""

    def __init__(self, is_causal: bool = True, gqa_layout: str = "AABB", global_window_size: Optional[int] = None,
                 local_window_size: Optional[int] = None):
        super().__init__()
        self._init_swa(is_causal, gqa_layout, global_window_size, local_window_size)

    def forward(
        self,
        query: torch.Tensor,
        key: torch.Tensor,
        value: torch.Tensor,
        cu_q_lens: torch.Tensor,
        cu_total_seq_lens: torch.Tensor,
        softmax_scale: Optional[float] = None,
    ):
        return MojoOperator.forward(self)


class MojoSdpa(MojoOperator):
    """Dense non-causal SDPA ``[B,Hq,Sq,D] x [B,Hkv,Skv,D]`` (inputs may be strided views)."""

    def __init__(self, scale: Optional[float] = None, enable_gqa: bool = False):
        super().__init__()
        self.scale = scale
        self.enable_gqa = enable_gqa

    def forward(
        self,
        query: torch.Tensor,
        key: torch.Tensor,
        value: torch.Tensor,
        attn_mask: Optional[torch.Tensor] = None,
    ):
        return MojoOperator.forward(self)

    def extra_repr(self) -> str:
        return f"scale={self.scale!r}, enable_gqa={self.enable_gqa!r}"
