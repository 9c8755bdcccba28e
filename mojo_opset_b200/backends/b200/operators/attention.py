from typing import Optional

import torch

from mojo_opset_b200 import functional as F
from mojo_opset_b200.core import MojoPagedDecodeGQA
from mojo_opset_b200.core import MojoPagedDecodeSWA
from mojo_opset_b200.core import MojoPagedPrefillGQA
from mojo_opset_b200.core import MojoPagedPrefillSWA
from mojo_opset_b200.core import MojoSdpa
from mojo_opset_b200.core import MojoSWA
from mojo_opset_b200.core.operators.attention import assert_paged_decode_contract
from mojo_opset_b200.core.operators.attention import assert_paged_prefill_contract


class B200PagedDecodeGQA(MojoPagedDecodeGQA):
    supported_platforms_list = ["b200"]

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
        assert_paged_decode_contract(block_tables, total_seq_lens)
        if not self.is_causal:
            # for one query token causal == full attention; the non-causal + mask variant is not built
            raise NotImplementedError("B200PagedDecodeGQA supports is_causal=True only")
        if mask is not None:
            raise NotImplementedError("B200PagedDecodeGQA does not take a mask")
        return F.paged_decode_gqa(query, key_cache, value_cache, total_seq_lens, block_tables, softmax_scale,
                                  self.gqa_layout, max_total_seq_len)


class B200PagedPrefillGQA(MojoPagedPrefillGQA):
    supported_platforms_list = ["b200"]

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
        assert_paged_prefill_contract(cu_q_lens, block_tables, cu_total_seq_lens)
        if not self.is_causal:
            raise NotImplementedError("B200PagedPrefillGQA supports is_causal=True only")
        if mask is not None:
            raise NotImplementedError("B200PagedPrefillGQA does not take a mask")
        return F.paged_prefill_gqa(query, key_cache, value_cache, cu_q_lens, block_tables, softmax_scale,
                                   cu_total_seq_lens, self.gqa_layout, max_q_len, max_total_seq_len)


class B200PagedPrefillSWA(MojoPagedPrefillSWA):
    supported_platforms_list = ["b200"]

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
        assert_paged_prefill_contract(cu_q_lens, block_table, cu_total_seq_lens)
        if not self.is_causal:  # the reference ignores the windows then: full attention over the paged keys
            raise NotImplementedError("B200PagedPrefillSWA supports is_causal=True only")
        return F.paged_prefill_gqa(query, key_cache, value_cache, cu_q_lens, block_table, softmax_scale,
                                   cu_total_seq_lens, self.gqa_layout, max_q_len, max_total_seq_len, True,
                                   self.local_window_size, self.global_window_size)


class B200PagedDecodeSWA(MojoPagedDecodeSWA):
    supported_platforms_list = ["b200"]

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
        assert_paged_decode_contract(block_table, total_seq_lens)
        if not self.is_causal:
            raise NotImplementedError("B200PagedDecodeSWA supports is_causal=True only")
        return F.paged_decode_swa(query, key_cache, value_cache, total_seq_lens, block_table, softmax_scale,
                                  self.gqa_layout, max_total_seq_len, self.local_window_size, self.global_window_size)


class B200SWA(MojoSWA):
    supported_platforms_list = ["b200"]

    def forward(
        self,
        query: torch.Tensor,
        key: torch.Tensor,
        value: torch.Tensor,
        cu_q_lens: torch.Tensor,
        cu_total_seq_lens: torch.Tensor,
        softmax_scale: Optional[float] = None,
    ):
        assert cu_q_lens.dtype == torch.int32
        assert cu_total_seq_lens.dtype == torch.int32
        if not self.is_causal:  # the reference ignores the windows then: full attention inside each sequence
            raise NotImplementedError("B200SWA supports is_causal=True only")
        return F.swa(query, key, value, cu_q_lens, cu_total_seq_lens, softmax_scale, self.gqa_layout,
                     self.local_window_size, self.global_window_size)


class B200Sdpa(MojoSdpa):
    supported_platforms_list = ["b200"]

    def forward(
        self,
        query: torch.Tensor,
        key: torch.Tensor,
        value: torch.Tensor,
        attn_mask: Optional[torch.Tensor] = None,
    ):
        return F.sdpa(query, key, value, self.scale, self.enable_gqa, attn_mask)
