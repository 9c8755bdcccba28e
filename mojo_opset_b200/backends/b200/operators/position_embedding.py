import torch

from mojo_opset_b200 import functional as F
from mojo_opset_b200.core import MojoApplyRoPE
from mojo_opset_b200.core import MojoGridRoPE
from mojo_opset_b200.core import MojoRotaryEmbedding


class B200ApplyRoPE(MojoApplyRoPE):
    supported_platforms_list = ["b200"]

    def forward(self, q, k, cos, sin, head_first: bool = True):
        self._check_rope_args(q, k, cos, sin)
        return F.apply_rope(q, k, cos, sin, head_first)


class B200RotaryEmbedding(MojoRotaryEmbedding):
    supported_platforms_list = ["b200"]

    def forward(self, x, cu_q_lens=None, total_seq_lens=None, position_ids=None):
        self._check_rotary_args(x, cu_q_lens, total_seq_lens, position_ids)
        table = {}
        if self.init_max_length is not None:
            table = dict(table_cos=self.cos, table_sin=self.sin)
        inv_freq = self.inv_freq if self.inv_freq.is_cuda else self.inv_freq.to(x.device)
        if cu_q_lens is not None:  # var-len prefill: one row per token of x [T, H]
            return F.rotary_cos_sin(x.shape[0], (x.shape[0],), inv_freq, self.attention_scaling,
                                    cu_q_lens=cu_q_lens, total_seq_lens=total_seq_lens, **table)
        if position_ids is not None:  # decode / explicit ids, any leading shape
            return F.rotary_cos_sin(position_ids.numel(), tuple(position_ids.shape), inv_freq,
                                    self.attention_scaling, position_ids=position_ids.reshape(-1), **table)
        seq = x.shape[1]  # padded prefill [B, S, H] -> [S, d]
        return F.rotary_cos_sin(seq, (seq,), inv_freq, self.attention_scaling, period=seq, **table)


class B200GridRoPE(MojoGridRoPE):
    supported_platforms_list = ["b200"]

    def forward(self, x, grid_sizes, freqs_list):
        # a sample's sequence length F*H*W is the length of its phase table: no host read of grid_sizes
        assert x.dim() == 4, "x must be 4D: [B, L, N, D]"
        assert x.size(-1) % 2 == 0, "D must be even for complex pairing"
        assert grid_sizes.dim() == 2 and grid_sizes.size(1) == 3, "grid_sizes must be [B, 3]"
        return F.grid_rope(x, freqs_list)
