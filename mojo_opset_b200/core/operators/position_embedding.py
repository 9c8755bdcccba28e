"""Rotary position embedding ops (reference ``mojo_opset/core/operators/position_embedding.py:9-175``)."""

from typing import Optional
from typing import Tuple

import torch

from ..operator import MojoOperator


class MojoRotaryEmbedding(MojoOperator):
    """cos/sin generator: ``inv_freq`` in fp32, optional precomputed table of ``init_max_length`` rows.

    Positions come from exactly one of: ``cu_q_lens`` (+ optional ``total_seq_lens`` for the cached
    prefix offset) for var-len prefill, ``position_ids`` for decode, or ``arange(x.shape[1])`` for padded
    prefill (reference ``:43-95``).  The table / ``inv_freq`` are built once at construction time.
    """

    def __init__(self, rope_theta, rope_dim, attention_scaling: float = 1.0,
                 init_max_length: Optional[int] = None, **kwargs):
        super().__init__(**kwargs)
        dev = self.tensor_factory_kwargs.get("device")
        self.rope_theta = rope_theta
        self.attention_scaling = attention_scaling
        self.init_max_length = None
        exponent = torch.arange(0, rope_dim, 2, dtype=torch.float32, device=dev) / rope_dim
        self.register_buffer("inv_freq", 1.0 / (rope_theta ** exponent), persistent=False)
        if init_max_length is not None:
            self._rope_init(init_max_length)

    def _angles(self, position_ids: torch.Tensor):
        freqs = position_ids[..., None] * self.inv_freq[None, :]
        emb = torch.cat((freqs, freqs), dim=-1)
        return emb.cos() * self.attention_scaling, emb.sin() * self.attention_scaling

    def _rope_init(self, max_length: int) -> None:
        self.init_max_length = max_length
        cos, sin = self._angles(torch.arange(max_length, device=self.inv_freq.device))
        self.register_buffer("cos", cos, persistent=False)
        self.register_buffer("sin", sin, persistent=False)

    def forward(self, x, cu_q_lens=None, total_seq_lens=None, position_ids=None) -> Tuple[torch.Tensor, torch.Tensor]:
        return MojoOperator.forward(self)

    @staticmethod
    def _check_rotary_args(x, cu_q_lens, total_seq_lens, position_ids) -> None:
        """Reference ``position_embedding.py:60-67,82-83``."""
        for t in (cu_q_lens, total_seq_lens, position_ids):
            assert t is None or t.dtype == torch.int32
        assert position_ids is None or cu_q_lens is None, "At most one of cu_q_lens or position_ids should be provided"
        if cu_q_lens is not None:
            assert x.dim() == 2, "x must be 2D: [T, D]"
        elif position_ids is not None:
            assert position_ids.shape == x.shape[:-1], (
                "position_ids must have the same shape as x except the hidden dimension"
            )


class MojoApplyRoPE(MojoOperator):
    """Rotate-half (NeoX) RoPE on the LAST ``cos.shape[-1]`` features of q and k; the rest passes through.

    q/k: ``[T,N,D]`` / ``[N,T,D]`` or ``[B,S,N,D]`` / ``[B,N,S,D]`` (``head_first`` selects the second of
    each pair); cos/sin ``[T,d]``, ``[S,d]`` or ``[B,S,d]``.  Math runs in ``promote(q.dtype, cos.dtype)``.
    """

    def __init__(self, interleaved: bool = False):
        super().__init__()
        assert not interleaved, "interleaved impl is not supported yet."
        self.interleaved = interleaved

    def forward(self, q, k, cos, sin, head_first: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        return MojoOperator.forward(self)

    @staticmethod
    def _check_rope_args(q, k, cos, sin) -> None:
        """Reference ``position_embedding.py:163-167``."""
        assert q.ndim == k.ndim, "q and k must have the same dimension"
        assert q.ndim == 3 or q.ndim == 4, "q and k must be 3D or 4D"
        assert cos.shape == sin.shape, "cos and sin must have the same shape"
        if q.ndim == 3:
            assert cos.ndim == 2, (
                "rotary position embedding (cos/sin) must be of shape [num_tokens, rope_dim] "
                "for varlen prefill or decode"
            )

    def extra_repr(self) -> str:
        return f"interleaved={self.interleaved!r}"


class MojoGridRoPE(MojoOperator):
    """3-D grid RoPE of the DiT block (reference ``experimental/operators/position_embedding.py:80-118``):
    ``x [B, L, N, D]`` as interleaved complex pairs times a per-sample complex phase table ``freqs_list[b]
    [seq_len_b, 1, D/2]``; tokens past ``seq_len_b = F*H*W`` are passed through."""

    def forward(self, x: torch.Tensor, grid_sizes: torch.Tensor, freqs_list) -> torch.Tensor:
        return MojoOperator.forward(self)
