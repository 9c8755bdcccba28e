"""``b200`` backend: ``B200<Op>`` classes subclassing the core ops (name prefix = backend key)."""

from .operators.activation import B200Gelu
from .operators.activation import B200Silu
from .operators.activation import B200SwiGLU
from .operators.attention import B200PagedDecodeGQA
from .operators.attention import B200PagedDecodeSWA
from .operators.attention import B200PagedPrefillGQA
from .operators.attention import B200PagedPrefillSWA
from .operators.attention import B200SWA
from .operators.attention import B200Sdpa
from .operators.compute_with_comm import B200GemmAllReduce
from .operators.fused_attention_input import B200NormRoPEStoreKV
from .operators.fused_attention_input import B200RoPEStoreKV
from .operators.kv_cache import B200StorePagedKVCache
from .operators.normalization import B200LayerNorm
from .operators.normalization import B200ResidualAddRMSNorm
from .operators.normalization import B200RMSNorm
from .operators.position_embedding import B200ApplyRoPE
from .operators.position_embedding import B200GridRoPE
from .operators.position_embedding import B200RotaryEmbedding

__all__ = [
    "B200Gelu",
    "B200Silu",
    "B200SwiGLU",
    "B200PagedDecodeGQA",
    "B200PagedPrefillGQA",
    "B200PagedPrefillSWA",
    "B200SWA",
    "B200PagedDecodeSWA",
    "B200Sdpa",
    "B200GemmAllReduce",
    "B200NormRoPEStoreKV",
    "B200RoPEStoreKV",
    "B200StorePagedKVCache",
    "B200LayerNorm",
    "B200ResidualAddRMSNorm",
    "B200RMSNorm",
    "B200ApplyRoPE",
    "B200GridRoPE",
    "B200RotaryEmbedding",
]
