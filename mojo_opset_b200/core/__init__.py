"""Core op interfaces of the paged-attention decoder hot path (the subset of the reference's
``mojo_opset/core/__init__.py`` that ``BASELINE.json:north_star`` names)."""

from .backend_registry import MojoBackendRegistry
from .operator import MojoOperator
from .operators.activation import MojoGelu
from .operators.activation import MojoSilu
from .operators.activation import MojoSwiGLU
from .operators.attention import MojoPagedDecodeGQA
from .operators.attention import MojoPagedDecodeSWA
from .operators.attention import MojoPagedPrefillGQA
from .operators.attention import MojoPagedPrefillSWA
from .operators.attention import MojoSWA
from .operators.attention import MojoSdpa
from .operators.compute_with_comm import MojoGemmAllReduce
from .operators.fused_attention_input import MojoNormRoPEStoreKV
from .operators.fused_attention_input import MojoRoPEStoreKV
from .operators.kv_cache import MojoStorePagedKVCache
from .operators.kv_cache import build_paged_kv_chunk_metadata
from .operators.normalization import MojoLayerNorm
from .operators.normalization import MojoResidualAddRMSNorm
from .operators.normalization import MojoRMSNorm
from .operators.position_embedding import MojoApplyRoPE
from .operators.position_embedding import MojoGridRoPE
from .operators.position_embedding import MojoRotaryEmbedding

__all__ = [
    "MojoBackendRegistry",
    "MojoOperator",
    "MojoGelu",
    "MojoSilu",
    "MojoSwiGLU",
    "MojoPagedDecodeGQA",
    "MojoPagedPrefillGQA",
    "MojoPagedPrefillSWA",
    "MojoSWA",
    "MojoPagedDecodeSWA",
    "MojoSdpa",
    "MojoGemmAllReduce",
    "MojoNormRoPEStoreKV",
    "MojoRoPEStoreKV",
    "MojoStorePagedKVCache",
    "build_paged_kv_chunk_metadata",
    "MojoLayerNorm",
    "MojoResidualAddRMSNorm",
    "MojoRMSNorm",
    "MojoApplyRoPE",
    "MojoGridRoPE",
    "MojoRotaryEmbedding",
]
