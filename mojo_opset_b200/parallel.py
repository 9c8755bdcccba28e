"""Multi-GPU plumbing of the hot path: one process per GPU (torchrun), ``torch.distributed`` for the plumbing.

The path shards without a data-path collective:

* data parallel (cfg2 decode, cfg3 prefill, cfg5 DiT): sequences / batch elements are independent; each rank owns
  its sequences' pages and block tables -> ``shard_range``;
* tensor parallel over KV-head groups (cfg4, Llama-70B-shaped decode): rank r runs the unchanged attention ops on
  q heads ``[r*Hq/tp, (r+1)*Hq/tp)`` and its KV heads - the reference's rule for paged attention is q ``Shard(-2)``,
  caches ``Shard(-3)``, out ``Shard(-2)`` (``mojo_opset/distributed/parallel/partitions.py:85-89``) with the QKV
  weight split of ``partitions.py:123-177`` (KV heads are REPLICATED over ``tp / Hkv`` ranks when ``tp > Hkv``) ->
  ``shard_heads``.  The only collective on the path is the all-reduce of the row-parallel ``o_proj`` output
  (``Partial() -> Replicate()``, ``partitions.py:42-47``, ``mojo_parallel.py:121-127``) -> ``RowParallelOutProj``.

Backend: NCCL over NVLink5/NVSwitch on B200, gloo on CPU (tests) - ``utils.platform.get_dist_backend``.
"""

from typing import NamedTuple
from typing import Optional
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) share of ``total`` independent units (sequences, batch elements): the first
    ``total % world`` ranks take one extra."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of world {world}")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class HeadShard(NamedTuple):
    q_begin: int
    q_end: int
    kv_begin: int
    kv_end: int
    kv_replicas: int  # ranks holding the same KV heads (1 unless tp > Hkv)


def shard_heads(num_q_heads: int, num_kv_heads: int, tp: int, rank: int) -> HeadShard:
    """AABB head partition of reference ``partitions.py:123-177``: q heads split evenly; KV heads split when
    ``Hkv >= tp``, otherwise each KV head is shared by ``tp / Hkv`` consecutive ranks."""
    if num_q_heads % tp:
        raise ValueError(f"num_q_heads ({num_q_heads}) must be divisible by tp ({tp})")
    if not ((num_kv_heads >= tp and num_kv_heads % tp == 0) or (tp > num_kv_heads and tp % num_kv_heads == 0)):
        raise ValueError(f"num_kv_heads ({num_kv_heads}) and tp ({tp}) must divide one another")
    if (num_q_heads // num_kv_heads) % max(1, tp // num_kv_heads):
        raise ValueError("a rank's q heads must belong to one KV head when KV heads are replicated")
    q_per = num_q_heads // tp
    kv_per = max(1, num_kv_heads // tp)
    replicas = max(1, tp // num_kv_heads)
    kv_begin = (rank // replicas) * kv_per
    return HeadShard(rank * q_per, (rank + 1) * q_per, kv_begin, kv_begin + kv_per, replicas)


def all_reduce_sum_(t: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """In-place sum over the TP group on the tensor's current stream (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class RowParallelOutProj(torch.nn.Module):
    """``o_proj`` of a TP attention block: ``y = all_reduce(x_local @ W[:, q_begin*D : q_end*D].T)``.

    The GEMM is a plain library GEMM (cuBLAS through ``torch.nn.functional.linear`` - not part of the hand-written
    path); what belongs to the path is the sharding rule and the single all-reduce behind it (reference
    ``RowwiseParallel`` Linear: input ``Shard(-1)``, output ``Partial()`` -> ``Replicate()``)."""

    def __init__(self, full_weight: torch.Tensor, shard: HeadShard, head_dim: int,
                 group: Optional[dist.ProcessGroup] = None):
        super().__init__()
        cols = slice(shard.q_begin * head_dim, shard.q_end * head_dim)
        self.weight = torch.nn.Parameter(full_weight[:, cols].contiguous(), requires_grad=False)
        self.group = group

    def forward(self, attn_out_local: torch.Tensor) -> torch.Tensor:
        x = attn_out_local.reshape(attn_out_local.shape[0], -1)
        return all_reduce_sum_(torch.nn.functional.linear(x, self.weight), self.group)
