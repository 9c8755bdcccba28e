"""Paged-attention runtime state and CUDA-graph runners for the decoder hot path.

Mirrors the reference's ``PagedAttentionRuntimeState`` / ``AttentionMetadata``
(``mojo_opset/runtime/runtime.py:17-228``) and ``DeviceGraphRunner`` / ``DeviceGraphPool``
(``mojo_opset/compile/device_graph.py:8-105``) with the same method names and return values, B200-first:

* the block allocator runs ON THE DEVICE (``csrc/runtime.cu``): ``_reserve`` is one kernel instead of a Python loop
  with two ``.item()`` host syncs per sequence, and allocates the very same blocks in the very same order;
* the KV-store plan is not materialised: the store / fused pre-attention ops walk ``(block_tables, cu_q_lens,
  context_kv_lens)`` on the device, so ``AttentionMetadata.chunk_metadata`` stays ``None`` (the reference builds it
  with boolean-mask indexing = a dynamic shape = a host sync, ``core/operators/kv_cache.py:60-101``);
* a decode step therefore reads nothing on the host and is captured whole - bookkeeping, ops, GEMMs - in a
  ``torch.cuda.CUDAGraph`` (the reference goes through ``xpu_graph``'s tracing runner).
"""

from dataclasses import dataclass
from typing import Callable
from typing import Dict
from typing import List
from typing import Optional
from typing import Tuple

import torch

from . import _lib


@dataclass
class AttentionMetadata:
    q_lens: torch.Tensor
    cu_q_lens: Optional[torch.Tensor]
    total_seq_lens: torch.Tensor
    block_tables: torch.Tensor
    chunk_metadata: Optional[torch.Tensor]  # always None here: slots are found on the device
    key_caches: List[torch.Tensor]
    value_caches: List[torch.Tensor]
    is_prefill: bool
    context_kv_lens: torch.Tensor = None    # lengths before this step's append (the store ops' third argument)


class PagedAttentionRuntimeState:
    """Block tables, sequence lengths, the free-block stack and the per-layer caches of ``batch_size`` sequences."""

    def __init__(self, num_layers: int, num_kv_heads: int, head_dim: int, batch_size: int,
                 max_position_embeddings: int, device, dtype, block_size: int = 128):
        self.batch_size = batch_size
        self.num_layers = num_layers
        self.device = torch.device(device)
        self.dtype = dtype
        self.block_size = block_size
        self.num_kv_heads = num_kv_heads
        self.head_dim = head_dim
        self.max_blocks_per_seq = (max_position_embeddings + block_size - 1) // block_size
        total_blocks = batch_size * self.max_blocks_per_seq
        self.block_tables = torch.full((batch_size, self.max_blocks_per_seq), -1, dtype=torch.int32, device=device)
        self.total_seq_lens = torch.zeros((batch_size,), dtype=torch.int32, device=device)
        self.free_blocks = torch.arange(total_blocks, dtype=torch.int32, device=device)
        self._num_free = torch.full((1,), total_blocks, dtype=torch.int32, device=device)  # device-resident counter
        self._error = torch.zeros((1,), dtype=torch.int32, device=device)
        self._ones = torch.ones(batch_size, dtype=torch.int32, device=device)
        cache_shape = (total_blocks, num_kv_heads, block_size, head_dim)
        self.key_caches = [torch.zeros(cache_shape, dtype=dtype, device=device) for _ in range(num_layers)]
        self.value_caches = [torch.zeros(cache_shape, dtype=dtype, device=device) for _ in range(num_layers)]
        self._backup = None

    @classmethod
    def from_config(cls, model_config, batch_size: int, device, dtype, block_size: int = 128):
        """``model_config`` with the reference's field names (``runtime/runtime.py:39-45``)."""
        return cls(model_config.num_layers, getattr(model_config, "local_num_kv_heads", model_config.num_kv_heads),
                   model_config.head_dim, batch_size, model_config.max_position_embeddings, device, dtype, block_size)

    @property
    def kv_cache(self):
        return self

    @property
    def num_free_blocks(self) -> int:
        """Host read of the device counter (synchronises; for tests and diagnostics only)."""
        return int(self._num_free.item())

    def check(self) -> None:
        """Raise what the reference raises eagerly; a host read, so call it outside captured regions."""
        code = int(self._error.item())
        if code == 1:
            raise ValueError("PagedAttentionRuntimeState: Out of paged KV cache memory.")
        if code == 2:
            raise ValueError("PagedAttentionRuntimeState: a sequence exceeds max_position_embeddings.")

    # ---- state snapshots around graph capture (reference device_graph.py:52-66) -----------------------------------
    def backup_state(self) -> None:
        self._backup = (self.block_tables.clone(), self.total_seq_lens.clone(), self._num_free.clone())

    def restore_state(self) -> None:
        tables, lens, free = self._backup
        self.block_tables.copy_(tables)
        self.total_seq_lens.copy_(lens)
        self._num_free.copy_(free)

    # ---- bookkeeping on the device -----------------------------------------------------------------------------
    def _reserve(self, q_lens: Optional[torch.Tensor]) -> torch.Tensor:
        """Append ``q_lens`` tokens (``None`` = one each) to every sequence; returns the previous lengths."""
        lib = _lib.load()
        context = torch.empty_like(self.total_seq_lens)
        ql = None if q_lens is None else q_lens.to(device=self.device, dtype=torch.int32).contiguous()
        rc = lib.mojo_b200_paged_reserve(
            self.block_tables.data_ptr(), self.block_tables.stride(0), self.max_blocks_per_seq,
            self.total_seq_lens.data_ptr(), _lib.ptr(ql), self.free_blocks.data_ptr(), self._num_free.data_ptr(),
            context.data_ptr(), self.batch_size, self.block_size, self._error.data_ptr(),
            _lib.stream_ptr(self.device))
        _lib.check(lib, rc, "paged_reserve")
        return context

    def _metadata(self, q_lens, cu_q_lens, context) -> AttentionMetadata:
        return AttentionMetadata(q_lens=q_lens, cu_q_lens=cu_q_lens, total_seq_lens=self.total_seq_lens,
                                 block_tables=self.block_tables, chunk_metadata=None, key_caches=self.key_caches,
                                 value_caches=self.value_caches, is_prefill=cu_q_lens is not None,
                                 context_kv_lens=context)

    def prepare_prefill_inputs(self, input_ids: torch.Tensor, q_lens: torch.Tensor
                               ) -> Tuple[torch.Tensor, torch.Tensor, AttentionMetadata]:
        """``input_ids`` = the concatenated prompt tokens, ``q_lens[b]`` their count per sequence (the caller keeps
        ``input_ids.numel() == q_lens.sum()``; the reference checks it with a host read)."""
        lib = _lib.load()
        input_ids = input_ids.reshape(-1).to(device=self.device, dtype=torch.int64)
        q_lens = q_lens.to(device=self.device, dtype=torch.int32)
        context = self._reserve(q_lens)
        cu_q_lens = torch.nn.functional.pad(q_lens.cumsum(-1, dtype=torch.int32), (1, 0))
        positions = torch.empty(input_ids.numel(), dtype=torch.int64, device=self.device)
        rc = lib.mojo_b200_paged_positions(positions.data_ptr(), cu_q_lens.data_ptr(), context.data_ptr(),
                                           self.batch_size, positions.numel(), _lib.stream_ptr(self.device))
        _lib.check(lib, rc, "paged_positions")
        return input_ids, positions, self._metadata(q_lens, cu_q_lens, context)

    def prepare_decode_inputs(self, input_ids: torch.Tensor
                              ) -> Tuple[torch.Tensor, torch.Tensor, AttentionMetadata]:
        input_ids = input_ids.reshape(-1).to(device=self.device, dtype=torch.int64)
        if input_ids.numel() != self.batch_size:
            raise ValueError(
                f"Decode input_ids must provide exactly one token per sequence: {input_ids.numel()} != {self.batch_size}")
        context = self._reserve(None)
        return input_ids, context.to(torch.int64), self._metadata(self._ones, None, context)


class DeviceGraphRunner:
    """Capture ``fn(*static_inputs)`` once in a ``torch.cuda.CUDAGraph`` and replay it with new input values.

    ``fn`` may mutate long-lived state (the runtime state above, KV caches): state mutated during warm-up and capture
    is rolled back through ``session.backup_state() / restore_state()`` exactly as the reference's runner does."""

    def __init__(self, fn: Callable, warmup: int = 2):
        self.fn = fn
        self.warmup = warmup
        self.graph = None
        self.static_inputs = None
        self.static_outputs = None

    def capture(self, *inputs: torch.Tensor, session=None) -> None:
        self.static_inputs = [t.clone() for t in inputs]
        if session is not None and hasattr(session, "backup_state"):
            session.backup_state()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.inference_mode():
            for _ in range(self.warmup):
                self.fn(*self.static_inputs)
                if session is not None and hasattr(session, "restore_state"):
                    session.restore_state()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.static_outputs = self.fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        if session is not None and hasattr(session, "restore_state"):
            session.restore_state()

    def replay(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs


class DeviceGraphPool:
    """Batch-size-keyed cache of runners bound to one session (reference ``device_graph.py:73-105``)."""

    def __init__(self, fn: Callable):
        self._fn = fn
        self._runners: Dict[int, DeviceGraphRunner] = {}
        self._bound_session_id = None

    def get_runner(self, input_ids: torch.Tensor, session) -> DeviceGraphRunner:
        if id(session) != self._bound_session_id:
            self._runners.clear()
            self._bound_session_id = id(session)
        bs = input_ids.shape[0]
        if bs not in self._runners:
            runner = DeviceGraphRunner(self._fn)
            runner.capture(input_ids, session=session)
            self._runners[bs] = runner
        return self._runners[bs]

    @property
    def captured_batch_sizes(self) -> List[int]:
        return sorted(self._runners.keys())
