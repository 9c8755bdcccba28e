"""TEST INFRASTRUCTURE - CPU restatement of the reference's paged-KV bookkeeping
(``mojo_opset/runtime/runtime.py:112-158``: ``_allocate_blocks``, ``_reserve``, ``_build_positions``), used by
``tests/test_gpu_runtime.py`` to check that the device-side allocator hands out the same blocks in the same order.
The product never imports it.

Parity pin: ``tests/golden/runtime.pt`` (made by ``tests/golden/make_runtime_golden.py`` from the UNMODIFIED reference class) -
``tests/test_oracle_golden.py::test_runtime_oracle_matches_reference_bookkeeping`` checks this file against it bit for bit."""

import torch


class ReserveOracle:
    def __init__(self, batch_size: int, max_position_embeddings: int, block_size: int):
        self.batch_size, self.block_size = batch_size, block_size
        self.max_blocks_per_seq = (max_position_embeddings + block_size - 1) // block_size
        total_blocks = batch_size * self.max_blocks_per_seq
        self.block_tables = torch.full((batch_size, self.max_blocks_per_seq), -1, dtype=torch.int32)
        self.total_seq_lens = torch.zeros((batch_size,), dtype=torch.int32)
        self.free_blocks = torch.arange(total_blocks, dtype=torch.int32)
        self.num_free_blocks = total_blocks

    def _allocate_blocks(self, num_blocks: int) -> torch.Tensor:  # runtime.py:112-117
        if num_blocks > self.num_free_blocks:
            raise ValueError("PagedAttentionRuntimeState: Out of paged KV cache memory.")
        allocated = self.free_blocks[self.num_free_blocks - num_blocks: self.num_free_blocks]
        self.num_free_blocks -= num_blocks
        return allocated

    def reserve(self, q_lens: torch.Tensor) -> torch.Tensor:  # runtime.py:124-141
        previous = self.total_seq_lens.clone()
        for b in range(self.batch_size):
            context_len, append_len = int(previous[b]), int(q_lens[b])
            old_n = (context_len + self.block_size - 1) // self.block_size
            new_n = (context_len + append_len + self.block_size - 1) // self.block_size
            if new_n > old_n:
                self.block_tables[b, old_n:new_n] = self._allocate_blocks(new_n - old_n)
        self.total_seq_lens = previous + q_lens.to(torch.int32)
        return previous

    def positions(self, context_kv_lens: torch.Tensor, q_lens: torch.Tensor) -> torch.Tensor:  # runtime.py:143-154
        out = [torch.arange(int(c), int(c) + int(q), dtype=torch.int64) for c, q in zip(context_kv_lens, q_lens) if q > 0]
        return torch.cat(out) if out else torch.empty((0,), dtype=torch.int64)
