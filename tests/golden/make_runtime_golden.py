#!/usr/bin/env python
"""Generates tests/golden/runtime.pt by running the UNMODIFIED reference's paged-KV bookkeeping
(``mojo_opset/runtime/runtime.py:29-228``: ``PagedAttentionRuntimeState._reserve / _allocate_blocks /
_build_positions / prepare_prefill_inputs / prepare_decode_inputs``) on CPU through a scripted sequence of ragged
prefill and decode steps.  The fixture pins ``oracle/runtime_ref.py`` (tests/test_oracle_golden.py) - and through it
the device-side allocator (tests/test_gpu_runtime.py).

    PYTHONPATH=/root/reference python tests/golden/make_runtime_golden.py
"""
import os
import sys
from types import SimpleNamespace

import torch

os.environ["MOJO_BACKEND"] = "torch"
os.environ["MOJO_OPSET_PLUGIN_AUTOLOAD"] = "0"
from mojo_opset.runtime.runtime import PagedAttentionRuntimeState  # noqa: E402  (the reference)

HERE = os.path.dirname(os.path.abspath(__file__))


def scenario(name, batch, max_pos, block_size, steps):
    cfg = SimpleNamespace(model_config=SimpleNamespace(num_layers=1, num_kv_heads=1, head_dim=8,
                                                       max_position_embeddings=max_pos))
    state = PagedAttentionRuntimeState(cfg, batch, device="cpu", dtype=torch.float32, block_size=block_size)
    trace = []
    for kind, q_lens in steps:
        if kind == "prefill":
            q = torch.tensor(q_lens, dtype=torch.int32)
            ids = torch.zeros(int(q.sum()), dtype=torch.int64)
            _, positions, meta = state.prepare_prefill_inputs(ids, q)
        else:
            q = torch.ones(batch, dtype=torch.int32)
            _, positions, meta = state.prepare_decode_inputs(torch.zeros(batch, dtype=torch.int64))
        trace.append(dict(kind=kind, q_lens=q.clone(), positions=positions.clone(),
                          block_tables=state.block_tables.clone(), total_seq_lens=state.total_seq_lens.clone(),
                          num_free_blocks=int(state.num_free_blocks), chunk_metadata=meta.chunk_metadata.clone(),
                          cu_q_lens=None if meta.cu_q_lens is None else meta.cu_q_lens.clone()))
    return dict(name=name, batch=batch, max_position_embeddings=max_pos, block_size=block_size, trace=trace)


def main():
    cases = [
        scenario("ragged_prefill_then_decode", 4, 256, 16,
                 [("prefill", [40, 0, 17, 100])] + [("decode", None)] * 20 + [("prefill", [3, 50, 0, 1])]
                 + [("decode", None)] * 5),
        scenario("chunked_prefill_bs128", 3, 1024, 128,
                 [("prefill", [300, 128, 1]), ("prefill", [212, 0, 127]), ("decode", None), ("decode", None),
                  ("prefill", [0, 512, 0])]),
        scenario("page_boundaries_bs8", 5, 64, 8,
                 [("prefill", [8, 7, 9, 16, 1])] + [("decode", None)] * 17),
    ]
    torch.save(cases, os.path.join(HERE, "runtime.pt"))
    print("runtime.pt:", [(c["name"], len(c["trace"])) for c in cases])


if __name__ == "__main__":
    main()
