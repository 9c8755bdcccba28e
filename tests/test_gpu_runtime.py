"""-m gpu: the device-side paged-KV runtime (csrc/runtime.cu + mojo_opset_b200/runtime.py) against the CPU
restatement of the reference's host-side bookkeeping, and the Qwen3-shaped demo (examples/qwen3_synthetic.py)
- eager prefill, then decode steps replayed from ONE CUDA graph - against the same network evaluated with the
oracle's ops."""

import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_reserve_matches_reference_order():
    from mojo_opset_b200.runtime import PagedAttentionRuntimeState
    from oracle.runtime_ref import ReserveOracle

    g = torch.Generator().manual_seed(0)
    for batch, bs, max_pos in ((5, 16, 400), (1500, 16, 64), (64, 128, 4096 + 300)):
        state = PagedAttentionRuntimeState(1, 1, 64, batch, max_pos, DEV, torch.bfloat16, block_size=bs)
        ref = ReserveOracle(batch, max_pos, bs)
        # ragged prefill, then decode steps across block boundaries, then a second prefill chunk
        steps = [torch.randint(0, max_pos // 3, (batch,), generator=g, dtype=torch.int32)]
        steps += [None] * (bs + 3)
        steps += [torch.randint(0, max_pos // 4, (batch,), generator=g, dtype=torch.int32)]
        for q in steps:
            if q is None:
                ids, pos, meta = state.prepare_decode_inputs(torch.zeros(batch, dtype=torch.int64, device=DEV))
                q_ref = torch.ones(batch, dtype=torch.int32)
            else:
                n = int(q.sum())
                ids, pos, meta = state.prepare_prefill_inputs(torch.zeros(n, dtype=torch.int64), q)
                q_ref = q
            ctx_ref = ref.reserve(q_ref)
            assert torch.equal(meta.context_kv_lens.cpu(), ctx_ref)
            assert torch.equal(pos.cpu(), ref.positions(ctx_ref, q_ref))
            assert torch.equal(state.block_tables.cpu(), ref.block_tables)
            assert torch.equal(state.total_seq_lens.cpu(), ref.total_seq_lens)
            assert state.num_free_blocks == ref.num_free_blocks
        state.check()


def test_device_allocator_replays_the_reference_trace():
    """The device-side allocator against the states the UNMODIFIED reference's PagedAttentionRuntimeState went through
    (tests/golden/runtime.pt, made by tests/golden/make_runtime_golden.py): same blocks, same order, same positions."""
    from conftest import load_golden
    from mojo_opset_b200.runtime import PagedAttentionRuntimeState

    for case in load_golden("runtime.pt"):
        state = PagedAttentionRuntimeState(1, 1, 8, case["batch"], case["max_position_embeddings"], DEV,
                                           torch.bfloat16, block_size=case["block_size"])
        for step in case["trace"]:
            if step["kind"] == "prefill":
                q = step["q_lens"]
                _, pos, meta = state.prepare_prefill_inputs(torch.zeros(int(q.sum()), dtype=torch.int64), q)
                assert torch.equal(meta.cu_q_lens.cpu(), step["cu_q_lens"])
            else:
                _, pos, meta = state.prepare_decode_inputs(torch.zeros(case["batch"], dtype=torch.int64, device=DEV))
            assert torch.equal(pos.cpu(), step["positions"].to(torch.int64)), case["name"]
            assert torch.equal(state.block_tables.cpu(), step["block_tables"]), case["name"]
            assert torch.equal(state.total_seq_lens.cpu(), step["total_seq_lens"])
            assert state.num_free_blocks == step["num_free_blocks"]
        state.check()


def test_reserve_out_of_memory_changes_nothing():
    from mojo_opset_b200.runtime import PagedAttentionRuntimeState

    state = PagedAttentionRuntimeState(1, 1, 64, 4, 64, DEV, torch.bfloat16, block_size=16)
    state.prepare_prefill_inputs(torch.zeros(40, dtype=torch.int64), torch.tensor([10, 10, 10, 10], dtype=torch.int32))
    tables, lens, free = state.block_tables.clone(), state.total_seq_lens.clone(), state.num_free_blocks
    state._reserve(torch.tensor([100, 1, 1, 1], dtype=torch.int32))  # would exceed max_blocks_per_seq
    assert torch.equal(state.block_tables, tables) and torch.equal(state.total_seq_lens, lens)
    assert state.num_free_blocks == free
    with pytest.raises(ValueError):
        state.check()


def test_qwen3_demo_graph_decode_matches_oracle():
    os.environ["MOJO_BACKEND"] = "b200"
    from qwen3_synthetic import Qwen3Config
    from qwen3_synthetic import Qwen3Synthetic
    from qwen3_synthetic import reference_forward

    from mojo_opset_b200.runtime import DeviceGraphRunner
    from mojo_opset_b200.runtime import PagedAttentionRuntimeState
    from oracle import golden
    from oracle.runtime_ref import ReserveOracle

    cfg = Qwen3Config(hidden_size=256, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=64, intermediate_size=512,
                      vocab_size=1000, max_position_embeddings=128)
    model = Qwen3Synthetic(cfg, DEV, torch.bfloat16, seed=3)
    B, bs = 3, 16
    state = PagedAttentionRuntimeState.from_config(cfg, B, DEV, torch.bfloat16, block_size=bs)
    ref_state = ReserveOracle(B, cfg.max_position_embeddings, bs)
    shape = state.key_caches[0].shape
    ref_k = [torch.zeros(shape, dtype=torch.bfloat16) for _ in range(cfg.num_layers)]
    ref_v = [torch.zeros(shape, dtype=torch.bfloat16) for _ in range(cfg.num_layers)]
    g = torch.Generator().manual_seed(1)
    q_lens = torch.tensor([20, 5, 33], dtype=torch.int32)
    prompt = torch.randint(0, cfg.vocab_size, (int(q_lens.sum()),), generator=g)

    def compare(logits, ref_logits):
        # bf16 network, fp32-accumulating GEMMs on both sides: logits agree to bf16 resolution of O(1) values
        torch.testing.assert_close(logits.float().cpu(), ref_logits.float(), atol=6e-2, rtol=6e-2)

    # ---- prefill (eager)
    ids, pos, meta = state.prepare_prefill_inputs(prompt, q_lens)
    logits = model(ids, pos, meta)
    ctx = ref_state.reserve(q_lens)
    ref_logits = reference_forward(model, golden, prompt, ref_state.positions(ctx, q_lens), q_lens, ctx,
                                   ref_state.block_tables, ref_k, ref_v, True)
    compare(logits, ref_logits)
    last = (q_lens.cumsum(0) - 1).long()
    next_ids = ref_logits[last].argmax(-1)  # both sides continue from the oracle's tokens

    # ---- decode: one captured graph, replayed
    def step(input_ids):
        i, p, m = state.prepare_decode_inputs(input_ids)
        return model(i, p, m)

    runner = DeviceGraphRunner(step)
    runner.capture(next_ids.to(DEV), session=state)
    ones = torch.ones(B, dtype=torch.int32)
    for _ in range(bs + 2):  # crosses a page boundary: the captured allocator must hand out new blocks
        logits = runner.replay(next_ids.to(DEV)).clone()
        ctx = ref_state.reserve(ones)
        ref_logits = reference_forward(model, golden, next_ids, ctx.long(), ones, ctx, ref_state.block_tables, ref_k,
                                       ref_v, False)
        compare(logits, ref_logits)
        assert torch.equal(state.block_tables.cpu(), ref_state.block_tables)
        next_ids = ref_logits.argmax(-1)
    state.check()
