"""The oracle restatement vs the vectors the unmodified reference produced (tests/golden/make_golden.py).

Bit-exact for every op whose arithmetic the oracle restates; 1e-2 for SDPA whose CPU flash kernel is ATen's.
"""

import pytest
import torch

from conftest import load_golden
from oracle import golden


def _ids(cases):
    return [c["name"] for c in cases]


DECODE = load_golden("paged_decode_gqa.pt")
PREFILL = load_golden("paged_prefill_gqa.pt")
SDPA = load_golden("sdpa.pt")
STORE = load_golden("store_paged_kv.pt")
NORM = load_golden("rmsnorm.pt")
ROPE = load_golden("apply_rope.pt")
ROTARY = load_golden("rotary_embedding.pt")
ACT = load_golden("activation.pt")
SWA = load_golden("paged_swa.pt")


@pytest.mark.parametrize("case", DECODE, ids=_ids(DECODE))
def test_decode_bit_exact(case):
    out = golden.paged_decode_gqa(case["query"], case["key_cache"], case["value_cache"], case["total_seq_lens"],
                                  case["block_tables"], case["softmax_scale"], case["gqa_layout"])
    assert torch.equal(out, case["out"])


@pytest.mark.parametrize("case", PREFILL, ids=_ids(PREFILL))
def test_prefill_bit_exact(case):
    out = golden.paged_prefill_gqa(case["query"], case["key_cache"], case["value_cache"], case["cu_q_lens"],
                                   case["block_tables"], case["softmax_scale"], case["cu_total_seq_lens"],
                                   case["gqa_layout"])
    assert torch.equal(out, case["out"])


@pytest.mark.parametrize("case", SWA, ids=_ids(SWA))
def test_swa_bit_exact(case):
    """MojoPagedPrefillSWA / MojoPagedDecodeSWA (reference attention.py:533-745): the restatement against the
    reference's own outputs."""
    win = dict(local_window_size=case["local_window_size"], global_window_size=case["global_window_size"])
    if case["op"] == "prefill":
        out = golden.paged_prefill_swa(case["query"], case["key_cache"], case["value_cache"], case["cu_q_lens"],
                                       case["block_table"], case["softmax_scale"], case["cu_total_seq_lens"],
                                       case["gqa_layout"], True, **win)
    else:
        out = golden.paged_decode_swa(case["query"], case["key_cache"], case["value_cache"], case["total_seq_lens"],
                                      case["block_table"], case["softmax_scale"], case["gqa_layout"], **win)
    assert torch.equal(out, case["out"])


def test_window_mask_matches_definition():
    """Row at position p sees key k iff k <= p and (p <= k + local or k < global) - brute force."""
    for q_len, kv_len, local, glob in [(5, 9, 2, None), (7, 7, None, 3), (4, 20, 6, 2), (3, 3, None, None), (6, 10, 0, 0)]:
        m = golden.window_mask(q_len, kv_len, local, glob)
        for t in range(q_len):
            p = kv_len - q_len + t
            for k in range(kv_len):
                win = True if local is None and glob is None else (
                    (local is not None and p <= k + local) or (glob is not None and k < glob))
                assert bool(m[t, k]) == (k <= p and win)


@pytest.mark.parametrize("case", SDPA, ids=_ids(SDPA))
def test_sdpa(case):
    out = golden.sdpa(case["query"], case["key"], case["value"], case["attn_mask"], case["scale"], case["enable_gqa"])
    torch.testing.assert_close(out.float(), case["out"].float(), atol=1e-2, rtol=1e-2)
    aten = golden.sdpa_aten(case["query"], case["key"], case["value"], case["attn_mask"], case["scale"],
                            case["enable_gqa"])
    assert torch.equal(aten, case["out"])


@pytest.mark.parametrize("case", STORE, ids=_ids(STORE))
def test_store_kv_bit_exact(case):
    bs = case["key_cache"].shape[2]
    plan = golden.build_chunk_plan(case["block_table"], case["cu_q_lens"], case["context_kv_lens"], bs)
    assert torch.equal(plan, case["chunk_metadata"])
    kc, vc = golden.store_paged_kv(case["key_states"], case["value_states"], case["key_cache"].clone(),
                                   case["value_cache"].clone(), plan)
    assert torch.equal(kc, case["key_cache_out"]) and torch.equal(vc, case["value_cache_out"])


@pytest.mark.parametrize("case", STORE, ids=_ids(STORE))
def test_product_plan_builder_matches_reference(case):
    from mojo_opset_b200.core import build_paged_kv_chunk_metadata

    bs = case["key_cache"].shape[2]
    plan = build_paged_kv_chunk_metadata(case["block_table"], case["cu_q_lens"], case["context_kv_lens"], bs)
    assert plan.dtype == torch.int32 and torch.equal(plan, case["chunk_metadata"])


@pytest.mark.parametrize("case", NORM, ids=_ids(NORM))
def test_rmsnorm_bit_exact(case):
    assert torch.equal(golden.rms_norm(case["hidden_state"], case["weight"], case["eps"]), case["rmsnorm_out"])
    if case["norm_pos"] is not None:
        y, r = golden.residual_add_rms_norm(case["hidden_state"], case["residual"], case["weight"], case["eps"],
                                            case["norm_pos"])
        assert torch.equal(y, case["out"]) and torch.equal(r, case["residual_out"])


@pytest.mark.parametrize("case", ROPE, ids=_ids(ROPE))
def test_rope_bit_exact(case):
    q, k = golden.apply_rope(case["q"], case["k"], case["cos"], case["sin"], case["head_first"])
    assert torch.equal(q, case["q_out"]) and torch.equal(k, case["k_out"])


@pytest.mark.parametrize("case", ROTARY, ids=_ids(ROTARY))
def test_rotary_embedding_bit_exact(case):
    x = torch.empty(*case["x_shape"])
    pos = golden.rotary_positions(x, case["cu_q_lens"], case["total_seq_lens"], case["position_ids"])
    inv_freq = 1.0 / (case["rope_theta"] ** (torch.arange(0, case["rope_dim"], 2, dtype=torch.float32) / case["rope_dim"]))
    cos, sin = golden.rotary_cos_sin(pos, inv_freq)
    assert torch.equal(cos, case["cos"]) and torch.equal(sin, case["sin"])


@pytest.mark.parametrize("case", ACT, ids=_ids(ACT))
def test_activation_bit_exact(case):
    assert torch.equal(golden.swiglu(case["gate"], case["up"], case["swiglu_limit"]), case["out"])
    assert torch.equal(golden.silu(case["gate"]), case["silu_out"])


@pytest.mark.parametrize("case", STORE, ids=_ids(STORE))
def test_c_oracle_store_kv_bit_exact(case):
    """The plain-C restatement (oracle/store_kv.c) against the reference's vectors."""
    from oracle import store_kv_c

    bs = case["key_cache"].shape[2]
    plan = store_kv_c.build_chunk_plan(case["block_table"], case["cu_q_lens"], case["context_kv_lens"], bs)
    assert torch.equal(plan, case["chunk_metadata"])
    kc = store_kv_c.store_paged_kv(case["key_states"].contiguous(), case["key_cache"].clone(), plan)
    vc = store_kv_c.store_paged_kv(case["value_states"].contiguous(), case["value_cache"].clone(), plan)
    assert torch.equal(kc, case["key_cache_out"]) and torch.equal(vc, case["value_cache_out"])


RUNTIME = load_golden("runtime.pt")


@pytest.mark.parametrize("case", RUNTIME, ids=[c["name"] for c in RUNTIME])
def test_runtime_oracle_matches_reference_bookkeeping(case):
    """oracle/runtime_ref.py (the checker of the device-side block allocator) against the state the UNMODIFIED
    reference's PagedAttentionRuntimeState went through (runtime/runtime.py:112-228): block tables, lengths, free
    count, positions and the KV-store plan after every prefill / decode step - bit exact."""
    from oracle import golden
    from oracle.runtime_ref import ReserveOracle

    o = ReserveOracle(case["batch"], case["max_position_embeddings"], case["block_size"])
    for step in case["trace"]:
        q = step["q_lens"]
        ctx = o.reserve(q)
        assert torch.equal(o.block_tables, step["block_tables"])
        assert torch.equal(o.total_seq_lens, step["total_seq_lens"])
        assert o.num_free_blocks == step["num_free_blocks"]
        if step["kind"] == "prefill":
            assert torch.equal(o.positions(ctx, q), step["positions"])
            cu = torch.nn.functional.pad(q.cumsum(-1, dtype=torch.int32), (1, 0))
            assert torch.equal(cu, step["cu_q_lens"])
            plan = golden.build_chunk_plan(o.block_tables, cu, ctx, case["block_size"])
        else:
            assert torch.equal(ctx.to(torch.int64), step["positions"].to(torch.int64))
            plan = golden.build_chunk_plan(o.block_tables, None, ctx, case["block_size"])
        assert torch.equal(plan, step["chunk_metadata"])


SWA_NONPAGED = load_golden("swa.pt")


@pytest.mark.parametrize("case", SWA_NONPAGED, ids=[c["name"] for c in SWA_NONPAGED])
def test_nonpaged_swa_oracle_matches_reference(case):
    """oracle.golden.swa against the UNMODIFIED reference's MojoSWA outputs (tests/golden/make_swa_golden.py): bit exact."""
    from oracle import golden

    out = golden.swa(case["query"], case["key"], case["value"], case["cu_q_lens"], case["cu_total_seq_lens"], None,
                     case["gqa_layout"], True, case["local_window_size"], case["global_window_size"])
    assert torch.equal(out, case["out"])
