"""-m gpu: the CUDA path (through the op classes -> C ABI) against the vectors the unmodified reference
produced.  KV store bit-exact; floating point ops within the tolerances north_star states (bf16/fp16
atol=rtol=2e-2; fp32 1e-5/1e-5 for decode) and, where the kernel reproduces the golden rounding points,
much tighter bounds that are asserted as well.
"""

import os

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
SWA = load_golden("paged_swa.pt")


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


@pytest.mark.parametrize("case", SWA, ids=_ids(SWA))
def test_paged_swa(ops, case):
    """MojoPagedPrefillSWA / MojoPagedDecodeSWA against the reference's outputs, through the op classes -> C ABI."""
    win = dict(global_window_size=case["global_window_size"], local_window_size=case["local_window_size"])
    if case["op"] == "prefill":
        op = ops.MojoPagedPrefillSWA(is_causal=True, gqa_layout=case["gqa_layout"], **win)
        assert type(op).__name__ == "B200PagedPrefillSWA"
        out = op(_cuda(case["query"]), _cuda(case["key_cache"]), _cuda(case["value_cache"]), _cuda(case["cu_q_lens"]),
                 _cuda(case["block_table"]), softmax_scale=case["softmax_scale"],
                 cu_total_seq_lens=_cuda(case["cu_total_seq_lens"]))
    else:
        op = ops.MojoPagedDecodeSWA(is_causal=True, gqa_layout=case["gqa_layout"], **win)
        assert type(op).__name__ == "B200PagedDecodeSWA"
        out = op(_cuda(case["query"]), _cuda(case["key_cache"]), _cuda(case["value_cache"]),
                 _cuda(case["total_seq_lens"]), _cuda(case["block_table"]), softmax_scale=case["softmax_scale"])
    assert out.shape == case["out"].shape
    torch.testing.assert_close(out.cpu().float(), case["out"].float(), atol=2e-2, rtol=2e-2)


def test_paged_swa_long_window_skips_tiles(ops):
    """A long sequence with a short window: the kernel loads only the KV tiles a query block can see (global prefix +
    the tiles under the window); the result must match the oracle, and - property - must not depend on the keys and
    values outside every window."""
    from oracle import golden

    g = torch.Generator().manual_seed(77)
    Hq, Hkv, D, bs, T, local, glob = 4, 2, 128, 16, 1500, 100, 40
    nb = T // bs + 3
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    q = torch.randn(T, Hq, D, generator=g).to(torch.bfloat16)
    table = torch.randperm(nb, generator=g)[: (T + bs - 1) // bs].view(1, -1).to(torch.int32)
    cu = torch.tensor([0, T], dtype=torch.int32)
    op = ops.MojoPagedPrefillSWA(global_window_size=glob, local_window_size=local)
    out = op(_cuda(q), _cuda(kc), _cuda(vc), _cuda(cu), _cuda(table))
    ref = golden.paged_prefill_swa(q, kc, vc, cu, table, None, None, "AABB", True, local, glob)
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=2e-2, rtol=2e-2)
    # rows >= 1000 see keys < 40 and keys >= 900 only: rewriting keys 64..799 cannot change them (bit-exact)
    kc2, vc2 = kc.clone(), vc.clone()
    mid = table[0, 4:50].long()
    kc2[mid] = torch.randn(kc2[mid].shape, generator=g).to(torch.bfloat16)
    vc2[mid] = torch.randn(vc2[mid].shape, generator=g).to(torch.bfloat16)
    out2 = op(_cuda(q), _cuda(kc2), _cuda(vc2), _cuda(cu), _cuda(table))
    assert torch.equal(out2[1000:], out[1000:])
    assert not torch.equal(out2[:900], out[:900])


@pytest.mark.parametrize("local,glob", [(100, None), (700, 130), (None, 200), (5000, 5000), (0, 1), (63, 64)])
@pytest.mark.parametrize("splits", [0, 3])
def test_paged_decode_swa_long(ops, local, glob, splits):
    """MojoPagedDecodeSWA on the split-KV decode kernel restricted to the visible KV tiles: contexts of many tiles,
    windows that leave a gap / touch / cover everything, forced split counts, empty and one-token rows - against the
    oracle; and the keys outside both windows cannot matter (bit-exact after rewriting them)."""
    from oracle import golden

    g = torch.Generator().manual_seed(1000 + (local or 0) + 7 * (glob or 0))
    Hq, Hkv, D, bs = 8, 2, 128, 16
    lens = [2900, 1, 0, 640, 1337, 64]
    need = [(n + bs - 1) // bs for n in lens]
    nb = sum(need) + 3
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    table = torch.full((len(lens), max(need)), -1, dtype=torch.int32)
    perm = torch.randperm(nb, generator=g).to(torch.int32)
    at = 0
    for i, n in enumerate(need):
        table[i, :n] = perm[at:at + n]
        at += n
    q = torch.randn(len(lens), Hq, D, generator=g).to(torch.bfloat16)
    seq = torch.tensor(lens, dtype=torch.int32)
    op = ops.MojoPagedDecodeSWA(global_window_size=glob, local_window_size=local)
    if splits:
        os.environ["MOJO_B200_DECODE_SPLITS"] = str(splits)
    try:
        out = op(_cuda(q), _cuda(kc), _cuda(vc), _cuda(seq), _cuda(table))
        ref = golden.paged_decode_swa(q, kc, vc, seq, table, None, "AABB", local, glob)
        torch.testing.assert_close(out.cpu().float(), ref.float(), atol=2e-2, rtol=2e-2)
        assert torch.count_nonzero(out[2]).item() == 0
        # sequence 0 (2900 keys): rewrite the keys strictly between the global prefix and the local window
        lo = 2899 - local if local is not None else 2900
        g_end = glob or 0
        first_page, last_page = (g_end + bs - 1) // bs, lo // bs
        if last_page - first_page >= 2:
            mid = table[0, first_page:last_page].long()
            kc2, vc2 = kc.clone(), vc.clone()
            kc2[mid] = torch.randn(kc2[mid].shape, generator=g).to(torch.bfloat16)
            vc2[mid] = torch.randn(vc2[mid].shape, generator=g).to(torch.bfloat16)  # masked in-tile keys: P = 0 exactly
            # pages of KV tiles (64 keys) that lie entirely outside both windows are never even loaded: poison them
            t_first, t_last = (g_end + 63) // 64, lo // 64
            if t_last > t_first:
                dead = table[0, t_first * 4:t_last * 4].long()
                kc2[dead] = float("nan")
                vc2[dead] = float("nan")
            out2 = op(_cuda(q), _cuda(kc2), _cuda(vc2), _cuda(seq), _cuda(table))
            assert torch.equal(out2[0], out[0])
    finally:
        os.environ.pop("MOJO_B200_DECODE_SPLITS", None)


@pytest.mark.parametrize("case", SDPA, ids=_ids(SDPA))
def test_sdpa(ops, case):
    op = ops.MojoSdpa(scale=case["scale"], enable_gqa=case["enable_gqa"])
    out = op(_cuda(case["query"]), _cuda(case["key"]), _cuda(case["value"]), _cuda(case["attn_mask"]))
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


SWA_NONPAGED = load_golden("swa.pt")


@pytest.mark.parametrize("case", SWA_NONPAGED, ids=_ids(SWA_NONPAGED))
def test_nonpaged_swa(ops, case):
    """MojoSWA (non-paged, packed var-len key / value) against the reference's outputs, through the op class -> C ABI."""
    op = ops.MojoSWA(is_causal=True, gqa_layout=case["gqa_layout"], global_window_size=case["global_window_size"],
                     local_window_size=case["local_window_size"])
    assert type(op).__name__ == "B200SWA"
    out = op(_cuda(case["query"]), _cuda(case["key"]), _cuda(case["value"]), _cuda(case["cu_q_lens"]),
             _cuda(case["cu_total_seq_lens"]))
    assert out.shape == case["out"].shape
    torch.testing.assert_close(out.cpu().float(), case["out"].float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("kernel", ["tcgen05", "mma"])
def test_nonpaged_swa_long_ragged_vs_oracle(ops, kernel):
    """Long ragged batch through both kernels (tile skipping, CTA pairs, keys of the NEXT sequence behind every tail)."""
    from oracle import golden

    os.environ["MOJO_B200_ATTN_IMPL"] = kernel
    try:
        g = torch.Generator().manual_seed(12)
        q_lens, prefix = [900, 257, 1, 1500], [0, 700, 333, 100]
        kv_lens = [a + b for a, b in zip(q_lens, prefix)]
        Hq, Hkv, D = 8, 2, 128
        q = torch.randn(sum(q_lens), Hq, D, generator=g).to(torch.bfloat16)
        k = torch.randn(sum(kv_lens), Hkv, D, generator=g).to(torch.bfloat16)
        v = torch.randn(sum(kv_lens), Hkv, D, generator=g).to(torch.bfloat16)
        cu_q = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32)
        cu_kv = torch.tensor([0] + torch.tensor(kv_lens).cumsum(0).tolist(), dtype=torch.int32)
        for local, glob in ((200, 40), (None, None), (1000, None)):
            op = ops.MojoSWA(global_window_size=glob, local_window_size=local)
            out = op(_cuda(q), _cuda(k), _cuda(v), _cuda(cu_q), _cuda(cu_kv))
            ref = golden.swa(q, k, v, cu_q, cu_kv, None, "AABB", True, local, glob)
            torch.testing.assert_close(out.cpu().float(), ref.float(), atol=2e-2, rtol=2e-2)
    finally:
        os.environ.pop("MOJO_B200_ATTN_IMPL", None)
