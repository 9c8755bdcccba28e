"""-m gpu: VALUE parity at BASELINE.json's stated sizes (VERDICT r1 "weak #1").

Every case runs the op class -> ctypes -> C ABI on the B200 and compares VALUES (not properties) with
``oracle.golden`` evaluated on the same tensors (the oracle is device-agnostic torch, so at config scale it runs on
CUDA tensors, where it finishes in seconds).  Bars: atol = rtol = 2e-2 for bf16/fp16 (north_star), 1e-5 /
1e-6 for fp32 (reference ``tests/accuracy/operators/test_attention.py:137-138``).  Long softmax rows of N(0,1) data
average thousands of values, so outputs are O(1e-2) and an absolute 2e-2 alone would accept almost anything: each case
ALSO bounds the relative Frobenius error ``|out - ref| / |ref|`` (bf16: 1.5e-2, fp32: 1e-5), which a wrong page
gather, head map, mask offset or 32-bit offset overflow cannot meet.

Cases:
  cfg1  MojoPagedDecodeGQA fp32, B8 32q/8kv hd128 page16 ctx1024 (exact config; oracle on the CPU as in BASELINE)
  cfg2  decode B64 32q/8kv ctx4096 page16 bf16 (all 64 sequences) + the whole layer step at that size
  cfg3  prefill T=8192 causal and the chunked 8192 + 8192-prefix variant
  cfg4  Llama-70B-shaped TP slices: TP8 (8q/1kv) and TP2 (32q/4kv), ctx 32768, 16 sequences whose pages are
        scattered over caches LARGER THAN 4 GiB (element offsets beyond 2^31: 64-bit addressing is exercised)
  ref   the reference's own decode list (test_attention.py:86-92: page 1024 at ctx 8192 / 2048, page 128 with
        head_dim 96, 16q/4kv page 32, the all-padding batch) in both GQA layouts, and a non-contiguous
        ``block_tables[layer]`` view of a ``[L, B, MB]`` table (reference ``modeling/qwen3/mojo_qwen3_dense.py:128-132``)
"""

import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    os.environ["MOJO_BACKEND"] = "b200"
    import mojo_opset_b200 as m

    assert m.MojoPagedDecodeGQA.get_registered_backends()[0] == "b200"
    return m


@pytest.fixture(scope="module")
def golden():
    from oracle import golden as g

    return g


@pytest.fixture(autouse=True)
def _free_cuda_memory():
    yield
    torch.cuda.empty_cache()


def check(out, ref, dtype, what=""):
    out, ref = out.float(), ref.float().to(out.device)
    assert out.shape == ref.shape
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    atol, rtol, rel = (1e-5, 1e-6, 1e-5) if dtype == torch.float32 else (2e-2, 2e-2, 1.5e-2)
    torch.testing.assert_close(out, ref, atol=atol, rtol=rtol, msg=lambda m: f"{what}: {m}")
    err = (out - ref).norm().item() / max(ref.norm().item(), 1e-30)
    assert err < rel, f"{what}: relative Frobenius error {err:.3e} >= {rel}"
    return err


def paged_kv(batch, lens, hkv, d, bs, dtype, seed, num_blocks=None, spare=10, device=DEV):
    """Caches filled on the device, a random-permutation block table (``-1`` padded), reference-test style
    (test_attention.py:56-81).  ``num_blocks`` larger than needed scatters the pages over a bigger cache."""
    g = torch.Generator().manual_seed(seed)
    need = [(n + bs - 1) // bs for n in lens]
    nb = (sum(need) + spare) if num_blocks is None else num_blocks
    gen = torch.Generator(device=device).manual_seed(seed)
    kc = torch.empty(nb, hkv, bs, d, dtype=dtype, device=device).normal_(generator=gen)
    vc = torch.empty(nb, hkv, bs, d, dtype=dtype, device=device).normal_(generator=gen)
    if nb <= (1 << 22):
        perm = torch.randperm(nb, generator=g)[: sum(need)]
    else:  # a huge cache: sample distinct ids without materialising the permutation; force the far end in
        perm = torch.unique(torch.randint(0, nb, (2 * sum(need),), generator=g))
        perm = perm[torch.randperm(perm.numel(), generator=g)][: sum(need)]
        perm[0], perm[-1] = nb - 1, nb - 2
    table = torch.full((batch, max(max(need), 1)), -1, dtype=torch.int32)
    pos = 0
    for i, n in enumerate(need):
        table[i, :n] = perm[pos:pos + n].to(torch.int32)
        pos += n
    return kc, vc, table.to(device)


# ------------------------------------------------------------------------------------------------------
# cfg1: the reference's own CPU-runnable case, exactly
# ------------------------------------------------------------------------------------------------------
def test_cfg1_decode_fp32_exact(ops, golden):
    B, Hq, Hkv, D, bs, ctx = 8, 32, 8, 128, 16, 1024
    g = torch.Generator().manual_seed(20260716 + 1)
    kc, vc, table = paged_kv(B, [ctx] * B, Hkv, D, bs, torch.float32, 20260716 + 1, device="cpu")
    q = torch.randn(B, Hq, D, generator=g)
    lens = torch.full((B,), ctx, dtype=torch.int32)
    ref = golden.paged_decode_gqa(q, kc, vc, lens, table)  # CPU, fp32: BASELINE.json configs[0]
    out = ops.MojoPagedDecodeGQA()(q.to(DEV), kc.to(DEV), vc.to(DEV), lens.to(DEV), table.to(DEV),
                                   max_total_seq_len=ctx)
    check(out.cpu(), ref, torch.float32, "cfg1")
    # ragged lengths of the same config (U[ctx/2, ctx]), one empty row
    lens2 = torch.randint(ctx // 2, ctx + 1, (B,), generator=g, dtype=torch.int32)
    lens2[3] = 0
    ref2 = golden.paged_decode_gqa(q, kc, vc, lens2, table)
    out2 = ops.MojoPagedDecodeGQA()(q.to(DEV), kc.to(DEV), vc.to(DEV), lens2.to(DEV), table.to(DEV))
    check(out2.cpu(), ref2, torch.float32, "cfg1 ragged")
    assert not out2[3].any()


# ------------------------------------------------------------------------------------------------------
# cfg2: Qwen3-8B-shaped decode, all 64 sequences, and the whole layer step
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ragged", [False, True], ids=["ctx4096", "ragged"])
@pytest.mark.parametrize("layout", ["AABB", "ABAB"])
def test_cfg2_decode_values(ops, golden, ragged, layout):
    B, Hq, Hkv, D, bs, ctx = 64, 32, 8, 128, 16, 4096
    g = torch.Generator().manual_seed(20260716 + 2)
    lens = torch.full((B,), ctx, dtype=torch.int32)
    if ragged:
        lens = torch.randint(ctx // 2, ctx + 1, (B,), generator=g, dtype=torch.int32)
    kc, vc, table = paged_kv(B, [ctx] * B, Hkv, D, bs, torch.bfloat16, 20260716 + 2)
    q = torch.randn(B, Hq, D, generator=g).to(torch.bfloat16).to(DEV)
    lens = lens.to(DEV)
    ref = golden.paged_decode_gqa(q, kc, vc, lens, table, None, layout)
    out = ops.MojoPagedDecodeGQA(gqa_layout=layout)(q, kc, vc, lens, table, max_total_seq_len=ctx)
    check(out, ref, torch.bfloat16, f"cfg2 decode {layout}")


def test_cfg2_layer_step_values(ops, golden):
    """bench.py's step at cfg2: ResidualAdd+RMSNorm 64x4096 -> RoPE -> StorePagedKVCache -> PagedDecodeGQA ->
    SwiGLU 64x12288; store bit-exact, RoPE bit-exact, the rest within their reference tolerances."""
    B, Hq, Hkv, D, bs, ctx, H, I = 64, 32, 8, 128, 16, 4096, 4096, 12288
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(7)
    kc, vc, table = paged_kv(B, [ctx] * B, Hkv, D, bs, dt, 11)
    r = lambda *s: torch.randn(*s, generator=g).to(dt).to(DEV)  # noqa: E731
    x, res, w = r(B, H), r(B, H), r(H)
    q, k, v = r(B, Hq, D), r(B, Hkv, D), r(B, Hkv, D)
    gate, up = r(B, I), r(B, I)
    ctx_lens = torch.full((B,), ctx - 1, dtype=torch.int32, device=DEV)
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))).to(DEV)
    cos, sin = golden.rotary_cos_sin(ctx_lens, inv_freq)

    y_ref, r_ref = golden.residual_add_rms_norm(x, res, w, 1e-6)
    q_ref, k_ref = golden.apply_rope(q, k, cos, sin, head_first=False)
    plan = golden.build_chunk_plan(table, None, ctx_lens, bs)
    kc_ref, vc_ref = golden.store_paged_kv(k_ref, v, kc.clone(), vc.clone(), plan)
    o_ref = golden.paged_decode_gqa(q_ref, kc_ref, vc_ref, ctx_lens + 1, table)
    s_ref = golden.swiglu(gate, up)

    norm = ops.MojoResidualAddRMSNorm(H, eps=1e-6, device=DEV, dtype=dt)
    with torch.no_grad():
        norm.weight.copy_(w)
        y, rs = norm(x, res)
    q_rot, k_rot = ops.MojoApplyRoPE()(q, k, cos, sin, head_first=False)
    ops.MojoStorePagedKVCache()(k_rot, v, kc, vc, table, None, ctx_lens)
    o = ops.MojoPagedDecodeGQA()(q_rot, kc, vc, ctx_lens + 1, table, max_total_seq_len=ctx)
    s = ops.MojoSwiGLU()(gate, up)

    assert torch.equal(rs, r_ref)
    torch.testing.assert_close(y.float(), y_ref.float(), atol=5e-2, rtol=1e-2)
    assert torch.equal(q_rot, q_ref) and torch.equal(k_rot, k_ref)
    assert torch.equal(kc, kc_ref) and torch.equal(vc, vc_ref)
    check(o, o_ref, dt, "cfg2 step decode")
    torch.testing.assert_close(s.float(), s_ref.float(), atol=1e-2, rtol=1e-2)


# ------------------------------------------------------------------------------------------------------
# cfg3: prefill T = 8192 and the chunked variant (8192 new tokens on an 8192-token cached prefix)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prefix", [0, 8192], ids=["T8192", "T8192+prefix8192"])
def test_cfg3_prefill_values(ops, golden, prefix):
    Hq, Hkv, D, bs, T = 32, 8, 128, 16, 8192
    g = torch.Generator().manual_seed(20260716 + 3)
    kv_len = T + prefix
    kc, vc, table = paged_kv(1, [kv_len], Hkv, D, bs, torch.bfloat16, 20260716 + 3)
    q = torch.randn(T, Hq, D, generator=g).to(torch.bfloat16).to(DEV)
    cu_q = torch.tensor([0, T], dtype=torch.int32, device=DEV)
    cu_kv = torch.tensor([0, kv_len], dtype=torch.int32, device=DEV)
    out = ops.MojoPagedPrefillGQA()(q, kc, vc, cu_q, table, cu_total_seq_lens=cu_kv if prefix else None,
                                    max_q_len=T, max_total_seq_len=kv_len)
    # the oracle materialises [T, Hq, kv] fp32 scores like the reference (8.6 / 17 GB): by head groups to bound memory
    ref = torch.empty_like(out)
    for h0 in range(0, Hq, 8):  # q heads h0..h0+7 <-> kv head h0/4.. (AABB, group 4): two kv heads per slice
        kv0, kv1 = h0 // 4, (h0 + 8) // 4
        ref[:, h0:h0 + 8] = golden.paged_prefill_gqa(q[:, h0:h0 + 8], kc[:, kv0:kv1], vc[:, kv0:kv1], cu_q, table,
                                                     None, cu_kv if prefix else None)
    check(out, ref, torch.bfloat16, f"cfg3 prefill prefix={prefix}")
    # early rows attend to few keys: errors there are not averaged away - bound them on their own
    check(out[:256], ref[:256], torch.bfloat16, "cfg3 first rows")


def test_cfg3_prefill_ragged_batch_values(ops, golden):
    """A ragged batch at cfg3 scale (the serving shape of chunked prefill): 5 sequences, 8192 query tokens in total,
    cached prefixes, one empty sequence."""
    Hq, Hkv, D, bs = 32, 8, 128, 16
    q_lens, prefixes = [3000, 0, 2500, 1, 2691], [0, 100, 5000, 4095, 777]
    kv_lens = [a + b if a else 0 for a, b in zip(q_lens, prefixes)]
    g = torch.Generator().manual_seed(33)
    kc, vc, table = paged_kv(len(q_lens), [max(n, 1) for n in kv_lens], Hkv, D, bs, torch.bfloat16, 33)
    q = torch.randn(sum(q_lens), Hq, D, generator=g).to(torch.bfloat16).to(DEV)
    cu_q = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32, device=DEV)
    cu_kv = torch.tensor([0] + torch.tensor(kv_lens).cumsum(0).tolist(), dtype=torch.int32, device=DEV)
    out = ops.MojoPagedPrefillGQA()(q, kc, vc, cu_q, table, cu_total_seq_lens=cu_kv, max_q_len=max(q_lens),
                                    max_total_seq_len=max(kv_lens))
    ref = golden.paged_prefill_gqa(q, kc, vc, cu_q, table, None, cu_kv)
    check(out, ref, torch.bfloat16, "cfg3 ragged")


# ------------------------------------------------------------------------------------------------------
# cfg4: Llama-3-70B-shaped decode, the per-rank slices of TP8 and TP2, caches beyond 4 GiB
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tp", [8, 2])
def test_cfg4_tp_slice_values(ops, golden, tp):
    from mojo_opset_b200.parallel import shard_heads

    Hq, Hkv, D, bs, ctx, B = 64, 8, 128, 16, 32768, 16
    shard = shard_heads(Hq, Hkv, tp, rank=tp - 1)
    hq_l, hkv_l = shard.q_end - shard.q_begin, shard.kv_end - shard.kv_begin
    assert (hq_l, hkv_l) == ((8, 1) if tp == 8 else (32, 4))
    block_bytes = hkv_l * bs * D * 2
    nb = (5 << 30) // block_bytes  # 5 GiB per cache tensor: page offsets cross 2^32 bytes and 2^31 elements
    g = torch.Generator().manual_seed(20260716 + 4)
    lens = torch.full((B,), ctx, dtype=torch.int32)
    lens[1::4] = torch.randint(ctx // 2, ctx, (B // 4,), generator=g, dtype=torch.int32)
    kc, vc, table = paged_kv(B, [ctx] * B, hkv_l, D, bs, torch.bfloat16, 20260716 + 4, num_blocks=nb)
    assert kc.numel() > (1 << 31) and int(table.max()) * block_bytes > (1 << 32)
    q = torch.randn(B, hq_l, D, generator=g).to(torch.bfloat16).to(DEV)
    lens = lens.to(DEV)
    out = ops.MojoPagedDecodeGQA()(q, kc, vc, lens, table, max_total_seq_len=ctx)
    ref = golden.paged_decode_gqa(q, kc, vc, lens, table)
    check(out, ref, torch.bfloat16, f"cfg4 TP{tp} slice")
    # the new token's KV lands in the far pages too: store (bit-exact) + decode again
    k_new = torch.randn(B, hkv_l, D, generator=g).to(torch.bfloat16).to(DEV)
    v_new = torch.randn(B, hkv_l, D, generator=g).to(torch.bfloat16).to(DEV)
    ctx_lens = lens - 1
    plan = golden.build_chunk_plan(table, None, ctx_lens, bs)
    pages = plan[:, 1].long()
    before_k, before_v = kc[pages].clone(), vc[pages].clone()
    ops.MojoStorePagedKVCache()(k_new, v_new, kc, vc, table, None, ctx_lens)
    exp_k, exp_v = before_k.clone(), before_v.clone()
    rows = torch.arange(B, device=DEV)
    exp_k[rows, :, plan[:, 2].long()] = k_new
    exp_v[rows, :, plan[:, 2].long()] = v_new
    assert torch.equal(kc[pages], exp_k) and torch.equal(vc[pages], exp_v)
    out2 = ops.MojoPagedDecodeGQA()(q, kc, vc, lens, table, max_total_seq_len=ctx)
    ref2 = golden.paged_decode_gqa(q, kc, vc, lens, table)
    check(out2, ref2, torch.bfloat16, f"cfg4 TP{tp} slice after store")


# ------------------------------------------------------------------------------------------------------
# the reference's own decode cases (tests/accuracy/operators/test_attention.py:33-151)
# ------------------------------------------------------------------------------------------------------
REF_DECODE = [
    (8, 16, 4, 128, 1024, 32, "M_BF16"),
    (8, 16, 4, 96, 1024, 128, "M_BF16_PADDIM"),
    (8, 8, 1, 128, 8192, 1024, "M_BF16_LONG"),
    (8, 8, 1, 128, 2048, 1024, "M_BF16_BIGPAGE"),
    (8, 8, 1, 128, 0, 1024, "M_BF16_PADSEQ"),
]


def _reference_decode_data(B, Hq, Hkv, D, max_seq_len, bs, seed):
    """Restates ``generate_paged_decode_data`` (test_attention.py:33-84) with a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    query = torch.randn(B, Hq, D, generator=g).to(torch.bfloat16)
    if max_seq_len > 0:
        lens = torch.randint(0, max_seq_len, (B,), generator=g, dtype=torch.int32).clamp(min=1)
    else:
        lens = torch.randperm(B, generator=g).to(torch.int32)  # 0 .. B-1: one row is pure padding
    max_len = int(lens.max())
    mb = (max_len + bs - 1) // bs
    need = ((lens + bs - 1) // bs).tolist()
    nb = (sum(need) or B * mb) + 10
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    table = torch.full((B, mb), -1, dtype=torch.int32)
    free = torch.randperm(nb, generator=g).to(torch.int32)
    pos = 0
    for i, n in enumerate(need):
        table[i, :n] = free[pos:pos + n]
        pos += n
    return query, kc, vc, lens, table, max_len


@pytest.mark.parametrize("layout", ["ABAB", "AABB"])
@pytest.mark.parametrize("cfg", REF_DECODE, ids=[c[-1] for c in REF_DECODE])
def test_reference_decode_list(ops, golden, cfg, layout):
    B, Hq, Hkv, D, max_seq_len, bs, _ = cfg
    q, kc, vc, lens, table, max_len = _reference_decode_data(B, Hq, Hkv, D, max_seq_len, bs, seed=86 + bs + D)
    ref = golden.paged_decode_gqa(q, kc, vc, lens, table, 1.0 / math.sqrt(D), layout)
    op = ops.MojoPagedDecodeGQA(is_causal=True, gqa_layout=layout)
    out = op(q.to(DEV), kc.to(DEV), vc.to(DEV), lens.to(DEV), table.to(DEV), softmax_scale=1.0 / math.sqrt(D),
             max_total_seq_len=max_len)
    check(out.cpu(), ref, torch.bfloat16, f"{cfg[-1]} {layout}")
    assert not out[lens.to(DEV) <= 0].any(), "padding rows must be zero"


def test_decode_layer_view_of_block_tables(ops, golden):
    """The in-tree models keep ONE table ``[L, B, MB_total]`` and hand each layer ``block_tables[layer, :, :mb]`` - a
    non-contiguous view whose row stride is not its width (mojo_qwen3_dense.py:50-74, 128-132)."""
    L, B, Hq, Hkv, D, bs, ctx, mb_total = 3, 6, 32, 8, 128, 16, 700, 80
    g = torch.Generator().manual_seed(5)
    mb = (ctx + bs - 1) // bs
    nb = L * B * mb + 7
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16).to(DEV)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16).to(DEV)
    full = torch.full((L, B, mb_total), -1, dtype=torch.int32)
    full[:, :, :mb] = torch.randperm(nb, generator=g)[: L * B * mb].view(L, B, mb).to(torch.int32)
    full = full.to(DEV)
    q = torch.randn(B, Hq, D, generator=g).to(torch.bfloat16).to(DEV)
    lens = torch.randint(1, ctx + 1, (B,), generator=g, dtype=torch.int32).to(DEV)
    for layer in range(L):
        view = full[layer, :, :mb]
        assert not view.is_contiguous() and view.stride(0) == mb_total
        out = ops.MojoPagedDecodeGQA()(q, kc, vc, lens, view)
        ref = golden.paged_decode_gqa(q, kc, vc, lens, view)
        check(out, ref, torch.bfloat16, f"layer {layer} view")
    # prefill through the same kind of view
    q_lens = [ctx // 2] * B
    cu = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32, device=DEV)
    qp = torch.randn(sum(q_lens), Hq, D, generator=g).to(torch.bfloat16).to(DEV)
    view = full[1, :, :mb]
    outp = ops.MojoPagedPrefillGQA()(qp, kc, vc, cu, view, max_q_len=ctx // 2)
    refp = golden.paged_prefill_gqa(qp, kc, vc, cu, view)
    check(outp, refp, torch.bfloat16, "prefill view")
