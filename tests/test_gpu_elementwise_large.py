"""-m gpu: the HBM-bound ops (RMSNorm family, RoPE, SwiGLU/SiLU, KV store) at shapes that take the
prefill-sized code paths (multi-row CTAs, warp-per-token RoPE, two-vector activation trips) against the oracle
on the same seeded inputs, through the C ABI.  Also: the library zero-fills prefill output rows that no query
block covers (the host no longer memsets the output)."""

import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def F():
    os.environ["MOJO_BACKEND"] = "b200"
    from mojo_opset_b200 import functional

    return functional


@pytest.fixture(scope="module")
def golden():
    from oracle import golden as g  # the checker

    return g


def _tol(dtype):
    return dict(atol=1e-5, rtol=1e-5) if dtype == torch.float32 else dict(atol=2e-2, rtol=2e-2)


# every row-group width of the kernel: 4 ... 1024 threads per row, register-resident and streamed remainder,
# 16-byte and narrower packs
@pytest.mark.parametrize("hidden", [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 40960, 734, 7338, 100])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_rms_norm_shapes(F, golden, hidden, dtype):
    g = torch.Generator().manual_seed(hidden)
    rows = 515 if hidden <= 4096 else 37
    x = torch.randn(rows, hidden, generator=g).to(dtype)
    r = torch.randn(rows, hidden, generator=g).to(dtype)
    w = torch.randn(hidden, generator=g).to(dtype)
    y_ref = golden.rms_norm(x, w, 1e-6)
    y = F.rms_norm(x.to(DEV), w.to(DEV), 1e-6)
    torch.testing.assert_close(y.cpu().float(), y_ref.float(), **_tol(dtype))
    y_ref, s_ref = golden.residual_add_rms_norm(x, r, w, 1e-6)
    y, s = F.residual_add_rms_norm(x.to(DEV), r.to(DEV), w.to(DEV), 1e-6)
    assert torch.equal(s.cpu(), s_ref)  # the rounded residual sum is bit exact
    torch.testing.assert_close(y.cpu().float(), y_ref.float(), **_tol(dtype))


def test_rms_norm_head_rows_3d(F, golden):
    """q/k-norm: [T, heads, 128] rows of 256 bytes, 64 rows per CTA."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(300, 32, 128, generator=g).to(torch.bfloat16)
    w = torch.randn(128, generator=g).to(torch.bfloat16)
    y = F.rms_norm(x.to(DEV), w.to(DEV), 1e-6)
    torch.testing.assert_close(y.cpu().float(), golden.rms_norm(x, w, 1e-6).float(), atol=2e-2, rtol=2e-2)


ROPE_CASES = [
    # (T, Hq, Hkv, D, rope_dim, dtype, cos dtype)
    (300, 32, 8, 128, 128, torch.bfloat16, torch.float32),   # 8 items/head: warp kernel
    (300, 32, 8, 128, 128, torch.bfloat16, torch.bfloat16),  # intermediates rounded to bf16
    (300, 32, 8, 128, 128, torch.float16, torch.float32),
    (257, 6, 2, 64, 64, torch.bfloat16, torch.float32),      # 4 items/head, ragged last CTA
    (260, 5, 3, 128, 128, torch.float32, torch.float32),     # 16 items/head
    (300, 8, 2, 96, 64, torch.bfloat16, torch.float32),      # pass-through lanes in the warp kernel
    (300, 8, 2, 128, 64, torch.bfloat16, torch.float32),     # 12 items/head: general kernel
    (64, 32, 8, 128, 128, torch.bfloat16, torch.float32),    # decode-sized: CTA per token
]


@pytest.mark.parametrize("case", ROPE_CASES, ids=[str(c[:5]) + str(c[5]).split(".")[-1] + str(c[6]).split(".")[-1]
                                                  for c in ROPE_CASES])
@pytest.mark.parametrize("head_first", [False, True])
def test_apply_rope_large(F, golden, case, head_first):
    T, Hq, Hkv, D, rd, dtype, cdt = case
    g = torch.Generator().manual_seed(T + D + rd)
    q = torch.randn(T, Hq, D, generator=g).to(dtype)
    k = torch.randn(T, Hkv, D, generator=g).to(dtype)
    ang = torch.rand(T, rd // 2, generator=g) * 6.28
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().to(cdt), emb.sin().to(cdt)
    if head_first:  # [N, T, D] views of the same memory
        q, k = q.transpose(0, 1), k.transpose(0, 1)
    q_ref, k_ref = golden.apply_rope(q, k, cos, sin, head_first=head_first)
    q_out, k_out = F.apply_rope(q.to(DEV), k.to(DEV), cos.to(DEV), sin.to(DEV), head_first=head_first)
    # same rounding points as the eager golden: bit exact
    assert torch.equal(q_out.cpu(), q_ref)
    assert torch.equal(k_out.cpu(), k_ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(300, 12288), (1, 7), (33, 1000), (5, 4096 * 3 + 8)])
def test_swiglu_large(F, golden, dtype, shape):
    g = torch.Generator().manual_seed(shape[1])
    gate = (torch.randn(*shape, generator=g) * 3).to(dtype)
    up = torch.randn(*shape, generator=g).to(dtype)
    sixteen = dtype != torch.float32
    tol = dict(atol=1e-2, rtol=1e-2) if sixteen else dict(atol=1e-6, rtol=1e-5)
    out = F.swiglu(gate.to(DEV), up.to(DEV))
    ref = golden.swiglu(gate, up)
    torch.testing.assert_close(out.cpu().float(), ref.float(), **tol)
    s = F.silu(gate.to(DEV))
    torch.testing.assert_close(s.cpu().float(), golden.silu(gate).float(), **tol)
    if sixteen:  # same rounding points: only last-place effects of the fp32 exp / reciprocal can show through
        assert (out.cpu() != ref).float().mean().item() < 2e-3
        assert (s.cpu() != golden.silu(gate)).float().mean().item() < 2e-3
    # strided halves of one fused projection
    fused = torch.cat((gate, up), -1).to(DEV)
    out2 = F.swiglu(fused[:, : shape[1]], fused[:, shape[1]:])
    assert torch.equal(out2, out)


def test_store_kv_prefill_sized(F, golden):
    g = torch.Generator().manual_seed(11)
    T, Hkv, D, bs = 2048 + 5, 8, 128, 16
    nb = (T + bs - 1) // bs + 7
    k = torch.randn(T, Hkv, D, generator=g).to(torch.bfloat16)
    v = torch.randn(T, Hkv, D, generator=g).to(torch.bfloat16)
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    table = torch.randperm(nb, generator=g)[: nb - 7].view(1, -1).to(torch.int32)
    cu = torch.tensor([0, T], dtype=torch.int32)
    ctx = torch.tensor([3], dtype=torch.int32)  # starts inside a page
    plan = golden.build_chunk_plan(table, cu, ctx, bs)
    kc_ref, vc_ref = golden.store_paged_kv(k, v, kc.clone(), vc.clone(), plan)
    kc_d, vc_d = kc.to(DEV), vc.to(DEV)
    F.store_paged_kv(k.to(DEV), v.to(DEV), kc_d, vc_d, block_table=table.to(DEV), cu_q_lens=cu.to(DEV),
                     context_kv_lens=ctx.to(DEV))
    assert torch.equal(kc_d.cpu(), kc_ref) and torch.equal(vc_d.cpu(), vc_ref)
    kc_d, vc_d = kc.to(DEV), vc.to(DEV)
    F.store_paged_kv(k.to(DEV), v.to(DEV), kc_d, vc_d, chunk_metadata=plan.to(DEV))
    assert torch.equal(kc_d.cpu(), kc_ref) and torch.equal(vc_d.cpu(), vc_ref)


@pytest.mark.parametrize("impl", ["tcgen05", "mma"])
def test_prefill_zero_fills_uncovered_rows(F, golden, impl):
    """Rows past cu_q_lens[-1], rows of sequences without keys and rows whose causal window is empty
    (kv_len < q_len) read as zeros although the output buffer starts out poisoned."""
    os.environ["MOJO_B200_ATTN_IMPL"] = impl
    try:
        g = torch.Generator().manual_seed(3)
        Hq, Hkv, D, bs = 4, 2, 128, 16
        q_lens, kv_lens = [300, 0, 260, 40], [300, 0, 200, 0]  # seq 2: 60 leading rows see nothing; seq 3: no keys
        tail = 9
        T = sum(q_lens) + tail
        mb = 20
        nb = 4 * mb + 3
        q = torch.randn(T, Hq, D, generator=g).to(torch.bfloat16)
        kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
        vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
        table = torch.randperm(nb, generator=g)[: 4 * mb].view(4, mb).to(torch.int32)
        cu_q = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32)
        cu_kv = torch.tensor([0] + list(torch.tensor(kv_lens).cumsum(0)), dtype=torch.int32)
        # poison the caching allocator's next block of this size
        junk = torch.full((T, Hq, D), float("nan"), dtype=torch.bfloat16, device=DEV)
        del junk
        out = F.paged_prefill_gqa(q.to(DEV), kc.to(DEV), vc.to(DEV), cu_q.to(DEV), table.to(DEV), None,
                                  cu_kv.to(DEV), max_q_len=300, max_total_seq_len=300).cpu()
        assert not torch.isnan(out.float()).any()
        assert torch.count_nonzero(out[sum(q_lens):]) == 0                      # tail
        assert torch.count_nonzero(out[300 + 0 + 260: 300 + 0 + 260 + 40]) == 0  # sequence without keys
        assert torch.count_nonzero(out[300: 300 + 60]) == 0                      # empty causal windows
        # the covered rows still match the oracle
        ref = golden.paged_prefill_gqa(q[:300], kc, vc, cu_q[:2], table[:1], None, cu_kv[:2])
        torch.testing.assert_close(out[:300].float(), ref.float(), atol=2e-2, rtol=2e-2)
        assert torch.count_nonzero(out[360:560]) > 0
    finally:
        os.environ.pop("MOJO_B200_ATTN_IMPL", None)
