"""-m gpu: the fused pre-attention pass (per-head RMSNorm -> RoPE -> paged KV store, csrc/norm_rope_store.cu)
against the oracle's composition of the four reference ops on the same seeded inputs, through the op classes."""

import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    os.environ["MOJO_BACKEND"] = "b200"
    import mojo_opset_b200 as m

    return m


@pytest.fixture(scope="module")
def golden():
    from oracle import golden as g  # the checker

    return g


def _case(seed, q_lens, ctx_lens, Hq, Hkv, D, rope_dim, bs, dtype, cos_dtype, decode):
    g = torch.Generator().manual_seed(seed)
    B = len(ctx_lens)
    T = B if decode else sum(q_lens)
    q = torch.randn(T, Hq, D, generator=g).to(dtype)
    k = torch.randn(T, Hkv, D, generator=g).to(dtype)
    v = torch.randn(T, Hkv, D, generator=g).to(dtype)
    ang = torch.rand(T, rope_dim // 2, generator=g) * 6.28
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().to(cos_dtype), emb.sin().to(cos_dtype)
    need = [(max(c, 0) + (1 if decode else ql) + bs - 1) // bs for c, ql in zip(ctx_lens, q_lens)]
    mb = max(need) + 1
    nb = sum(need) + 5
    perm = torch.randperm(nb, generator=g).tolist()
    table = torch.full((B, mb), -1, dtype=torch.int32)
    it = iter(perm)
    for b in range(B):
        for j in range(need[b]):
            table[b, j] = next(it)
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(dtype)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(dtype)
    cu = None if decode else torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32)
    ctx = torch.tensor(ctx_lens, dtype=torch.int32)
    wq = (1 + 0.1 * torch.randn(D, generator=g)).to(dtype)
    wk = (1 + 0.1 * torch.randn(D, generator=g)).to(dtype)
    return dict(q=q, k=k, v=v, cos=cos, sin=sin, kc=kc, vc=vc, table=table, cu=cu, ctx=ctx, wq=wq, wk=wk)


CASES = [
    # q_lens, ctx_lens, Hq, Hkv, D, rope_dim, bs, dtype, cos dtype, decode
    ([1] * 5, [0, 17, 31, 32, -1], 32, 8, 128, 128, 16, torch.bfloat16, torch.float32, True),
    ([1] * 3, [5, 0, 100], 8, 2, 128, 128, 16, torch.float16, torch.float32, True),
    ([40, 0, 7, 130], [0, 3, 29, 16], 32, 8, 128, 128, 16, torch.bfloat16, torch.float32, False),
    ([33, 20], [5, 0], 6, 2, 64, 64, 8, torch.bfloat16, torch.bfloat16, False),
    ([9, 50], [0, 12], 4, 4, 128, 64, 32, torch.bfloat16, torch.float32, False),   # partial rotary
    ([21], [64], 3, 1, 256, 128, 16, torch.float16, torch.float16, False),
]


@pytest.mark.parametrize("case", CASES, ids=[f"D{c[4]}r{c[5]}bs{c[6]}{'dec' if c[9] else 'pre'}{i}" for i, c in enumerate(CASES)])
@pytest.mark.parametrize("with_norm", [True, False])
def test_fused_matches_composition(ops, golden, case, with_norm):
    q_lens, ctx_lens, Hq, Hkv, D, rope_dim, bs, dtype, cos_dtype, decode = case
    c = _case(len(q_lens) * 31 + D, q_lens, ctx_lens, Hq, Hkv, D, rope_dim, bs, dtype, cos_dtype, decode)
    kc_ref, vc_ref = c["kc"].clone(), c["vc"].clone()
    q_ref, k_ref = golden.norm_rope_store_kv(c["q"], c["k"], c["v"], c["cos"], c["sin"], kc_ref, vc_ref, c["table"],
                                             c["cu"], c["ctx"], c["wq"] if with_norm else None,
                                             c["wk"] if with_norm else None, 1e-6)
    d = {n: (t.to(DEV) if torch.is_tensor(t) else t) for n, t in c.items()}
    if with_norm:
        op = ops.MojoNormRoPEStoreKV(D, eps=1e-6, device=DEV, dtype=dtype)
        assert type(op).__name__ == "B200NormRoPEStoreKV"
        with torch.no_grad():
            op.q_weight.copy_(d["wq"])
            op.k_weight.copy_(d["wk"])
    else:
        op = ops.MojoRoPEStoreKV()
        assert type(op).__name__ == "B200RoPEStoreKV"
    q_out = op(d["q"], d["k"], d["v"], d["cos"], d["sin"], d["kc"], d["vc"], d["table"], d["cu"], d["ctx"])
    # V is moved bit-exactly; without a norm every rounding point is reproduced, so q and K are bit exact as well
    assert torch.equal(d["vc"].cpu(), vc_ref)
    if with_norm:
        torch.testing.assert_close(q_out.cpu().float(), q_ref.float(), atol=2e-2, rtol=2e-2)
        torch.testing.assert_close(d["kc"].cpu().float(), kc_ref.float(), atol=2e-2, rtol=2e-2)
        assert (q_out.cpu() != q_ref).float().mean().item() < 5e-3  # last-place flips of the fp32 row sum only
    else:
        assert torch.equal(q_out.cpu(), q_ref)
        assert torch.equal(d["kc"].cpu(), kc_ref)


def test_fused_equals_unfused_b200_chain(ops):
    """Same kernels' arithmetic as the four separate b200 ops on a strided fused-QKV projection output."""
    from mojo_opset_b200 import functional as F

    g = torch.Generator().manual_seed(77)
    T, Hq, Hkv, D, bs = 300, 32, 8, 128, 16
    qkv = torch.randn(T, (Hq + 2 * Hkv) * D, generator=g).to(torch.bfloat16).to(DEV)
    q = qkv[:, : Hq * D].view(T, Hq, D)
    k = qkv[:, Hq * D: (Hq + Hkv) * D].view(T, Hkv, D)
    v = qkv[:, (Hq + Hkv) * D:].view(T, Hkv, D)
    ang = torch.rand(T, D // 2, generator=g) * 6.28
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().to(DEV), emb.sin().to(DEV)
    nb = T // bs + 4
    table = torch.randperm(nb, generator=g)[: T // bs + 1].view(1, -1).to(torch.int32).to(DEV)
    cu = torch.tensor([0, T], dtype=torch.int32, device=DEV)
    ctx = torch.tensor([0], dtype=torch.int32, device=DEV)
    kc1 = torch.zeros(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV)
    vc1, kc2, vc2 = torch.zeros_like(kc1), torch.zeros_like(kc1), torch.zeros_like(kc1)
    q1, k1 = F.apply_rope(q, k, cos, sin, head_first=False)
    F.store_paged_kv(k1, v, kc1, vc1, block_table=table, cu_q_lens=cu, context_kv_lens=ctx)
    q2, k2 = F.norm_rope_store_kv(q, k, v, cos, sin, kc2, vc2, table, cu, ctx, want_k=True)
    assert torch.equal(q1, q2) and torch.equal(k1, k2)
    assert torch.equal(kc1, kc2) and torch.equal(vc1, vc2)
