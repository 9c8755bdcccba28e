"""Generate the golden input/output vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 MOJO_BACKEND=torch python tests/golden/make_golden.py

Every case is: seeded inputs -> the reference's torch-native op (``MojoXxx._registry.get("torch")``) on
CPU -> inputs + outputs saved as one ``.pt`` dict per op.  Shapes are small so the fixtures stay a few MB;
they cover the edge cases the reference's own tests exercise (zero-length rows, -1 block ids, AABB/ABAB,
ragged var-len prefill with a cached prefix, partial RoPE, non-power-of-two hidden sizes, pre/post norm,
negative-context rows in the KV store).
"""

import math
import os
import sys

import torch

os.environ.setdefault("MOJO_BACKEND", "torch")
sys.path.insert(0, os.environ.get("MOJO_REFERENCE_ROOT", "/root/reference"))

import mojo_opset as ref  # noqa: E402  (the reference)
from mojo_opset.core.operators.kv_cache import build_paged_kv_chunk_metadata  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def torch_op(cls, *args, **kwargs):
    return cls._registry.get("torch")(*args, **kwargs)


def paged_cache(gen, lens, num_kv_heads, head_dim, block_size, dtype, spare=3):
    """Random cache + a randomly permuted block table (same recipe as the reference's
    tests/accuracy/operators/test_attention.py:33-83)."""
    need = [(n + block_size - 1) // block_size for n in lens]
    width = max(max(need), 1)
    total = sum(need) + spare
    kc = torch.randn(total, num_kv_heads, block_size, head_dim, generator=gen).to(dtype)
    vc = torch.randn(total, num_kv_heads, block_size, head_dim, generator=gen).to(dtype)
    table = torch.full((len(lens), width), -1, dtype=torch.int32)
    free = torch.randperm(total, generator=gen).to(torch.int32)
    at = 0
    for i, n in enumerate(need):
        table[i, :n] = free[at:at + n]
        at += n
    return kc, vc, table


def gen_decode():
    cases = []
    specs = [
        # name, lens, Hq, Hkv, D, bs, dtype, layout
        ("bf16_aabb", [37, 1, 128, 95, 16], 8, 2, 128, 16, torch.bfloat16, "AABB"),
        ("bf16_abab", [60, 33, 17], 8, 2, 128, 32, torch.bfloat16, "ABAB"),
        ("bf16_padseq", [0, 50, 0, 9], 8, 1, 128, 16, torch.bfloat16, "AABB"),
        ("fp16_d64", [70, 41], 4, 4, 64, 16, torch.float16, "AABB"),
        ("fp32_cfg1_small", [100, 64, 7], 8, 2, 128, 16, torch.float32, "AABB"),
        ("bf16_d96", [40, 23], 8, 2, 96, 16, torch.bfloat16, "AABB"),
    ]
    for i, (name, lens, hq, hkv, d, bs, dtype, layout) in enumerate(specs):
        gen = torch.Generator().manual_seed(1000 + i)
        kc, vc, table = paged_cache(gen, lens, hkv, d, bs, dtype)
        q = torch.randn(len(lens), hq, d, generator=gen).to(dtype)
        seq = torch.tensor(lens, dtype=torch.int32)
        scale = 1.0 / math.sqrt(d)
        out = torch_op(ref.MojoPagedDecodeGQA, is_causal=True, gqa_layout=layout)(
            q, kc, vc, seq, table, softmax_scale=scale, max_total_seq_len=max(lens)
        )
        cases.append(dict(name=name, gqa_layout=layout, query=q, key_cache=kc, value_cache=vc, total_seq_lens=seq,
                          block_tables=table, softmax_scale=scale, out=out))
    torch.save(cases, os.path.join(HERE, "paged_decode_gqa.pt"))


def gen_prefill():
    cases = []
    specs = [
        # name, q_lens, ctx_lens (cached prefix), Hq, Hkv, D, bs, dtype, layout
        ("bf16_nocache", [33, 70, 5], None, 4, 2, 128, 16, torch.bfloat16, "AABB"),
        ("bf16_prefix", [20, 48], [37, 16], 4, 2, 128, 16, torch.bfloat16, "AABB"),
        ("bf16_abab_empty", [17, 0, 40], [0, 0, 25], 4, 2, 128, 32, torch.bfloat16, "ABAB"),
        ("fp16_d64", [64, 31], None, 4, 1, 64, 16, torch.float16, "AABB"),
    ]
    for i, (name, q_lens, ctx, hq, hkv, d, bs, dtype, layout) in enumerate(specs):
        gen = torch.Generator().manual_seed(2000 + i)
        kv_lens = q_lens if ctx is None else [a + b for a, b in zip(q_lens, ctx)]
        kc, vc, table = paged_cache(gen, kv_lens, hkv, d, bs, dtype)
        q = torch.randn(sum(q_lens), hq, d, generator=gen).to(dtype)
        cu_q = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32)
        cu_kv = None if ctx is None else torch.tensor([0] + list(torch.tensor(kv_lens).cumsum(0)), dtype=torch.int32)
        scale = 1.0 / math.sqrt(d)
        out = torch_op(ref.MojoPagedPrefillGQA, is_causal=True, gqa_layout=layout)(
            q, kc, vc, cu_q, table, softmax_scale=scale, cu_total_seq_lens=cu_kv
        )
        cases.append(dict(name=name, gqa_layout=layout, query=q, key_cache=kc, value_cache=vc, cu_q_lens=cu_q,
                          block_tables=table, softmax_scale=scale, cu_total_seq_lens=cu_kv, out=out))
    torch.save(cases, os.path.join(HERE, "paged_prefill_gqa.pt"))


def gen_swa():
    """MojoPagedPrefillSWA / MojoPagedDecodeSWA (reference attention.py:533-745): local / global / both / no window,
    windows shorter and longer than the sequences, cached prefixes, ABAB, a zero-length row.  Written by
    ``python tests/golden/make_golden.py swa`` (only this fixture is regenerated)."""
    cases = []
    specs = [
        # name, q_lens, ctx_lens, Hq, Hkv, D, bs, dtype, layout, local, global
        ("local_only", [150, 70, 5], None, 4, 2, 128, 16, torch.bfloat16, "AABB", 32, None),
        ("local_global_prefix", [90, 130], [137, 16], 4, 2, 128, 16, torch.bfloat16, "AABB", 48, 8),
        ("global_only_abab", [77, 0, 40], [0, 0, 25], 4, 2, 128, 32, torch.bfloat16, "ABAB", None, 20),
        ("wide_windows", [64, 31], None, 4, 1, 64, 16, torch.float16, "AABB", 1000, 1000),
        ("no_window", [40, 21], [3, 0], 4, 2, 128, 16, torch.bfloat16, "AABB", None, None),
        ("local_zero", [33], [100], 2, 2, 128, 16, torch.bfloat16, "AABB", 0, 4),
    ]
    for i, (name, q_lens, ctx, hq, hkv, d, bs, dtype, layout, local, glob) in enumerate(specs):
        gen = torch.Generator().manual_seed(8000 + i)
        kv_lens = q_lens if ctx is None else [a + b for a, b in zip(q_lens, ctx)]
        kc, vc, table = paged_cache(gen, kv_lens, hkv, d, bs, dtype)
        q = torch.randn(sum(q_lens), hq, d, generator=gen).to(dtype)
        cu_q = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32)
        cu_kv = None if ctx is None else torch.tensor([0] + list(torch.tensor(kv_lens).cumsum(0)), dtype=torch.int32)
        scale = 1.0 / math.sqrt(d)
        out = torch_op(ref.MojoPagedPrefillSWA, is_causal=True, gqa_layout=layout, global_window_size=glob,
                       local_window_size=local)(q, kc, vc, cu_q, table, softmax_scale=scale, cu_total_seq_lens=cu_kv)
        cases.append(dict(op="prefill", name=name, gqa_layout=layout, local_window_size=local, global_window_size=glob,
                          query=q, key_cache=kc, value_cache=vc, cu_q_lens=cu_q, block_table=table,
                          softmax_scale=scale, cu_total_seq_lens=cu_kv, out=out))
    dspecs = [
        # name, lens, Hq, Hkv, D, bs, dtype, layout, local, global
        ("dec_local", [137, 1, 128, 95, 16], 8, 2, 128, 16, torch.bfloat16, "AABB", 32, None),
        ("dec_both_abab", [260, 33, 17], 8, 2, 128, 32, torch.bfloat16, "ABAB", 64, 4),
        ("dec_padseq_global", [0, 150, 0, 9], 8, 1, 128, 16, torch.bfloat16, "AABB", 16, 100),
        ("dec_fp16_d64", [70, 41], 4, 4, 64, 16, torch.float16, "AABB", 7, 3),
    ]
    for i, (name, lens, hq, hkv, d, bs, dtype, layout, local, glob) in enumerate(dspecs):
        gen = torch.Generator().manual_seed(8100 + i)
        kc, vc, table = paged_cache(gen, lens, hkv, d, bs, dtype)
        q = torch.randn(len(lens), hq, d, generator=gen).to(dtype)
        seq = torch.tensor(lens, dtype=torch.int32)
        scale = 1.0 / math.sqrt(d)
        out = torch_op(ref.MojoPagedDecodeSWA, is_causal=True, gqa_layout=layout, global_window_size=glob,
                       local_window_size=local)(q, kc, vc, seq, table, softmax_scale=scale)
        cases.append(dict(op="decode", name=name, gqa_layout=layout, local_window_size=local, global_window_size=glob,
                          query=q, key_cache=kc, value_cache=vc, total_seq_lens=seq, block_table=table,
                          softmax_scale=scale, out=out))
    torch.save(cases, os.path.join(HERE, "paged_swa.pt"))


def gen_sdpa():
    cases = []
    gen = torch.Generator().manual_seed(3000)
    # DiT-style: transposed views of [B,S,H,D] memory, no mask, scale None
    b, s, h, d = 2, 96, 3, 128
    q, k, v = (torch.randn(b, s, h, d, generator=gen).to(torch.bfloat16).transpose(1, 2) for _ in range(3))
    out = torch_op(ref.MojoSdpa)(q, k, v)
    cases.append(dict(name="dit_bf16_strided", scale=None, enable_gqa=False, query=q, key=k, value=v,
                      attn_mask=None, out=out))
    # cross attention Skv != Sq
    q = torch.randn(1, 2, 80, 128, generator=gen).to(torch.bfloat16)
    k = torch.randn(1, 2, 48, 128, generator=gen).to(torch.bfloat16)
    v = torch.randn(1, 2, 48, 128, generator=gen).to(torch.bfloat16)
    cases.append(dict(name="cross_bf16", scale=None, enable_gqa=False, query=q, key=k, value=v, attn_mask=None,
                      out=torch_op(ref.MojoSdpa)(q, k, v)))
    # masked GQA (reference test_attention.py:899-949 style)
    q = torch.randn(1, 4, 64, 64, generator=gen).to(torch.bfloat16)
    k = torch.randn(1, 2, 64, 64, generator=gen).to(torch.bfloat16)
    v = torch.randn(1, 2, 64, 64, generator=gen).to(torch.bfloat16)
    mask = torch.rand(64, 64, generator=gen) > 0.3
    mask[:, 0] = True
    cases.append(dict(name="masked_gqa_bf16", scale=0.2, enable_gqa=True, query=q, key=k, value=v, attn_mask=mask,
                      out=torch_op(ref.MojoSdpa, scale=0.2, enable_gqa=True)(q, k, v, mask)))
    torch.save(cases, os.path.join(HERE, "sdpa.pt"))


def gen_store_kv():
    cases = []
    specs = [
        # name, q_lens (None = decode), ctx_lens, Hkv, D, bs, dtype
        ("decode_bf16", None, [5, 16, -1, 31, 0], 2, 128, 16, torch.bfloat16),
        ("prefill_bf16", [20, 1, 0, 40], [3, 15, 7, 0], 2, 128, 16, torch.bfloat16),
        ("prefill_fp16_big_page", [100, 30], [60, 0], 1, 64, 128, torch.float16),
        ("prefill_fp32_neg_ctx", [9, 12], [-1, 4], 2, 32, 8, torch.float32),
    ]
    for i, (name, q_lens, ctx, hkv, d, bs, dtype) in enumerate(specs):
        gen = torch.Generator().manual_seed(4000 + i)
        n_seq = len(ctx)
        new = [1] * n_seq if q_lens is None else q_lens
        final = [max(c, 0) + n for c, n in zip(ctx, new)]
        kc, vc, table = paged_cache(gen, final, hkv, d, bs, dtype)
        if name == "decode_bf16":
            table[3, 1] = -1  # ctx 31 -> logical block 1 unmapped: row must be dropped
        tokens = sum(new)
        ks = torch.randn(tokens, hkv, d, generator=gen).to(dtype)
        vs = torch.randn(tokens, hkv, d, generator=gen).to(dtype)
        cu_q = None if q_lens is None else torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32)
        ctx_t = torch.tensor(ctx, dtype=torch.int32)
        plan = build_paged_kv_chunk_metadata(table, cu_q, ctx_t, bs)
        kc_out, vc_out = torch_op(ref.MojoStorePagedKVCache)(ks, vs, kc.clone(), vc.clone(), table, cu_q, ctx_t)
        kc_out2, vc_out2 = torch_op(ref.MojoStorePagedKVCache)(ks, vs, kc.clone(), vc.clone(), chunk_metadata=plan)
        assert torch.equal(kc_out, kc_out2) and torch.equal(vc_out, vc_out2)
        cases.append(dict(name=name, key_states=ks, value_states=vs, key_cache=kc, value_cache=vc, block_table=table,
                          cu_q_lens=cu_q, context_kv_lens=ctx_t, chunk_metadata=plan, key_cache_out=kc_out,
                          value_cache_out=vc_out))
    torch.save(cases, os.path.join(HERE, "store_paged_kv.pt"))


def gen_norm():
    cases = []
    specs = [
        ("pre_bf16", (6, 4096), torch.bfloat16, "pre", 1e-6),
        ("post_bf16", (2, 3, 1024), torch.bfloat16, "post", 1e-5),
        ("pre_fp16_nonpow2", (5, 734), torch.float16, "pre", 1e-5),
        ("pre_fp32", (3, 512), torch.float32, "pre", 1e-5),
    ]
    for i, (name, shape, dtype, pos, eps) in enumerate(specs):
        gen = torch.Generator().manual_seed(5000 + i)
        x = torch.randn(*shape, generator=gen).to(dtype)
        r = torch.randn(*shape, generator=gen).to(dtype)
        w = torch.randn(shape[-1], generator=gen).to(dtype)
        op = torch_op(ref.MojoResidualAddRMSNorm, norm_size=shape[-1], eps=eps, norm_pos=pos, dtype=dtype)
        with torch.no_grad():
            op.weight.copy_(w)
            y, res = op(x, r)
        rn = torch_op(ref.MojoRMSNorm, norm_size=shape[-1], eps=eps, dtype=dtype)
        with torch.no_grad():
            rn.weight.copy_(w)
            y_plain = rn(x)
        cases.append(dict(name=name, norm_pos=pos, eps=eps, hidden_state=x, residual=r, weight=w, out=y.detach(),
                          residual_out=res.detach(), rmsnorm_out=y_plain.detach()))
    # q/k-norm shape: normalise over head_dim
    gen = torch.Generator().manual_seed(5100)
    x = torch.randn(7, 8, 128, generator=gen).to(torch.bfloat16)
    w = torch.randn(128, generator=gen).to(torch.bfloat16)
    rn = torch_op(ref.MojoRMSNorm, norm_size=128, eps=1e-6, dtype=torch.bfloat16)
    with torch.no_grad():
        rn.weight.copy_(w)
        cases.append(dict(name="qk_norm_bf16", norm_pos=None, eps=1e-6, hidden_state=x, residual=None, weight=w,
                          out=None, residual_out=None, rmsnorm_out=rn(x).detach()))
    torch.save(cases, os.path.join(HERE, "rmsnorm.pt"))


def gen_rope():
    cases = []
    gen = torch.Generator().manual_seed(6000)

    def table(n, d, dtype):
        rot = torch_op(ref.MojoRotaryEmbedding, rope_theta=1e6, rope_dim=d)
        pos = torch.randint(0, 4096, (n,), generator=gen, dtype=torch.int32)
        cos, sin = rot(torch.empty(n, 8), position_ids=pos)
        return cos.to(dtype), sin.to(dtype)

    # decode / varlen [T,N,D] token-first, fp32 cos
    q = torch.randn(9, 8, 128, generator=gen).to(torch.bfloat16)
    k = torch.randn(9, 2, 128, generator=gen).to(torch.bfloat16)
    cos, sin = table(9, 128, torch.float32)
    qo, ko = torch_op(ref.MojoApplyRoPE)(q, k, cos, sin, head_first=False)
    cases.append(dict(name="tnd_fp32cos", q=q, k=k, cos=cos, sin=sin, head_first=False, q_out=qo, k_out=ko))
    # [N,T,D] head-first transposed view, bf16 cos (math in bf16)
    q = torch.randn(9, 8, 128, generator=gen).to(torch.bfloat16).transpose(0, 1)
    k = torch.randn(9, 2, 128, generator=gen).to(torch.bfloat16).transpose(0, 1)
    cos, sin = table(9, 128, torch.bfloat16)
    qo, ko = torch_op(ref.MojoApplyRoPE)(q, k, cos, sin, head_first=True)
    cases.append(dict(name="ntd_bf16cos", q=q, k=k, cos=cos, sin=sin, head_first=True, q_out=qo, k_out=ko))
    # [B,N,S,D] view of BSND memory (Qwen3 call site), cos [B,S,d], partial rope 96 of 128
    q = torch.randn(2, 11, 4, 128, generator=gen).to(torch.bfloat16).transpose(1, 2)
    k = torch.randn(2, 11, 2, 128, generator=gen).to(torch.bfloat16).transpose(1, 2)
    cos, sin = table(22, 96, torch.float32)
    cos, sin = cos.view(2, 11, 96), sin.view(2, 11, 96)
    qo, ko = torch_op(ref.MojoApplyRoPE)(q, k, cos, sin, head_first=True)
    cases.append(dict(name="bnsd_partial", q=q, k=k, cos=cos, sin=sin, head_first=True, q_out=qo, k_out=ko))
    # [B,S,N,D], cos [S,d], fp16, D=88
    q = torch.randn(2, 5, 3, 88, generator=gen).to(torch.float16)
    k = torch.randn(2, 5, 1, 88, generator=gen).to(torch.float16)
    cos, sin = table(5, 88, torch.float32)
    qo, ko = torch_op(ref.MojoApplyRoPE)(q, k, cos, sin, head_first=False)
    cases.append(dict(name="bsnd_d88_fp16", q=q, k=k, cos=cos, sin=sin, head_first=False, q_out=qo, k_out=ko))
    torch.save(cases, os.path.join(HERE, "apply_rope.pt"))

    # rotary embedding: varlen with prefix, decode ids, padded
    rot_cases = []
    rot = torch_op(ref.MojoRotaryEmbedding, rope_theta=1e6, rope_dim=128)
    cu = torch.tensor([0, 5, 5, 12], dtype=torch.int32)
    tot = torch.tensor([9, 3, 7], dtype=torch.int32)
    c, s = rot(torch.empty(12, 16), cu_q_lens=cu, total_seq_lens=tot)
    rot_cases.append(dict(name="varlen_prefix", rope_theta=1e6, rope_dim=128, x_shape=(12, 16), cu_q_lens=cu,
                          total_seq_lens=tot, position_ids=None, cos=c, sin=s))
    pos = torch.tensor([0, 17, 4095, 31999], dtype=torch.int32)
    c, s = rot(torch.empty(4, 16), position_ids=pos)
    rot_cases.append(dict(name="decode_ids", rope_theta=1e6, rope_dim=128, x_shape=(4, 16), cu_q_lens=None,
                          total_seq_lens=None, position_ids=pos, cos=c, sin=s))
    c, s = rot(torch.empty(2, 6, 16))
    rot_cases.append(dict(name="padded", rope_theta=1e6, rope_dim=128, x_shape=(2, 6, 16), cu_q_lens=None,
                          total_seq_lens=None, position_ids=None, cos=c, sin=s))
    torch.save(rot_cases, os.path.join(HERE, "rotary_embedding.pt"))


def gen_act():
    cases = []
    for i, (shape, dtype, limit) in enumerate([
        ((5, 768), torch.bfloat16, 0.0),
        ((3, 7, 33), torch.float16, 0.0),
        ((9, 99), torch.float32, 0.0),
        ((4, 256), torch.bfloat16, 1.5),
    ]):
        gen = torch.Generator().manual_seed(7000 + i)
        g = (torch.randn(*shape, generator=gen) * 3).to(dtype)
        u = (torch.randn(*shape, generator=gen) * 3).to(dtype)
        out = torch_op(ref.MojoSwiGLU, swiglu_limit=limit)(g, u)
        cases.append(dict(name=f"swiglu_{i}", swiglu_limit=limit, gate=g, up=u, out=out,
                          silu_out=torch_op(ref.MojoSilu)(g)))
    torch.save(cases, os.path.join(HERE, "activation.pt"))


if __name__ == "__main__":
    torch.manual_seed(0)
    print("reference:", ref.__file__, "torch", torch.__version__, "threads", torch.get_num_threads())
    if "swa" in sys.argv[1:]:  # later addition: regenerate this fixture alone (the others stay byte-identical)
        gen_swa()
        print("wrote gen_swa")
        sys.exit(0)
    for fn in (gen_decode, gen_prefill, gen_sdpa, gen_store_kv, gen_norm, gen_rope, gen_act, gen_swa):
        fn()
        print("wrote", fn.__name__)
