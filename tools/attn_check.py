#!/usr/bin/env python
"""Developer check of the two attention kernels (tcgen05 vs mma.sync general path) on a B200: numerics against
an fp32 torch restatement computed on the GPU, then timings at cfg3 / cfg5.  Not part of the product path.

    python tools/attn_check.py [--no-bench] [--only sdpa|prefill]
"""

import argparse
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
import mojo_opset_b200 as m  # noqa: E402
from mojo_opset_b200 import functional as F  # noqa: E402

DEV = "cuda"


def set_impl(name, min_q=None):
    if name is None:
        os.environ.pop("MOJO_B200_ATTN_IMPL", None)
    else:
        os.environ["MOJO_B200_ATTN_IMPL"] = name
    if min_q is not None:
        os.environ["MOJO_B200_ATTN_TCGEN05_MIN_Q"] = str(min_q)


def ref_attn(q, k, v, scale, causal_off=None):
    """q [H,Sq,D], k/v [H,Skv,D] fp32 math.  causal_off: query row t sees keys <= causal_off + t."""
    s = torch.einsum("hqd,hkd->hqk", q.float(), k.float()) * scale
    if causal_off is not None:
        sq, sk = s.shape[-2:]
        keep = torch.arange(sk, device=s.device)[None, :] <= (torch.arange(sq, device=s.device)[:, None] + causal_off)
        s = s.masked_fill(~keep, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.einsum("hqk,hkd->hqd", p, v.float())


def check_sdpa(B, Hq, Hkv, Sq, Skv, D=128, dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    q = torch.randn(B, Sq, Hq, D, device=DEV, generator=g).to(dtype).transpose(1, 2)
    k = torch.randn(B, Skv, Hkv, D, device=DEV, generator=g).to(dtype).transpose(1, 2)
    v = torch.randn(B, Skv, Hkv, D, device=DEV, generator=g).to(dtype).transpose(1, 2)
    G = Hq // Hkv
    ref = torch.stack([ref_attn(q[b], k[b].repeat_interleave(G, 0), v[b].repeat_interleave(G, 0), 1 / math.sqrt(D))
                       for b in range(B)])
    out = {}
    for impl in ("mma", "tcgen05"):
        set_impl(impl)
        try:
            o = F.sdpa(q, k, v, None, enable_gqa=G > 1)
            torch.cuda.synchronize()
            out[impl] = (o.float() - ref).abs().max().item()
        except Exception as e:  # noqa: BLE001
            out[impl] = f"ERR {type(e).__name__}: {e}"
    set_impl(None)
    print(f"sdpa B{B} Hq{Hq} Hkv{Hkv} Sq{Sq} Skv{Skv} {str(dtype)[6:]}: max|err| " +
          " ".join(f"{k_}={v_ if isinstance(v_, str) else format(v_, '.4f')}" for k_, v_ in out.items()), flush=True)
    return out


def check_prefill(q_lens, prefix_lens, Hq, Hkv, bs, D=128, dtype=torch.bfloat16, layout="AABB", seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    B = len(q_lens)
    kv_lens = [a + b for a, b in zip(q_lens, prefix_lens)]
    blocks = [(n + bs - 1) // bs for n in kv_lens]
    nb = sum(blocks) + 5
    mb = max(blocks) + 2
    kc = torch.randn(nb, Hkv, bs, D, device=DEV, generator=g).to(dtype)
    vc = torch.randn(nb, Hkv, bs, D, device=DEV, generator=g).to(dtype)
    perm = torch.randperm(nb, device=DEV, generator=g).to(torch.int32)
    table = torch.full((B, mb), -1, dtype=torch.int32, device=DEV)
    pos = 0
    for i, n in enumerate(blocks):
        table[i, :n] = perm[pos:pos + n]
        pos += n
    T = sum(q_lens)
    q = torch.randn(T, Hq, D, device=DEV, generator=g).to(dtype)
    cu_q = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32, device=DEV)
    cu_kv = torch.tensor([0] + list(torch.tensor(kv_lens).cumsum(0)), dtype=torch.int32, device=DEV)
    G = Hq // Hkv
    ref = torch.zeros(T, Hq, D, device=DEV)
    for i in range(B):
        if q_lens[i] == 0:
            continue
        ids = table[i, :blocks[i]].long()
        kk = kc[ids].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :kv_lens[i]]
        vv = vc[ids].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :kv_lens[i]]
        if layout == "AABB":
            kk, vv = kk.repeat_interleave(G, 0), vv.repeat_interleave(G, 0)
        else:
            kk, vv = kk.repeat(G, 1, 1), vv.repeat(G, 1, 1)
        qq = q[cu_q[i]:cu_q[i + 1]].transpose(0, 1)
        ref[cu_q[i]:cu_q[i + 1]] = ref_attn(qq, kk, vv, 1 / math.sqrt(D), causal_off=prefix_lens[i]).transpose(0, 1)
    out = {}
    for impl in ("mma", "tcgen05"):
        set_impl(impl)
        try:
            o = F.paged_prefill_gqa(q, kc, vc, cu_q, table, None, cu_kv, layout, max(q_lens), max(kv_lens))
            torch.cuda.synchronize()
            out[impl] = (o.float() - ref).abs().max().item()
        except Exception as e:  # noqa: BLE001
            out[impl] = f"ERR {type(e).__name__}: {e}"
    set_impl(None)
    print(f"prefill q{q_lens} prefix{prefix_lens} Hq{Hq} Hkv{Hkv} bs{bs} {layout} {str(dtype)[6:]}: max|err| " +
          " ".join(f"{k_}={v_ if isinstance(v_, str) else format(v_, '.4f')}" for k_, v_ in out.items()), flush=True)
    return out


def time_it(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def bench(impls=("mma", "tcgen05")):
    Hq, Hkv, D, bs, T = 32, 8, 128, 16, 8192
    nb = T // bs + 10
    kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
    vc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
    q = torch.empty(T, Hq, D, dtype=torch.bfloat16, device=DEV).normal_()
    table = torch.randperm(nb, device=DEV)[: T // bs].view(1, -1).to(torch.int32)
    cu = torch.tensor([0, T], dtype=torch.int32, device=DEV)
    flops = 4 * Hq * D * (T * (T + 1) // 2)
    for impl in impls:
        set_impl(impl)
        ms = time_it(lambda: F.paged_prefill_gqa(q, kc, vc, cu, table, None, None, "AABB", T, T))
        print(f"cfg3 prefill T=8192 causal [{impl}]: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    Bd, H, S = 2, 24, 4096
    qs, ks, vs = (torch.empty(Bd, S, H, D, dtype=torch.bfloat16, device=DEV).normal_().transpose(1, 2) for _ in range(3))
    flops = 4 * Bd * H * S * S * D
    for impl in impls:
        set_impl(impl)
        ms = time_it(lambda: F.sdpa(qs, ks, vs))
        print(f"cfg5 sdpa B2 H24 S4096 [{impl}]: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    set_impl(None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-bench", action="store_true")
    ap.add_argument("--only", default=None)
    ap.add_argument("--bench-impl", default=None)
    args = ap.parse_args()
    if args.only in (None, "sdpa"):
        check_sdpa(1, 1, 1, 128, 128)
        check_sdpa(1, 1, 1, 256, 256)
        check_sdpa(1, 2, 2, 256, 512)
        check_sdpa(2, 3, 3, 1024, 1024)
        check_sdpa(1, 4, 2, 300, 200)
        check_sdpa(1, 2, 2, 4096, 512)
        check_sdpa(1, 2, 1, 777, 1555, dtype=torch.float16)
    if args.only in (None, "prefill"):
        check_prefill([256], [0], 4, 4, 128)
        check_prefill([256], [0], 4, 1, 16)
        check_prefill([512, 300], [0, 0], 8, 2, 16)
        check_prefill([200, 1000, 77], [512, 0, 33], 8, 2, 32)
        check_prefill([640], [1024], 4, 1, 1024)
        check_prefill([384, 0, 129], [100, 50, 7], 4, 2, 16, layout="ABAB")
        check_prefill([333], [95], 2, 2, 8, dtype=torch.float16)
    if not args.no_bench:
        bench((args.bench_impl,) if args.bench_impl else ("mma", "tcgen05"))


if __name__ == "__main__":
    main()
