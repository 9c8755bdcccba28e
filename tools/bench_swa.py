#!/usr/bin/env python
"""Sliding-window attention (MojoPagedPrefillSWA / MojoPagedDecodeSWA) at Qwen3-8B-shaped sizes: time, effective
TFLOP/s over the VISIBLE (query, key) pairs, and the speed-up over full causal attention of the same length (the
window makes the work O(window), the tile skipping must make the time follow).  Developer bench, one GPU.

    python tools/bench_swa.py [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
from mojo_opset_b200 import functional as F  # noqa: E402

DEV = "cuda"


def time_us(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def visible_pairs(q_len, kv_len, local, glob):
    off = kv_len - q_len
    n = 0
    for t in range(q_len):
        pos = off + t
        lo = 0 if local is None else max(0, pos - local)
        if local is None and glob is not None:
            n += min(glob, pos + 1)
        elif glob is None:
            n += pos + 1 - lo
        else:
            n += (pos + 1 - lo) + min(glob, lo)
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    Hq, Hkv, D, bs = 32, 8, 128, 16
    rows = []
    # prefill: one sequence of T tokens
    for T, local, glob in ((8192, None, None), (8192, 4096, None), (8192, 1024, None), (8192, 1024, 64),
                           (32768, 4096, None)):
        nb = T // bs + 10
        kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
        vc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
        q = torch.empty(T, Hq, D, dtype=torch.bfloat16, device=DEV).normal_()
        table = torch.randperm(nb)[: T // bs].view(1, -1).to(torch.int32).to(DEV)
        cu = torch.tensor([0, T], dtype=torch.int32, device=DEV)
        us = time_us(lambda: F.paged_prefill_gqa(q, kc, vc, cu, table, None, cu, "AABB", T, T, True, local, glob))
        pairs = visible_pairs(T, T, local, glob) if (local, glob) != (None, None) else T * (T + 1) // 2
        rows.append(dict(op="prefill", T=T, local=local, glob=glob, us=us, tflops=4 * Hq * D * pairs / us / 1e6))
    # decode: batch 64 at context 32768
    B, ctx = 64, 32768
    nb = B * ctx // bs + 10
    kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
    vc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
    q = torch.empty(B, Hq, D, dtype=torch.bfloat16, device=DEV).normal_()
    table = torch.randperm(nb)[: B * ctx // bs].view(B, -1).to(torch.int32).to(DEV)
    lens = torch.full((B,), ctx, dtype=torch.int32, device=DEV)
    for local, glob in ((None, None), (4096, None), (1024, 64)):
        us = time_us(lambda: F.paged_decode_swa(q, kc, vc, lens, table, None, "AABB", ctx, local, glob))
        keys = ctx if local is None else min(ctx, local + 1 + (glob or 0))
        rows.append(dict(op="decode", B=B, ctx=ctx, local=local, glob=glob, us=us,
                         kv_gbs=2 * B * keys * Hkv * D * 2 / us / 1e3))
    for r in rows:
        print(json.dumps(r))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
