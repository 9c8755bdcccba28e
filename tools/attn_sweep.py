#!/usr/bin/env python
"""Developer sweep: time cfg3 paged prefill / cfg5 SDPA under a list of environment-variable variants of the
tcgen05 attention kernel (the library reads them at launch time).  Not part of the product path.

    python tools/attn_sweep.py "MOJO_B200_ATTN_EMU=1" "MOJO_B200_ATTN_EMU=2,MOJO_B200_ATTN_NOROUND=1" ...
"""

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
from mojo_opset_b200 import functional as F  # noqa: E402

DEV = "cuda"


def time_ms(fn, iters=10, warmup=3):
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


def main():
    variants = [""] + sys.argv[1:]
    g = torch.Generator().manual_seed(3)
    Hq, Hkv, D, bs = 32, 8, 128, 16
    cases = {}
    for name, T, ctx in (("prefill8k", 8192, 0), ("chunk8k+8k", 8192, 8192), ("prefill2k x4", 2048, 0), ("prefill4k", 4096, 0), ("prefill16k", 16384, 0)):
        B = 4 if "x4" in name else 1
        kv = T + ctx
        nb = B * kv // bs + 10
        kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
        vc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
        q = torch.empty(B * T, Hq, D, dtype=torch.bfloat16, device=DEV).normal_()
        table = torch.randperm(nb, generator=g)[: B * kv // bs].view(B, -1).to(torch.int32).to(DEV)
        cu_q = (torch.arange(B + 1, dtype=torch.int32) * T).to(DEV)
        cu_kv = (torch.arange(B + 1, dtype=torch.int32) * kv).to(DEV)
        flops = B * 4 * Hq * D * sum(ctx + t + 1 for t in range(T))
        cases[name] = (lambda q=q, kc=kc, vc=vc, cu_q=cu_q, table=table, cu_kv=cu_kv, T=T, kv=kv:
                       F.paged_prefill_gqa(q, kc, vc, cu_q, table, None, cu_kv, max_q_len=T, max_total_seq_len=kv),
                       flops)
    for name, Bd in (("sdpa b1", 1), ("sdpa b2", 2), ("sdpa b4", 4), ("sdpa b16", 16)):
        H, S = 24, 4096
        qs, ks, vs = (torch.empty(Bd, S, H, D, dtype=torch.bfloat16, device=DEV).normal_().transpose(1, 2)
                      for _ in range(3))
        cases[name] = (lambda qs=qs, ks=ks, vs=vs: F.sdpa(qs, ks, vs), 4 * Bd * H * S * S * D)

    for name, Bd in (("sdpa d64 b2", 2), ("sdpa d64 b8", 8)):
        H, S = 24, 4096
        qs, ks, vs = (torch.empty(Bd, S, H, 64, dtype=torch.bfloat16, device=DEV).normal_().transpose(1, 2)
                      for _ in range(3))
        cases[name] = (lambda qs=qs, ks=ks, vs=vs: F.sdpa(qs, ks, vs), 4 * Bd * H * S * S * 64)

    # round-robin over the variants, several rounds, so that clock / power drift hits every variant alike;
    # report min and median over the rounds
    import statistics
    import time
    results = {var: {name: [] for name in cases} for var in variants}
    rounds = int(os.environ.get("SWEEP_ROUNDS", 5))
    for _ in range(rounds):
        for name, (fn, flops) in cases.items():
            for var in variants:
                keys = []
                for kv_ in filter(None, var.split(",")):
                    k, v = kv_.split("=")
                    os.environ[k] = v
                    keys.append(k)
                try:
                    results[var][name].append(time_ms(fn, iters=4, warmup=1))
                except Exception as e:  # noqa: BLE001
                    print(f"{var} {name}: ERR {e}")
                for k in keys:
                    os.environ.pop(k, None)
                time.sleep(0.05)
    for var in variants:
        row = []
        for name, (fn, flops) in cases.items():
            r = results[var][name]
            if r:
                row.append(f"{name}: {min(r) * 1e3:7.1f}/{statistics.median(r) * 1e3:7.1f} us "
                           f"{flops / min(r) / 1e9:6.0f}/{flops / statistics.median(r) / 1e9:6.0f} TF/s")
        print(f"[{var or 'default'}]  " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()
