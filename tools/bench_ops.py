#!/usr/bin/env python
"""Roofline table of every HBM-bound op of the path at its prefill-sized and decode-sized shapes (SURVEY.md 8d
algorithmic bytes / CUDA-event time vs the HBM peak), plus paged decode at the cfg4 per-GPU slices and a varlen
batch.  Buffers rotate over several copies so that every launch streams cold data (working set > L2).

    python tools/bench_ops.py [--json gpurun_out/ops.json]
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
BF = torch.bfloat16


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_us(fns, iters=20, warmup=4):
    """Device time per call: the calls are captured into ONE CUDA graph (which also proves every op is graph
    capturable: no sync, no host read) and the graph is replayed, so Python/ctypes launch overhead is excluded.
    Returns (graph_us, eager_us); eager = back-to-back Python calls, host overhead included."""
    n = len(fns)
    for i in range(warmup):
        fns[i % n]()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fns[i % n]()
    b.record()
    torch.cuda.synchronize()
    eager = a.elapsed_time(b) / iters * 1e3
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(iters):
                fns[i % n]()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters * 1e3)
    del graph
    return best, eager


def rnd(*shape, dtype=BF):
    return torch.empty(*shape, dtype=dtype, device=DEV).normal_()


def copies_for(bytes_per_call):
    return max(2, min(8, int(400e6 // max(bytes_per_call, 1)) + 1))


def bench_elementwise(rows_list):
    out = []
    H, I, Hq, Hkv, D, bs = 4096, 12288, 32, 8, 128, 16
    for T in rows_list:
        # ResidualAdd + RMSNorm: 4*T*H*2 + H*2
        nbytes = 4 * T * H * 2 + H * 2
        n = copies_for(nbytes)
        w = rnd(H)
        sets = [(rnd(T, H), rnd(T, H)) for _ in range(n)]
        us = time_us([lambda s=s: F.residual_add_rms_norm(s[0], s[1], w, 1e-6) for s in sets])
        out.append(("residual_add_rms_norm", f"{T}x{H}", nbytes, us))
        nbytes = 2 * T * H * 2 + H * 2
        us = time_us([lambda s=s: F.rms_norm(s[0], w, 1e-6) for s in sets])
        out.append(("rms_norm", f"{T}x{H}", nbytes, us))
        # q/k head norm (D = 128 rows)
        sets_h = [rnd(T, Hq, D) for _ in range(n)]
        wd = rnd(D)
        nbytes = 2 * T * Hq * D * 2 + D * 2
        us = time_us([lambda s=s: F.rms_norm(s, wd, 1e-6) for s in sets_h])
        out.append(("rms_norm(head)", f"{T * Hq}x{D}", nbytes, us))
        del sets
        # RoPE: 2*T*(Hq+Hkv)*D*2 + 2*T*D*4
        nbytes = 2 * T * (Hq + Hkv) * D * 2 + 2 * T * D * 4
        n = copies_for(nbytes)
        sets = [(rnd(T, Hq, D), rnd(T, Hkv, D), rnd(T, D, dtype=torch.float32), rnd(T, D, dtype=torch.float32))
                for _ in range(n)]
        us = time_us([lambda s=s: F.apply_rope(s[0], s[1], s[2], s[3], head_first=False) for s in sets])
        out.append(("apply_rope", f"T={T} {Hq}+{Hkv} heads", nbytes, us))
        del sets
        # SwiGLU: 3*T*I*2
        nbytes = 3 * T * I * 2
        n = copies_for(nbytes)
        sets = [(rnd(T, I), rnd(T, I)) for _ in range(n)]
        us = time_us([lambda s=s: F.swiglu(s[0], s[1]) for s in sets])
        out.append(("swiglu", f"{T}x{I}", nbytes, us))
        del sets
        # Store KV: one sequence of T new tokens (prefill) or T sequences of one token (decode)
        prefill = T >= 1024
        chunks = T // bs if prefill else T
        nbytes = 2 * (2 * T * Hkv * D * 2) + 16 * chunks
        n = copies_for(nbytes)
        nb = (T // bs if prefill else T) + 10
        sets = []
        for _ in range(n):
            perm = torch.randperm(nb)[: nb - 10].to(torch.int32)
            if prefill:
                meta = torch.stack((torch.arange(chunks, dtype=torch.int32) * bs, perm[:chunks],
                                    torch.zeros(chunks, dtype=torch.int32),
                                    torch.full((chunks,), bs, dtype=torch.int32)), -1)
            else:
                meta = torch.stack((torch.arange(T, dtype=torch.int32), perm[:T],
                                    torch.full((T,), 5, dtype=torch.int32), torch.ones(T, dtype=torch.int32)), -1)
            sets.append((rnd(T, Hkv, D), rnd(T, Hkv, D), rnd(nb, Hkv, bs, D), rnd(nb, Hkv, bs, D),
                         meta.contiguous().to(DEV)))
        us = time_us([lambda s=s: F.store_paged_kv(s[0], s[1], s[2], s[3], chunk_metadata=s[4]) for s in sets])
        out.append(("store_paged_kv", f"T={T} chunks={chunks}", nbytes, us))
        del sets
        # DiT block ops: LayerNorm (2 streams + affine), exact GELU (2 streams)
        nbytes = 2 * T * H * 2 + 2 * H * 2
        n = copies_for(nbytes)
        sets = [rnd(T, H) for _ in range(n)]
        lw, lb = rnd(H), rnd(H)
        us = time_us([lambda s=s: F.layer_norm(s, lw, lb, 1e-6) for s in sets])
        out.append(("layer_norm", f"{T}x{H}", nbytes, us))
        del sets
        nbytes = 2 * T * I * 2
        n = copies_for(nbytes)
        sets = [rnd(T, I) for _ in range(n)]
        us = time_us([lambda s=s: F.gelu(s) for s in sets])
        out.append(("gelu", f"{T}x{I}", nbytes, us))
        del sets
        # fused q/k-norm + RoPE + KV store (one pass): reads q, k, v, writes q' and the k', v slots
        nbytes = (2 * Hq + 4 * Hkv) * D * 2 * T + 2 * T * D * 4 + 2 * D * 2
        n = copies_for(nbytes)
        sets = []
        for _ in range(n):
            if prefill:
                table = torch.randperm(nb)[: T // bs].view(1, -1).to(torch.int32).to(DEV)
                cu = torch.tensor([0, T], dtype=torch.int32, device=DEV)
                ctx = torch.zeros(1, dtype=torch.int32, device=DEV)
            else:
                table = torch.randperm(nb)[:T].view(T, 1).to(torch.int32).to(DEV)
                cu, ctx = None, torch.full((T,), 5, dtype=torch.int32, device=DEV)
            sets.append((rnd(T, Hq, D), rnd(T, Hkv, D), rnd(T, Hkv, D), rnd(T, D, dtype=torch.float32),
                         rnd(T, D, dtype=torch.float32), rnd(nb, Hkv, bs, D), rnd(nb, Hkv, bs, D), table, cu, ctx))
        wq, wk = rnd(D), rnd(D)
        us = time_us([lambda s=s: F.norm_rope_store_kv(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], s[9],
                                                        wq, wk, 1e-6) for s in sets])
        out.append(("norm_rope_store_kv (fused)", f"T={T} {Hq}q+{Hkv}k+{Hkv}v heads", nbytes, us))
        del sets
        torch.cuda.empty_cache()
    return out


def bench_decode():
    out = []
    D, bs = 128, 16
    g = torch.Generator().manual_seed(7)
    cases = [
        ("cfg2 B64 32q/8kv ctx4096", 64, 32, 8, 4096, None, 3),
        ("cfg2 varlen U[2048,4096]", 64, 32, 8, 4096, "varlen", 3),
        ("cfg4 TP8 slice B256 8q/1kv ctx32768", 256, 8, 1, 32768, None, 2),
        ("cfg4 TP4 slice B256 16q/2kv ctx32768", 256, 16, 2, 32768, None, 2),
        ("cfg4 TP2 slice B256 32q/4kv ctx32768", 256, 32, 4, 32768, None, 1),
        ("small batch B4 32q/8kv ctx8192 (split-KV)", 4, 32, 8, 8192, None, 8),
        ("B1 32q/8kv ctx32768 (split-KV)", 1, 32, 8, 32768, None, 8),
    ]
    for name, B, Hq, Hkv, ctx, mode, n in cases:
        mb = ctx // bs
        nb = B * mb + 10
        sets = []
        lens = torch.full((B,), ctx, dtype=torch.int32)
        if mode == "varlen":
            lens = torch.randint(ctx // 2, ctx + 1, (B,), generator=g, dtype=torch.int32)
        for _ in range(n):
            kc = torch.empty(nb, Hkv, bs, D, dtype=BF, device=DEV).normal_()
            vc = torch.empty(nb, Hkv, bs, D, dtype=BF, device=DEV).normal_()
            table = torch.randperm(nb, generator=g)[: B * mb].view(B, mb).to(torch.int32).to(DEV)
            sets.append((rnd(B, Hq, D), kc, vc, table))
        lens_d = lens.to(DEV)
        nbytes = int(2 * int(lens.sum()) * Hkv * D * 2 + 2 * B * Hq * D * 2
                     + int(((lens + bs - 1) // bs).sum()) * 4 + B * 4)
        us = time_us([lambda s=s: F.paged_decode_gqa(s[0], s[1], s[2], lens_d, s[3], max_total_seq_len=ctx)
                      for s in sets], iters=10, warmup=3)
        out.append(("paged_decode_gqa", name, nbytes, us))
        del sets
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--skip-decode", action="store_true")
    args = ap.parse_args()
    peak, src = peak_hbm()
    rows = bench_elementwise([64, 8192])
    if not args.skip_decode:
        rows += bench_decode()
    table = []
    print(f"HBM peak {peak:.0f} GB/s ({src})")
    print(f"{'op':28s} {'shape':42s} {'MB':>9s} {'us':>9s} {'GB/s':>8s} {'frac':>6s} {'eager us':>9s}")
    for op, shape, nbytes, (us, eager) in rows:
        gbs = nbytes / us / 1e3
        print(f"{op:28s} {shape:42s} {nbytes / 1e6:9.2f} {us:9.1f} {gbs:8.0f} {gbs / peak:6.2f} {eager:9.1f}")
        table.append(dict(op=op, shape=shape, algorithmic_bytes=nbytes, us=us, gbs=gbs, frac=gbs / peak,
                          eager_us=eager))
    if args.json:
        os.makedirs(os.path.dirname(args.json) or ".", exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(dict(peak_gbs=peak, peak_source=src, rows=table), f, indent=1)


if __name__ == "__main__":
    main()
