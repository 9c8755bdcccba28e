#!/usr/bin/env python
"""cfg4 (BASELINE.json): Llama-3-70B-shaped paged decode, 64 q / 8 kv heads, hd 128, batch 256, ctx 32k, bf16,
KV heads sharded TP = world size, one NCCL all-reduce on the o_proj output per layer.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_tp_decode.py --steps 10 --warmup 3 [--ctx 32768] [--batch 256]

Per step and rank: MojoStorePagedKVCache (new token) -> MojoPagedDecodeGQA on the local head shard -> o_proj GEMM on the
local columns (cuBLAS, plumbing) -> all-reduce [batch, 8192] bf16.  Device-timed with CUDA events, max over ranks.
Prints one JSON line on rank 0: tokens/s, per-rank decode GB/s vs the measured HBM peak, all-reduce time.
"""

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ctx", type=int, default=32768)
    ap.add_argument("--hq", type=int, default=64)
    ap.add_argument("--hkv", type=int, default=8)
    ap.add_argument("--hidden", type=int, default=8192)
    ap.add_argument("--layers", type=int, default=2, help="distinct KV caches rotated across steps (cold KV)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    os.environ["MOJO_BACKEND"] = "b200"
    import mojo_opset_b200 as m
    from mojo_opset_b200.parallel import RowParallelOutProj
    from mojo_opset_b200.parallel import shard_heads

    B, D, bs, ctx = args.batch, 128, 16, args.ctx
    shard = shard_heads(args.hq, args.hkv, world, rank)
    hq_l, hkv_l = shard.q_end - shard.q_begin, shard.kv_end - shard.kv_begin
    blocks_per_seq = ctx // bs
    nb = B * blocks_per_seq + 8
    g = torch.Generator().manual_seed(20260716 + 4)
    caches, tables, metas = [], [], []
    for _ in range(args.layers):
        kc = torch.empty(nb, hkv_l, bs, D, dtype=torch.bfloat16, device=dev).normal_()
        vc = torch.empty(nb, hkv_l, bs, D, dtype=torch.bfloat16, device=dev).normal_()
        perm = torch.randperm(nb, generator=g)[: B * blocks_per_seq].view(B, blocks_per_seq).to(torch.int32)
        meta = torch.stack((torch.arange(B, dtype=torch.int32), perm[:, (ctx - 1) // bs],
                            torch.full((B,), (ctx - 1) % bs, dtype=torch.int32), torch.ones(B, dtype=torch.int32)), -1)
        caches.append((kc, vc))
        tables.append(perm.to(dev))
        metas.append(meta.contiguous().to(dev))
    q = torch.randn(B, hq_l, D, generator=g).to(torch.bfloat16).to(dev)
    k_new = torch.randn(B, hkv_l, D, generator=g).to(torch.bfloat16).to(dev)
    v_new = torch.randn(B, hkv_l, D, generator=g).to(torch.bfloat16).to(dev)
    lens = torch.full((B,), ctx, dtype=torch.int32, device=dev)
    w_o = (torch.randn(args.hidden, args.hq * D, generator=g) / (args.hq * D) ** 0.5).to(torch.bfloat16).to(dev)
    o_proj = RowParallelOutProj(w_o, shard, D).to(dev)
    del w_o
    store, decode = m.MojoStorePagedKVCache(), m.MojoPagedDecodeGQA()
    assert type(decode).__name__ == "B200PagedDecodeGQA"

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    dec_ev, ar_ev = [], []

    def step(i, timed):
        L = i % args.layers
        kc, vc = caches[L]
        store(k_new, v_new, kc, vc, chunk_metadata=metas[L])
        e = [ev() for _ in range(4)] if timed else None
        if timed:
            e[0].record()
        o = decode(q, kc, vc, lens, tables[L], max_total_seq_len=ctx)
        if timed:
            e[1].record()
        y = torch.nn.functional.linear(o.reshape(B, -1), o_proj.weight)
        if timed:
            e[2].record()
        if world > 1:
            dist.all_reduce(y)
        if timed:
            e[3].record()
            dec_ev.append((e[0], e[1]))
            ar_ev.append((e[2], e[3]))
        return y

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i, False)
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(args.steps):
        step(args.warmup + i, True)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    dec_ms = sum(a.elapsed_time(b) for a, b in dec_ev) / len(dec_ev)
    ar_ms = sum(a.elapsed_time(b) for a, b in ar_ev) / len(ar_ev)
    stats = torch.tensor([ms, dec_ms, ar_ms], device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms, dec_ms, ar_ms = stats.tolist()
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
        kv_bytes = 2 * B * ctx * hkv_l * D * 2 + 2 * B * hq_l * D * 2 + B * blocks_per_seq * 4 + B * 4
        gbs = kv_bytes / (dec_ms * 1e-3) / 1e9
        print(json.dumps({
            "workload": f"cfg4 Llama-3-70B-shaped decode layer: batch {B}, {args.hq}q/{args.hkv}kv heads, hd 128, ctx {ctx}, "
                        f"page 16, bf16, TP={world} over KV heads + o_proj all-reduce [{B},{args.hidden}] bf16",
            "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps,
            "tokens_per_s": B * args.steps / (ms * 1e-3),
            "decode_us_max_rank": dec_ms * 1e3, "decode_bytes_per_rank": kv_bytes, "decode_gbs_per_rank": gbs,
            "decode_frac_of_measured_hbm": gbs / peaks["hbm_gbs"], "allreduce_us_max_rank": ar_ms * 1e3,
            "allreduce_bytes": B * args.hidden * 2, "local_heads": [hq_l, hkv_l],
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
