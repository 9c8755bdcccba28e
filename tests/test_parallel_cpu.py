"""World-size-2 ``gloo`` tests (CPU) of the N>1 path: data-parallel sequence sharding (no collective on the data
path) and KV-head tensor parallelism with the single o_proj all-reduce (reference
``mojo_opset/tests/distributed/test_paged_gqa_tp.py:348-416`` compares a TP=2 block with the single-rank block the
same way).  The per-rank attention runs the oracle (CPU); what is under test is ``mojo_opset_b200.parallel``."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mojo_opset_b200.parallel import RowParallelOutProj
from mojo_opset_b200.parallel import shard_heads
from mojo_opset_b200.parallel import shard_range


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_shard_heads_matches_reference_rule():
    # Llama-3-70B-shaped: 64 q / 8 kv heads (cfg4) at TP 2/4/8
    for tp in (2, 4, 8):
        seen_q, seen_kv = [], []
        for r in range(tp):
            s = shard_heads(64, 8, tp, r)
            assert s.q_end - s.q_begin == 64 // tp and s.kv_end - s.kv_begin == 8 // tp and s.kv_replicas == 1
            # AABB: a rank's q heads map onto exactly its own kv heads
            assert s.q_begin // 8 == s.kv_begin and (s.q_end - 1) // 8 == s.kv_end - 1
            seen_q += range(s.q_begin, s.q_end)
            seen_kv += range(s.kv_begin, s.kv_end)
        assert seen_q == list(range(64)) and seen_kv == list(range(8))
    # tp > Hkv: KV heads replicated over tp / Hkv consecutive ranks (reference partitions.py:147-150)
    shards = [shard_heads(32, 2, 4, r) for r in range(4)]
    assert [s.kv_begin for s in shards] == [0, 0, 1, 1] and all(s.kv_replicas == 2 for s in shards)
    with pytest.raises(ValueError):
        shard_heads(30, 8, 4, 0)
    with pytest.raises(ValueError):
        shard_heads(32, 6, 4, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_case(seed=0, B=6, Hq=8, Hkv=2, D=32, bs=16, ctx=70, hidden=48):
    g = torch.Generator().manual_seed(seed)
    blocks = (ctx + bs - 1) // bs
    nb = B * blocks + 3
    kc = torch.randn(nb, Hkv, bs, D, generator=g)
    vc = torch.randn(nb, Hkv, bs, D, generator=g)
    table = torch.randperm(nb, generator=g)[: B * blocks].view(B, blocks).to(torch.int32)
    lens = torch.randint(1, ctx + 1, (B,), generator=g).to(torch.int32)
    q = torch.randn(B, Hq, D, generator=g)
    w_o = torch.randn(hidden, Hq * D, generator=g) / (Hq * D) ** 0.5
    return q, kc, vc, lens, table, w_o


def _worker(rank, world, port, mode, out_queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import golden

        q, kc, vc, lens, table, w_o = _make_case()
        B, Hq, D = q.shape
        Hkv = kc.shape[1]
        full_attn = golden.paged_decode_gqa(q, kc, vc, lens, table)
        full = torch.nn.functional.linear(full_attn.reshape(B, -1), w_o)
        if mode == "tp":
            s = shard_heads(Hq, Hkv, world, rank)
            local = golden.paged_decode_gqa(q[:, s.q_begin:s.q_end].contiguous(), kc[:, s.kv_begin:s.kv_end].contiguous(),
                                            vc[:, s.kv_begin:s.kv_end].contiguous(), lens, table)
            y = RowParallelOutProj(w_o, s, D)(local)  # the one collective on the path
            ok = torch.allclose(y, full, atol=1e-4, rtol=1e-4)
        elif mode == "gemm_allreduce":
            # MojoGemmAllReduce through the registry with the oracle backend under gloo: the o_proj of the TP block as
            # the reference's fused op (core/operators/compute_with_comm.py:57-117), bias added on every rank
            import mojo_opset_b200 as mj
            import oracle.torch_backend  # noqa: F401  registers backend "torch"

            os.environ["MOJO_BACKEND"] = "torch"
            s = shard_heads(Hq, Hkv, world, rank)
            w_local = w_o[:, s.q_begin * D: s.q_end * D].contiguous()
            bias = torch.arange(w_o.shape[0], dtype=torch.float32) / 10
            x_local = full_attn[:, s.q_begin:s.q_end].reshape(B, -1)
            op = mj.MojoGemmAllReduce(w_local, bias, trans_weight=False)
            op_t = mj.MojoGemmAllReduce(w_local.t().contiguous(), bias, trans_weight=True)
            y, y_t = op(x_local), op_t(x_local)
            ok = (type(op).__name__ == "TorchGemmAllReduce" and torch.allclose(y, full + world * bias, atol=1e-4, rtol=1e-4)
                  and torch.allclose(y_t, y, atol=1e-5, rtol=1e-5))
            emu = golden.gemm_allreduce_emulated(
                [full_attn[:, shard_heads(Hq, Hkv, world, r).q_begin:shard_heads(Hq, Hkv, world, r).q_end].reshape(B, -1)
                 for r in range(world)],
                [w_o[:, shard_heads(Hq, Hkv, world, r).q_begin * D: shard_heads(Hq, Hkv, world, r).q_end * D]
                 for r in range(world)], [bias] * world)
            ok = ok and torch.allclose(emu, y, atol=1e-5, rtol=1e-5)
        else:  # data parallel: every rank owns a contiguous share of the sequences, no collective on the data path
            b0, b1 = shard_range(B, world, rank)
            local = golden.paged_decode_gqa(q[b0:b1], kc, vc, lens[b0:b1], table[b0:b1])
            gathered = [None] * world
            dist.all_gather_object(gathered, (b0, b1, local))  # test-side only: collect for the comparison
            rebuilt = torch.cat([t for _, _, t in sorted(gathered, key=lambda x: x[0])])
            ok = torch.equal(rebuilt, full_attn) and sum(e - b for b, e, _ in gathered) == B
        out_queue.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["tp", "dp", "gemm_allreduce"])
def test_world_size_2_gloo(mode):
    world = 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
