"""-m gpu: programmatic dependent launch (csrc/common.cuh `launch_pdl`) along the decode layer chain.

The chain's kernels (ResidualAdd+RMSNorm -> RoPE -> StorePagedKVCache -> PagedDecodeGQA + split fold -> SwiGLU) are
launched with `cudaLaunchAttributeProgrammaticStreamSerialization`: a kernel's CTAs may become resident while its
predecessor drains and every kernel waits (`griddepcontrol.wait`) before it touches memory.  A missing wait is a race,
so the chain is replayed many times from ONE captured CUDA graph (programmatic edges in the graph) and back to back
eagerly; every replay must reproduce, bit for bit, the result of the same chain launched with MOJO_B200_PDL=0 (plain
stream order).  The decode runs with forced splits, so the fold kernel's wait on the partials is exercised too."""

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


def _chain(ops, B=24, ctx=700, Hq=32, Hkv=8, D=128, bs=16, hidden=1024, inter=3072, seed=3):
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(DEV)  # noqa: E731
    nblk = (ctx + bs - 1) // bs
    nb = B * nblk + 3
    d = dict(hidden=rnd(B, hidden), residual=rnd(B, hidden), q=rnd(B, Hq, D), k=rnd(B, Hkv, D), v=rnd(B, Hkv, D),
             gate=rnd(B, inter), up=rnd(B, inter), kc=rnd(nb, Hkv, bs, D), vc=rnd(nb, Hkv, bs, D),
             cos=torch.randn(B, D, generator=g).to(DEV), sin=torch.randn(B, D, generator=g).to(DEV))
    table = torch.randperm(nb, generator=g)[: B * nblk].view(B, nblk).to(torch.int32)
    meta = torch.stack((torch.arange(B, dtype=torch.int32), table[:, (ctx - 1) // bs],
                        torch.full((B,), (ctx - 1) % bs, dtype=torch.int32), torch.ones(B, dtype=torch.int32)), -1)
    d["table"], d["meta"] = table.to(DEV), meta.contiguous().to(DEV)
    d["lens"] = torch.full((B,), ctx, dtype=torch.int32, device=DEV)
    norm = ops.MojoResidualAddRMSNorm(hidden, eps=1e-6, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():  # (the op allocates its weight like the reference does: uninitialised memory)
        norm.weight.copy_(1 + 0.1 * torch.randn(hidden, generator=g))
    rope, store, decode, swiglu = ops.MojoApplyRoPE(), ops.MojoStorePagedKVCache(), ops.MojoPagedDecodeGQA(), ops.MojoSwiGLU()

    def step():
        y, r = norm(d["hidden"], d["residual"])
        q_rot, k_rot = rope(d["q"], d["k"], d["cos"], d["sin"], head_first=False)
        store(k_rot, d["v"], d["kc"], d["vc"], chunk_metadata=d["meta"])
        o = decode(q_rot, d["kc"], d["vc"], d["lens"], d["table"], max_total_seq_len=ctx)
        a = swiglu(d["gate"], d["up"])
        return y, r, o, a

    return step


def test_chain_with_pdl_equals_plain_stream_order(ops, monkeypatch):
    monkeypatch.setenv("MOJO_B200_DECODE_SPLITS", "3")  # main kernel + fold kernel
    step = _chain(ops)
    monkeypatch.setenv("MOJO_B200_PDL", "0")
    ref = [t.clone() for t in step()]
    torch.cuda.synchronize()
    monkeypatch.setenv("MOJO_B200_PDL", "1")
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = [step() for _ in range(4)]  # four chained steps in ONE graph: the next step's norm follows this SwiGLU
    for _ in range(25):
        graph.replay()
    torch.cuda.synchronize()
    for k, got in enumerate(outs):
        for name, a, b in zip("yroa", got, ref):
            assert torch.equal(a, b), (f"step {k} of the replayed graph, output {name}: {int((a != b).sum())} of {a.numel()} "
                                       "elements differ from plain stream order")
    for _ in range(25):  # eager, back to back
        got = step()
    torch.cuda.synchronize()
    for name, a, b in zip("yroa", got, ref):
        assert torch.equal(a, b), f"eager, output {name}: {int((a != b).sum())} of {a.numel()} elements differ"
