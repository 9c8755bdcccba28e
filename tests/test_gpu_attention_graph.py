"""-m gpu: the attention ops inside CUDA graphs, the way the reference's graph tests use them
(``tests/accuracy/operators/test_attention_cudagraph.py`` and ``test_attention.py:196-353``):
capture once with max-shape static buffers, then overwrite query / caches / lengths / block tables IN PLACE - padding
rows get ``seq_len = 0`` and block ids ``-1`` - and replay.  The replayed outputs must match the oracle for the new
contents and padding rows must read as zeros.  Covers the split-KV decode kernel, its windowed variant, the tcgen05
prefill kernel launched as 2-CTA clusters and the DiT SDPA."""

import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = dict(atol=2e-2, rtol=2e-2)


@pytest.fixture(scope="module")
def F():
    os.environ["MOJO_BACKEND"] = "b200"
    from mojo_opset_b200 import functional

    return functional


def _capture(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):  # warm-up outside the capture (library load, tensor-map cache)
            fn()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = fn()
    return graph, out


def _fill_tables(table, lens, bs, perm):
    table.fill_(-1)
    at = 0
    for i, n in enumerate(lens):
        need = (n + bs - 1) // bs
        table[i, :need] = perm[at:at + need]
        at += need


@pytest.mark.parametrize("window", [None, (200, 30)])
def test_decode_graph_replay(F, window):
    from oracle import golden

    g = torch.Generator().manual_seed(5)
    B, Hq, Hkv, D, bs, max_len = 6, 8, 2, 128, 16, 1024
    nb = B * max_len // bs + 4
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16).to(DEV)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16).to(DEV)
    q = torch.randn(B, Hq, D, generator=g).to(torch.bfloat16).to(DEV)
    lens = torch.zeros(B, dtype=torch.int32, device=DEV)
    table = torch.full((B, max_len // bs), -1, dtype=torch.int32, device=DEV)
    local, glob = window if window else (None, None)

    def step():  # max_total_seq_len = the static bound: the launch shape must not depend on the data
        return F.paged_decode_swa(q, kc, vc, lens, table, None, "AABB", max_len, local, glob)

    host_lens = [1000, 37, 512, 3, 700, 64]
    perm = torch.randperm(nb, generator=g).to(torch.int32)
    host_table = torch.empty(B, max_len // bs, dtype=torch.int32)
    _fill_tables(host_table, host_lens, bs, perm)
    table.copy_(host_table)
    lens.copy_(torch.tensor(host_lens, dtype=torch.int32))
    graph, out = _capture(step)
    for host_lens in ([1024, 0, 9, 333, 0, 801], [5, 5, 0, 0, 1024, 64]):  # padding rows: length 0, ids -1
        perm = torch.randperm(nb, generator=g).to(torch.int32)
        _fill_tables(host_table, host_lens, bs, perm)
        table.copy_(host_table)
        lens.copy_(torch.tensor(host_lens, dtype=torch.int32))
        q.copy_(torch.randn(B, Hq, D, generator=g).to(torch.bfloat16))
        kc.copy_(torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16))
        graph.replay()
        torch.cuda.synchronize()
        ref = golden.paged_decode_swa(q.cpu(), kc.cpu(), vc.cpu(), lens.cpu(), table.cpu(), None, "AABB", local, glob)
        torch.testing.assert_close(out.cpu().float(), ref.float(), **TOL)
        for i, n in enumerate(host_lens):
            if n == 0:
                assert torch.count_nonzero(out[i]).item() == 0


def test_prefill_and_sdpa_graph_replay(F):
    from oracle import golden

    g = torch.Generator().manual_seed(6)
    Hq, Hkv, D, bs = 4, 2, 128, 16
    B, max_q, max_kv = 3, 512, 768
    T = B * max_q
    nb = B * max_kv // bs + 4
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16).to(DEV)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16).to(DEV)
    q = torch.randn(T, Hq, D, generator=g).to(torch.bfloat16).to(DEV)
    cu_q = torch.zeros(B + 1, dtype=torch.int32, device=DEV)
    cu_kv = torch.zeros(B + 1, dtype=torch.int32, device=DEV)
    table = torch.full((B, max_kv // bs), -1, dtype=torch.int32, device=DEV)
    qs, ks, vs = (torch.randn(2, 300, 3, D, generator=g).to(torch.bfloat16).to(DEV).transpose(1, 2) for _ in range(3))

    def step():
        return (F.paged_prefill_gqa(q, kc, vc, cu_q, table, None, cu_kv, "AABB", max_q, max_kv), F.sdpa(qs, ks, vs))

    def load(q_lens, prefix):
        kv = [a + b for a, b in zip(q_lens, prefix)]
        host_table = torch.empty(B, max_kv // bs, dtype=torch.int32)
        _fill_tables(host_table, kv, bs, torch.randperm(nb, generator=g).to(torch.int32))
        table.copy_(host_table)
        cu_q.copy_(torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32))
        cu_kv.copy_(torch.tensor([0] + torch.tensor(kv).cumsum(0).tolist(), dtype=torch.int32))

    load([512, 300, 200], [256, 0, 100])
    graph, (out, out_s) = _capture(step)
    for q_lens, prefix in (([256, 0, 512], [512, 0, 17]), ([33, 400, 1], [0, 368, 700])):  # row 2 of case 1: padding
        load(q_lens, prefix)
        q.copy_(torch.randn(T, Hq, D, generator=g).to(torch.bfloat16))
        qs.copy_(torch.randn(2, 3, 300, D, generator=g).to(torch.bfloat16))
        graph.replay()
        torch.cuda.synchronize()
        ref = golden.paged_prefill_gqa(q.cpu(), kc.cpu(), vc.cpu(), cu_q.cpu(), table.cpu(), None, cu_kv.cpu(), "AABB")
        n = sum(q_lens)
        torch.testing.assert_close(out[:n].cpu().float(), ref[:n].float(), **TOL)
        assert torch.count_nonzero(out[n:]).item() == 0  # tokens past cu_q_lens[-1] read as zeros
        torch.testing.assert_close(out_s.cpu().float(), golden.sdpa(qs.cpu(), ks.cpu(), vs.cpu()).float(), **TOL)
