"""-m gpu: MojoGemmAllReduce (csrc/gemm_allreduce.cu) through the C ABI.

* world == 1: the tcgen05 GEMM alone against the oracle (``F.linear`` semantics) over ragged shapes;
* world in {2, 4, 8} on ONE GPU: the full push / reduce / broadcast protocol (self-validating 16-byte lines: three
  payload words + the call's epoch per 128-bit store, polled by the reader) with `world` VIRTUAL ranks in this
  process (``comm.LocalRanks``: one workspace and one stream per rank, plain pointers instead of IPC mappings),
  several back-to-back calls so that both slot parities and the epochs are exercised; shapes with one and with several
  128-row blocks, so both the single-CTA and the CTA-pair (cta_group::2) GEMM run, ragged m / n (rows past m are
  neither sent nor read);
* >= 2 real GPUs (skipped otherwise): tools/bench_gemm_allreduce.py under torchrun, IPC + NVLink.
"""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def F():
    os.environ["MOJO_BACKEND"] = "b200"
    from mojo_opset_b200 import functional

    return functional


@pytest.fixture(scope="module")
def golden():
    from oracle import golden as g  # the checker

    return g


def _tol(k):
    # bf16 output of a length-k dot product of N(0,1) terms: |y| ~ sqrt(k); one bf16 ulp is 2^-8 relative
    return dict(atol=2e-2 * (k ** 0.5), rtol=2e-2)


@pytest.mark.parametrize("m,n,k", [(256, 8192, 1024), (1, 128, 64), (7, 100, 72), (130, 264, 200), (64, 4096, 4096),
                                    (300, 1000, 8), (128, 128, 4104)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("with_bias", [False, True])
@pytest.mark.parametrize("pair", ["0", "1"], ids=["single-cta", "cta-pairs"])  # (the launcher picks by shape; force both)
def test_gemm_single_rank(F, golden, m, n, k, dtype, with_bias, pair, monkeypatch):
    monkeypatch.setenv("MOJO_B200_GAR_PAIR", pair)
    g = torch.Generator().manual_seed(m * 7 + n + k)
    x = torch.randn(m, k, generator=g).to(dtype)
    w = torch.randn(n, k, generator=g).to(dtype)
    b = torch.randn(n, generator=g).to(dtype) if with_bias else None
    out = F.gemm_allreduce(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV))
    ref = golden.gemm(x.float(), w.float(), None if b is None else b.float())
    torch.testing.assert_close(out.cpu().float(), ref, **_tol(k))


def test_gemm_leading_dims_and_strided_rows(F, golden):
    g = torch.Generator().manual_seed(1)
    big = torch.randn(4, 33, 512, generator=g).to(torch.bfloat16).to(DEV)
    x = big[..., :256]  # rows of 256 inside a 512 pitch
    w = torch.randn(320, 256, generator=g).to(torch.bfloat16).to(DEV)
    out = F.gemm_allreduce(x, w)
    assert out.shape == (4, 33, 320)
    ref = golden.gemm(x.float().cpu(), w.float().cpu())
    torch.testing.assert_close(out.cpu().float(), ref, **_tol(256))


@pytest.mark.parametrize("mode", ["one", "two"])  # one-shot (push to all, local reduce) / two-shot (owner reduces)
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("m,n,k_local,dtype", [(256, 1024, 256, torch.bfloat16), (40, 384, 128, torch.bfloat16),
                                               (300, 520, 72, torch.bfloat16), (130, 264, 136, torch.float16)])
@pytest.mark.parametrize("pair", ["0", "1"], ids=["single-cta", "cta-pairs"])
def test_virtual_ranks_protocol(F, golden, world, m, n, k_local, dtype, mode, pair, monkeypatch):
    from mojo_opset_b200.comm import LocalRanks

    monkeypatch.setenv("MOJO_B200_GAR_MODE", mode)
    monkeypatch.setenv("MOJO_B200_GAR_PAIR", pair)
    g = torch.Generator().manual_seed(world * 1000 + m)
    nbytes = F.gemm_allreduce_workspace_bytes(m, n, world)
    ranks = LocalRanks(world, nbytes)
    try:
        for call in range(4):  # both parities twice
            xs = [torch.randn(m, k_local, generator=g).to(dtype) for _ in range(world)]
            ws = [torch.randn(n, k_local, generator=g).to(dtype) for _ in range(world)]
            bs = [torch.randn(n, generator=g).to(dtype) for _ in range(world)]
            ref = golden.gemm_allreduce_emulated(xs, ws, bs)
            xd, wd, bd = ([t.to(DEV) for t in ts] for ts in (xs, ws, bs))
            torch.cuda.synchronize()
            outs = []
            for r in range(world):
                with torch.cuda.stream(ranks.streams[r]):
                    outs.append(F.gemm_allreduce(xd[r], wd[r], bd[r], ranks.view(r), m))
            torch.cuda.synchronize()
            for r in range(world):
                assert torch.equal(outs[r], outs[0]), "ranks must hold bit-identical results"
            torch.testing.assert_close(outs[0].cpu().float(), ref.float(), **_tol(k_local * world))
    finally:
        ranks.close()


def test_op_class_single_process(F, golden):
    """MojoGemmAllReduce resolves to the b200 class and, without a process group, is the plain projection."""
    os.environ["MOJO_BACKEND"] = "b200"
    import mojo_opset_b200 as m

    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 50, 192, generator=g).to(torch.bfloat16).to(DEV)
    w = torch.randn(192, 136, generator=g).to(torch.bfloat16).to(DEV)  # trans_weight layout [in, out]
    b = torch.randn(136, generator=g).to(torch.bfloat16).to(DEV)
    op = m.MojoGemmAllReduce(w, b, trans_weight=True)
    assert type(op).__name__ == "B200GemmAllReduce"
    out = op(x)
    ref = golden.gemm(x.float().cpu(), w.float().cpu(), b.float().cpu(), trans_weight=True)
    torch.testing.assert_close(out.cpu().float(), ref, **_tol(192))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (IPC + NVLink)")
def test_two_gpus_ipc():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "bench_gemm_allreduce.py"), "--check",
           "--steps", "3", "--warmup", "1", "--tokens", "256", "--out-features", "2048", "--in-features", "1024"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "PARITY OK" in res.stdout
