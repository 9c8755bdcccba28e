"""-m gpu: the DiT block's non-GEMM ops (GELU, LayerNorm, grid RoPE; csrc/activation.cu, csrc/dit_ops.cu) against the
oracle on the same seeded inputs, through the op classes -> C ABI."""

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


@pytest.fixture(scope="module")
def golden():
    from oracle import golden as g  # the checker

    return g


def _tol(dtype):
    return dict(atol=1e-5, rtol=1e-5) if dtype == torch.float32 else dict(atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 300, 1536), (7,), (33, 1000), (3, 5, 13)])
def test_gelu(ops, golden, dtype, shape):
    g = torch.Generator().manual_seed(len(shape))
    x = (torch.randn(*shape, generator=g) * 3).to(dtype)
    op = ops.MojoGelu()
    assert type(op).__name__ == "B200Gelu"
    out = op(x.to(DEV))
    ref = golden.gelu(x)
    torch.testing.assert_close(out.cpu().float(), ref.float(), **_tol(dtype))


@pytest.mark.parametrize("hidden", [64, 128, 1536, 3072, 5120, 8192, 20480, 100, 734])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("affine", [True, False])
def test_layer_norm(ops, golden, hidden, dtype, affine):
    g = torch.Generator().manual_seed(hidden)
    rows = 300 if hidden <= 4096 else 17
    x = (torch.randn(2, rows, hidden, generator=g) * 2 + 0.5).to(dtype)
    op = ops.MojoLayerNorm(hidden, eps=1e-6, elementwise_affine=affine, device=DEV, dtype=dtype)
    assert type(op).__name__ == "B200LayerNorm"
    w = b = None
    if affine:
        w, b = torch.randn(hidden, generator=g).to(dtype), torch.randn(hidden, generator=g).to(dtype)
        with torch.no_grad():
            op.weight.copy_(w)
            op.bias.copy_(b)
    out = op(x.to(DEV))
    ref = golden.layer_norm(x, w, b, 1e-6)
    tol = dict(atol=2e-5, rtol=2e-5) if dtype == torch.float32 else dict(atol=3e-2, rtol=2e-2)
    torch.testing.assert_close(out.cpu().float(), ref.float(), **tol)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_grid_rope(ops, golden, dtype):
    g = torch.Generator().manual_seed(4)
    B, L, N, D = 2, 260, 12, 128
    grids = torch.tensor([[2, 10, 13], [1, 12, 20]])  # seq_len 260 (full) and 240 (20 padding tokens)
    x = torch.randn(B, L, N, D, generator=g).to(dtype)
    freqs = []
    for f, h, w in grids.tolist():
        ang = torch.rand(f * h * w, 1, D // 2, generator=g, dtype=torch.float64) * 6.28
        freqs.append(torch.polar(torch.ones_like(ang), ang))  # complex128, as the Wan model builds it
    op = ops.MojoGridRoPE()
    assert type(op).__name__ == "B200GridRoPE"
    out = op(x.to(DEV), grids, [fr.to(DEV) for fr in freqs])
    ref = golden.grid_rope(x, grids, freqs)
    torch.testing.assert_close(out.cpu().float(), ref.float(), **_tol(dtype))
    assert torch.equal(out[1, 240:].cpu(), x[1, 240:])  # padding tokens pass through bit-exactly
