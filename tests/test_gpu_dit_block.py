"""-m gpu: the Wan2.2-shaped DiT attention block (``examples/dit_block_synthetic.py``, mirror of the reference's
``modeling/wan2_2/mojo_wan_model.py:39-187``) on the b200 ops against the same block on the oracle's ops (CPU, same
weights and inputs): LayerNorm / RMSNorm / GridRoPE / Sdpa (self, S = 1024, and cross, Skv = 77) / GELU composed."""

import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dit_block_matches_oracle(torch_backend):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import dit_block_synthetic as ex

    import mojo_opset_b200 as ops

    dim, heads, ffn, grid, text = 512, 4, 1024, (2, 16, 32), 77  # head_dim 128, 1024 image tokens
    dtype = torch.bfloat16
    os.environ["MOJO_BACKEND"] = "b200"
    blk = ex.DiTBlock(ops, dim, ffn, heads, device="cuda", dtype=dtype, seed=3).eval()
    assert type(blk.self_attn.sdpa).__name__ == "B200Sdpa" and type(blk.norm1).__name__ == "B200LayerNorm"
    assert type(blk.self_attn.grid_rope).__name__ == "B200GridRoPE" and type(blk.ffn[1]).__name__ == "B200Gelu"
    os.environ["MOJO_BACKEND"] = "torch"
    try:
        ref = ex.DiTBlock(ops, dim, ffn, heads, device="cpu", dtype=dtype, seed=3).eval()
    finally:
        os.environ["MOJO_BACKEND"] = "b200"
    assert type(ref.self_attn.sdpa).__name__ == "TorchSdpa"
    ref.load_state_dict({k: v.cpu() for k, v in blk.state_dict().items()})
    x, e, grid_sizes, freqs, context = ex.make_inputs(2, grid, text, dim, dim // heads, "cpu", dtype)
    with torch.inference_mode():
        want = ref(x, e, grid_sizes, freqs, context)
        got = blk(x.cuda(), e.cuda(), grid_sizes.cuda(), [f.cuda() for f in freqs], context.cuda())
    assert got.shape == want.shape == (2, 1024, dim)
    # a whole block in bf16: activations of O(1..10) after two attention layers and an FFN
    torch.testing.assert_close(got.cpu().float(), want.float(), atol=1e-1, rtol=5e-2)
