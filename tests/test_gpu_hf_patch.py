"""-m gpu: ``apply_mojo_to_qwen3`` (reference ``mojo_opset/utils/patching.py:4-59``, driver ``examples/qwen3_patch.py``):
a HuggingFace Qwen3 built while patched runs its RoPE / RMSNorm / SwiGLU through the b200 kernels and produces the
logits of the unpatched model (same state dict) within the bf16 tolerance; HF ``generate`` works on it."""

import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pair():
    os.environ["MOJO_BACKEND"] = "b200"
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import qwen3_patch_synthetic as ex

    return ex, ex.build_pair(layers=3, hidden=512, heads=8, kv_heads=2, head_dim=128, inter=1536, vocab=2048)


def test_patched_modules_are_b200(pair):
    ex, (patched, plain) = pair
    counts = ex.mojo_module_counts(patched)
    # per layer: input/post norms + q/k norms, + the final norm; one rotary op shared by the module
    assert counts.get("B200RMSNorm") == 3 * 4 + 1
    assert counts.get("MojoSwiGLUMLP") == 3
    assert not ex.mojo_module_counts(plain)
    from transformers.models.qwen3 import modeling_qwen3

    assert modeling_qwen3.Qwen3RMSNorm.__name__ == "Qwen3RMSNorm"  # reverted


@pytest.mark.parametrize("batch,seq", [(1, 37), (3, 128)])
def test_logits_match_unpatched(pair, batch, seq):
    _, (patched, plain) = pair
    ids = torch.randint(0, 2048, (batch, seq), generator=torch.Generator().manual_seed(seq)).cuda()
    with torch.inference_mode():
        a, b = patched(ids).logits.float(), plain(ids).logits.float()
    # three bf16 layers of (norm, rope, swiglu) with different rounding points from HF's eager math
    torch.testing.assert_close(a, b, atol=6e-2, rtol=6e-2)


def test_generate_runs(pair):
    _, (patched, plain) = pair
    ids = torch.randint(0, 2048, (2, 16), generator=torch.Generator().manual_seed(0)).cuda()
    with torch.inference_mode():
        out = patched.generate(ids, max_new_tokens=8, do_sample=False, pad_token_id=0)
    assert out.shape == (2, 24)
    assert torch.equal(out[:, :16], ids)
