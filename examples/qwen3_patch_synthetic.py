#!/usr/bin/env python
"""The reference's ``examples/qwen3_patch.py`` on the b200 backend, without a checkpoint.

``apply_mojo_to_qwen3()`` (``mojo_opset_b200/utils/patching.py``, mirror of reference ``mojo_opset/utils/patching.py``)
swaps HuggingFace Qwen3's rotary function, RMSNorm class and MLP class for the Mojo ops before the model is built;
then plain HF ``generate`` runs.  The GPU box has no network and no weights, so the model is
built from a config with random bf16 weights; an unpatched twin with the same state dict is the comparison.

    python examples/qwen3_patch_synthetic.py --layers 4 --prompt-len 64 --max-new-tokens 32
"""

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MOJO_BACKEND", "b200")


def build_pair(layers=4, hidden=4096, heads=32, kv_heads=8, head_dim=128, inter=12288, vocab=151936, seed=0,
               device="cuda", dtype=torch.bfloat16):
    """(patched model, unpatched model) sharing one random state dict."""
    from transformers import Qwen3Config
    from transformers import Qwen3ForCausalLM

    from mojo_opset_b200.utils.patching import apply_mojo_to_qwen3
    from mojo_opset_b200.utils.patching import revert_mojo_from_qwen3

    cfg = Qwen3Config(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                      num_key_value_heads=kv_heads, head_dim=head_dim, intermediate_size=inter, vocab_size=vocab,
                      rms_norm_eps=1e-6, rope_theta=1e6, max_position_embeddings=8192, tie_word_embeddings=False)
    torch.manual_seed(seed)
    plain = Qwen3ForCausalLM(cfg).to(dtype).to(device).eval()
    with torch.no_grad():  # non-trivial norm weights so that the norm kernels' scaling is exercised
        for name, p in plain.named_parameters():
            if name.endswith("norm.weight"):
                p.copy_(1 + 0.1 * torch.randn_like(p, dtype=torch.float32).to(dtype))
    apply_mojo_to_qwen3()
    try:
        patched = Qwen3ForCausalLM(cfg).to(dtype).to(device).eval()
    finally:
        revert_mojo_from_qwen3()
    missing, unexpected = patched.load_state_dict(plain.state_dict(), strict=True)
    assert not missing and not unexpected
    return patched, plain


def mojo_module_counts(model):
    counts = {}
    for m in model.modules():
        n = type(m).__name__
        if n.startswith(("B200", "Mojo")):
            counts[n] = counts.get(n, 0) + 1
    return counts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prompt-len", type=int, default=64)
    ap.add_argument("--max-new-tokens", type=int, default=32)
    args = ap.parse_args()

    patched, plain = build_pair(layers=args.layers)
    ids = torch.randint(0, 151936, (args.batch, args.prompt_len), device="cuda",
                        generator=torch.Generator(device="cuda").manual_seed(1))
    out = {"modules": mojo_module_counts(patched)}
    with torch.inference_mode():
        lp, lr = patched(ids).logits.float(), plain(ids).logits.float()
        out["logits_max_abs_diff"] = (lp - lr).abs().max().item()
        out["logits_ref_absmax"] = lr.abs().max().item()
        for name, model in (("patched", patched), ("plain", plain)):
            kw = dict(max_new_tokens=args.max_new_tokens, do_sample=False, pad_token_id=0)
            model.generate(ids, **kw)  # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gen = model.generate(ids, **kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out[f"{name}_tokens_per_s"] = args.batch * (gen.shape[1] - ids.shape[1]) / dt
    print(json.dumps(out))


if __name__ == "__main__":
    main()
