#!/usr/bin/env python
"""Generates tests/golden/swa.pt: inputs and outputs of the UNMODIFIED reference's non-paged ``MojoSWA``
(``mojo_opset/core/operators/attention.py:747-838``, torch backend, CPU) on seeded ragged batches.

    PYTHONPATH=/root/reference python tests/golden/make_swa_golden.py
"""
import os

import torch

os.environ["MOJO_BACKEND"] = "torch"
os.environ["MOJO_OPSET_PLUGIN_AUTOLOAD"] = "0"
import mojo_opset  # noqa: E402  (the reference)

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, q_lens, prefix, hq, hkv, d, dtype, layout, local, glob, seed):
    g = torch.Generator().manual_seed(seed)
    kv_lens = [a + b for a, b in zip(q_lens, prefix)]
    q = torch.randn(sum(q_lens), hq, d, generator=g).to(dtype)
    k = torch.randn(sum(kv_lens), hkv, d, generator=g).to(dtype)
    v = torch.randn(sum(kv_lens), hkv, d, generator=g).to(dtype)
    cu_q = torch.tensor([0] + torch.tensor(q_lens).cumsum(0).tolist(), dtype=torch.int32)
    cu_kv = torch.tensor([0] + torch.tensor(kv_lens).cumsum(0).tolist(), dtype=torch.int32)
    op = mojo_opset.MojoSWA(is_causal=True, gqa_layout=layout, global_window_size=glob, local_window_size=local)
    assert type(op).__name__ == "TorchSWA"
    out = op(q, k, v, cu_q, cu_kv, None)
    return dict(name=name, query=q, key=k, value=v, cu_q_lens=cu_q, cu_total_seq_lens=cu_kv, gqa_layout=layout,
                local_window_size=local, global_window_size=glob, out=out)


def main():
    bf, fh = torch.bfloat16, torch.float16
    cases = [
        case("ragged_local_bf16", [70, 33, 129], [0, 40, 7], 8, 2, 128, bf, "AABB", 31, None, 1),
        case("ragged_local_global_bf16", [200, 5], [60, 300], 4, 2, 128, bf, "AABB", 64, 16, 2),
        case("abab_global_only_fp16", [90, 150], [10, 0], 4, 2, 128, fh, "ABAB", None, 48, 3),
        case("plain_causal_d64_bf16", [64, 100], [0, 28], 4, 4, 64, bf, "AABB", None, None, 4),
        case("long_local_bf16", [700], [100], 2, 1, 128, bf, "AABB", 130, 5, 5),
    ]
    torch.save(cases, os.path.join(HERE, "swa.pt"))
    print("swa.pt:", [c["name"] for c in cases])


if __name__ == "__main__":
    main()
