#!/usr/bin/env python
"""Run each HBM-bound op once or twice at its prefill-sized shape (for ncu captures; not a benchmark)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
from mojo_opset_b200 import functional as F  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
T, H, I, Hq, Hkv, D, bs = 8192, 4096, 12288, 32, 8, 128, 16
rnd = lambda *s, dtype=BF: torch.empty(*s, dtype=dtype, device=DEV).normal_()  # noqa: E731
x, r, w = rnd(T, H), rnd(T, H), rnd(H)
q, k = rnd(T, Hq, D), rnd(T, Hkv, D)
cos, sin = rnd(T, D, dtype=torch.float32), rnd(T, D, dtype=torch.float32)
g, u = rnd(T, I), rnd(T, I)
nb = T // bs + 10
kc, vc = rnd(nb, Hkv, bs, D), rnd(nb, Hkv, bs, D)
table = torch.randperm(nb)[: T // bs].view(1, -1).to(torch.int32).to(DEV)
cu = torch.tensor([0, T], dtype=torch.int32, device=DEV)
ctx = torch.zeros(1, dtype=torch.int32, device=DEV)
wd = rnd(D)
lw, lb = rnd(H), rnd(H)
gx = rnd(256, 4096)
gw = rnd(8192, 4096)
for _ in range(2):
    F.norm_rope_store_kv(q, k, k, cos, sin, kc, vc, table, cu, ctx, wd, wd, 1e-6)
    F.layer_norm(x, lw, lb, 1e-6)
    F.gelu(g)
    F.gemm_allreduce(gx, gw, None, None)
    F.residual_add_rms_norm(x, r, w, 1e-6)
    F.rms_norm(x, w, 1e-6)
    F.rms_norm(q, wd, 1e-6)
    F.apply_rope(q, k, cos, sin, head_first=False)
    F.swiglu(g, u)
    F.store_paged_kv(k, k, kc, vc, block_table=table, cu_q_lens=cu, context_kv_lens=ctx)
torch.cuda.synchronize()
