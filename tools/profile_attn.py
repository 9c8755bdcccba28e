#!/usr/bin/env python
"""Run the tcgen05 attention kernel at cfg3 (paged prefill T=8192) and cfg5 (SDPA, one GPU's share) a few times, for
ncu captures (not a benchmark):
    ncu --set full --clock-control none --import-source on -k regex:attn_fwd_sm100 -c 4 -o gpurun_out/attn python tools/profile_attn.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
from mojo_opset_b200 import functional as F  # noqa: E402

DEV, D = "cuda", 128
Hq, Hkv, bs, T = 32, 8, 16, 8192
nb = T // bs + 10
kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
vc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=DEV).normal_()
q = torch.empty(T, Hq, D, dtype=torch.bfloat16, device=DEV).normal_()
table = torch.randperm(nb)[: T // bs].view(1, -1).to(torch.int32).to(DEV)
cu = torch.tensor([0, T], dtype=torch.int32, device=DEV)
Bd, H, S = 2, 24, 4096
qs, ks, vs = (torch.empty(Bd, S, H, D, dtype=torch.bfloat16, device=DEV).normal_().transpose(1, 2) for _ in range(3))
for _ in range(2):
    F.paged_prefill_gqa(q, kc, vc, cu, table, None, cu, max_q_len=T, max_total_seq_len=T)
    F.sdpa(qs, ks, vs)
torch.cuda.synchronize()
