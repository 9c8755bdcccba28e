#!/usr/bin/env python
"""Developer tool: per-iteration timeline of CTA 0 of the tcgen05 attention kernel (needs a build with
MOJO_B200_EXTRA_NVCC_FLAGS=-DMOJO_ATTN_TRACE)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
os.environ["MOJO_B200_ATTN_IMPL"] = "tcgen05"
buf = torch.zeros(6 * 32 * 8, dtype=torch.int64, device="cuda")
os.environ["MOJO_B200_ATTN_TRACE_PTR"] = str(buf.data_ptr())
from mojo_opset_b200 import functional as F  # noqa: E402

D = 128
if "--prefill" in sys.argv:  # cfg3: block 0 is the longest query block (LPT order)
    Hq, Hkv, bs, T = 32, 8, 16, 8192
    nb = T // bs + 10
    kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device="cuda").normal_()
    vc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device="cuda").normal_()
    q = torch.empty(T, Hq, D, dtype=torch.bfloat16, device="cuda").normal_()
    table = torch.randperm(nb)[: T // bs].view(1, -1).to(torch.int32).cuda()
    cu = torch.tensor([0, T], dtype=torch.int32, device="cuda")
    run = lambda: F.paged_prefill_gqa(q, kc, vc, cu, table, None, cu, max_q_len=T, max_total_seq_len=T)
else:
    Bd, H, S = 2, 24, 4096
    qs, ks, vs = (torch.empty(Bd, S, H, D, dtype=torch.bfloat16, device="cuda").normal_().transpose(1, 2) for _ in range(3))
    run = lambda: F.sdpa(qs, ks, vs)
if "--zeros" in sys.argv:
    for t_ in ([q, kc, vc] if "--prefill" in sys.argv else [qs, ks, vs]):
        t_.zero_()
for _ in range(3):
    run()
torch.cuda.synchronize()
t = buf.cpu().view(6, 32, 8)
base = int(t[0, 4, 0])
names = ["sm0", "sm1", "mma", "sm0'", "sm1'", "mma'"]  # primed: block 1 (the peer CTA in pair mode)
for j in range(4, 14):
    for r in range(6):
        ev = [int(x) - base for x in t[r, j, :8]]
        print(f"j={j:2d} {names[r]}: " + " ".join(f"{e:7d}" for e in ev))
print("softmax events: 0 S ready | 1 S in regs | 2 max+rescale done | 3 exp done | 4 arrived")
print("mma events:     0 QK1(j+1) issued | 1 PV0(j) issued | 2 QK0(j+2) issued | 3 PV1(j) issued | 4 S free for QK0(j+1) | 5 S free for QK1(j+1) | 6 P0(j) seen | 7 P1(j) seen")
