#!/usr/bin/env python
"""Developer tool: per-CTA timeline (globaltimer, ns) of the split-KV decode kernel at a small batch.  Needs a build with
-DMOJO_DECODE_TRACE (tools/build_variant.sh dtr paged_decode.cu -DMOJO_DECODE_TRACE; MOJO_B200_LIB=...)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
B, ctx = (int(v) for v in (sys.argv[1:3] if len(sys.argv) >= 3 else (1, 32768)))
Hq, Hkv, D, bs = 32, 8, 128, 16
buf = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
os.environ["MOJO_B200_DECODE_TRACE_PTR"] = str(buf.data_ptr())
from mojo_opset_b200 import functional as F  # noqa: E402
nb = B * ctx // bs + 10
kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device="cuda").normal_(); vc = torch.empty_like(kc).normal_()
q = torch.empty(B, Hq, D, dtype=torch.bfloat16, device="cuda").normal_()
table = torch.randperm(nb)[: B * ctx // bs].view(B, -1).to(torch.int32).cuda()
lens = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
fn = lambda: F.paged_decode_gqa(q, kc, vc, lens, table, max_total_seq_len=ctx)
for _ in range(3): fn()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(6): fn()
g.replay(); torch.cuda.synchronize()
buf.zero_(); g.replay(); torch.cuda.synchronize()       # the trace holds the LAST launch of the graph
t = buf.cpu().view(-1, 8)
t = t[t[:, 0] != 0]
t0 = int(t[:, 0].min())
import statistics as st
names = ["entry", "past pdl_wait", "first tile landed", "last tile consumed", "partial written"]
print(f"B={B} ctx={ctx}: {t.shape[0]} CTAs; ns relative to the first CTA's entry")
for i, n in enumerate(names):
    col = sorted(int(x) - t0 for x in t[:, i])
    print(f"  {n:20s} min {col[0]:6d}  p10 {col[len(col)//10]:6d}  median {col[len(col)//2]:6d}  p90 {col[9*len(col)//10]:6d}  max {col[-1]:6d}")
fold = sorted(int(r[5]) - t0 for r in t if int(r[5]) != 0 and int(r[5]) - int(r[4]) > 300)
if fold:
    print(f"  fold done (the groups' last CTAs, {len(fold)}): min {fold[0]} median {fold[len(fold)//2]} max {fold[-1]}")
dur = sorted(int(r[3] - r[2]) for r in t)
print(f"  streaming (first landed -> last consumed) per CTA: min {dur[0]} median {dur[len(dur)//2]} max {dur[-1]} ns")
