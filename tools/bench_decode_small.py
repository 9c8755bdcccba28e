#!/usr/bin/env python
"""Developer bench: small-batch split-KV decode (the batch-1 serving case) replayed in a CUDA graph."""
import os, sys, torch
sys.path.insert(0, "/root/repo"); os.environ["MOJO_BACKEND"]="b200"
from mojo_opset_b200 import functional as F
def bench(B, ctx, Hq=32, Hkv=8, D=128, bs=16):
    nb = B*ctx//bs + 10
    kc = torch.empty(nb,Hkv,bs,D,dtype=torch.bfloat16,device="cuda").normal_(); vc = torch.empty_like(kc).normal_()
    q = torch.empty(B,Hq,D,dtype=torch.bfloat16,device="cuda").normal_()
    table = torch.randperm(nb)[:B*ctx//bs].view(B,-1).to(torch.int32).cuda()
    lens = torch.full((B,),ctx,dtype=torch.int32,device="cuda")
    fn = lambda: F.paged_decode_gqa(q,kc,vc,lens,table,max_total_seq_len=ctx)
    for _ in range(5): fn()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(20): out = fn()
    g.replay(); torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b)/20*1e3
    byt = 2*B*ctx*Hkv*D*2
    return us, byt/us/1e3
for B, ctx in ((1, 32768), (4, 8192), (1, 8192), (8, 4096)):
    us, gbs = bench(B, ctx)
    print(f"B={B} ctx={ctx}: {us:.1f} us {gbs:.0f} GB/s of KV")
