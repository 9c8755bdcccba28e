#!/usr/bin/env python
"""Developer bench (one GPU): the GEMM half of MojoGemmAllReduce (world 1) against cuBLAS at the cfg4 o_proj shapes,
CTA pairs on / off, graph replay over rotating inputs."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MOJO_BACKEND"] = "b200"
from mojo_opset_b200 import functional as F

def timed(fn, steps=20, warmup=3):
    for i in range(warmup): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(steps): fn(i)
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / steps * 1e3)
    return best

rows = []
for m, n, k in ((256, 8192, 1024), (256, 8192, 4096), (512, 8192, 4096), (8192, 8192, 1024), (64, 8192, 1024)):
    xs = [torch.randn(m, k, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
    ws = [(torch.randn(n, k, device="cuda") / k ** 0.5).to(torch.bfloat16) for _ in range(4)]   # rotating weights: a layer stack
    ref = timed(lambda i: torch.nn.functional.linear(xs[i % 4], ws[i % 4]))
    r = {"m": m, "n": n, "k": k, "cublas_us": ref}
    for pair in ("1", "0"):
        os.environ["MOJO_B200_GAR_PAIR"] = pair
        r["own_pair%s_us" % pair] = timed(lambda i: F.gemm_allreduce(xs[i % 4], ws[i % 4], None, None))
    out = F.gemm_allreduce(xs[0], ws[0], None, None)
    err = (out.float() - torch.nn.functional.linear(xs[0], ws[0]).float()).abs().max().item()
    r["max_abs_diff_vs_cublas"] = err
    rows.append(r); print(json.dumps(r), flush=True)
