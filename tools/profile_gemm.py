#!/usr/bin/env python
"""ncu target: the GEMM half of MojoGemmAllReduce and cuBLAS at the cfg4 o_proj shapes (a few launches each)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MOJO_BACKEND"] = "b200"
from mojo_opset_b200 import functional as F
for m, n, k in ((256, 8192, 4096), (256, 8192, 1024)):
    xs = [torch.randn(m, k, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
    ws = [(torch.randn(n, k, device="cuda") / k ** 0.5).to(torch.bfloat16) for _ in range(4)]
    for i in range(3):
        F.gemm_allreduce(xs[i], ws[i], None, None)
    for i in range(3):
        torch.nn.functional.linear(xs[i], ws[i])
    torch.cuda.synchronize()
