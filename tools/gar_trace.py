#!/usr/bin/env python
"""Developer tool: timeline (SM clocks) of CTAs 0 and 1 of the GEMM half of MojoGemmAllReduce (world 1).  Needs a build
with -DMOJO_GAR_TRACE (tools/build_variant.sh tr gemm_allreduce.cu -DMOJO_GAR_TRACE; MOJO_B200_LIB=...)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
buf = torch.zeros(6 * 96, dtype=torch.int64, device="cuda")
os.environ["MOJO_B200_GAR_TRACE_PTR"] = str(buf.data_ptr())
from mojo_opset_b200 import functional as F  # noqa: E402
m, n, k = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (256, 8192, 4096)))
cold = "--cold" in sys.argv
xs = [torch.randn(m, k, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
ws = [(torch.randn(n, k, device="cuda") / k ** 0.5).to(torch.bfloat16) for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(4):
    if cold:
        flush.zero_()
    buf.zero_()
    F.gemm_allreduce(xs[i], ws[i], None, None)
torch.cuda.synchronize()
t = buf.cpu().view(6, 96)
names = ["prod0", "mma0", "epi0", "prod1", "relay1", "epi1"]
base = int(t[2, 0])
for r in range(6):
    ev = [int(x) - (base if r < 3 else int(t[5, 0])) for x in t[r] if int(x) != 0]
    print(names[r], " ".join(f"{e}" for e in ev))
print("prod: [0] before pdl_wait, [c+1] empty wait of k-block c returned | mma/relay: [c+1] full wait of k-block c returned | "
      "epi: [0] kernel entry, [1] accumulator ready, [2] tile stored")
