#!/usr/bin/env python
"""Developer diagnostic: is the tcgen05 attention kernel cycle-bound or power-bound?  Runs cfg5 SDPA (and a cuBLAS
bf16 GEMM for comparison) back to back for a few seconds while sampling SM clock and board power through NVML.
Not part of the product path."""
import os, sys, threading, time, statistics, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOJO_BACKEND"] = "b200"
from mojo_opset_b200 import functional as F  # noqa: E402
import pynvml

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def sampled(fn, flops, seconds=3.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    clocks, power, stop = [], [], False

    def sampler():
        while not stop:
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.01)
    th = threading.Thread(target=sampler); th.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    a.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    stop = True; th.join()
    ms = a.elapsed_time(b) / n
    k = len(clocks) // 3
    return dict(ms=ms, tflops=flops / ms / 1e9, sm_mhz_median=statistics.median(clocks[k:]), sm_mhz_min=min(clocks[k:]),
                power_w_median=statistics.median(power[k:]), power_w_max=max(power))


D = 128
out = {}
H, S, Bd = 24, 4096, 16
qs, ks, vs = (torch.empty(Bd, S, H, D, dtype=torch.bfloat16, device="cuda").normal_().transpose(1, 2) for _ in range(3))
for emu in ("0", "1", "2"):
    os.environ["MOJO_B200_ATTN_EMU"] = emu
    out[f"sdpa_b16_emu{emu}"] = sampled(lambda: F.sdpa(qs, ks, vs), 4 * Bd * H * S * S * D)
os.environ.pop("MOJO_B200_ATTN_EMU")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
out["cublas_8192"] = sampled(lambda: torch.matmul(a, b), 2 * 8192 ** 3)
z = torch.zeros(8192, 8192, device="cuda", dtype=torch.bfloat16)
out["cublas_8192_zeros"] = sampled(lambda: torch.matmul(z, z), 2 * 8192 ** 3)
qz, kz, vz = (torch.zeros_like(t) for t in (qs, ks, vs))
out["sdpa_b16_zeros"] = sampled(lambda: F.sdpa(qz, kz, vz), 4 * Bd * H * S * S * D)
for k, v in out.items():
    print(k, json.dumps(v))
