#!/usr/bin/env python
"""Headline benchmark of the paged-attention decoder hot path (BASELINE.json: metric / configs[1]).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this repo's CUDA path (impl b200)
    torchrun ... bench.py --gpus N --steps K --warmup W             # N ranks, one per GPU, data parallel
    python bench.py --impl reference --steps 3 --warmup 1           # torch-native golden port on host cores

A "step" is one pass of the hot path over one batch of the Qwen3-8B-shaped decode workload (cfg2: batch 64,
32 q / 8 kv heads, head_dim 128, page 16, context 4096, bf16): ResidualAdd+RMSNorm -> RoPE -> StorePagedKVCache
-> PagedDecodeGQA -> SwiGLU, exactly one decoder layer's share of these ops.  Consecutive steps rotate over
`--layers` distinct KV caches (1.07 GB each) so every step streams cold KV (inputs larger than L2).

One JSON line is printed by rank 0:
  value     tokens/s (= batch * ranks / step time) with every input already resident in HBM;
  e2e       the same through the op modules with HOST (pinned) inputs: H2D of the step's inputs and D2H of its
            outputs inside the timed region, pipelined two deep over copy-in / compute / copy-out streams (the KV
            cache itself is device-resident state, as in serving);
  roofline  the dominant kernel (paged decode): algorithmic bytes / CUDA-event time vs the measured HBM peak;
  cpu_baseline  the oracle port of the same step on the host cores, bounded sample;
  extra     prefill (cfg3) TFLOP/s and DiT SDPA (cfg5, per-GPU slice) TFLOP/s vs the measured bf16 peak.
Multi-GPU: sequences shard across ranks with no data-path collective (weak scaling, max-over-ranks timing).
"""

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(batch=64, hq=32, hkv=8, d=128, bs=16, ctx=4096, hidden=4096, inter=12288, eps=1e-6, theta=1e6)
METRIC = "paged decode-GQA hot-path step throughput (Qwen3-8B-shaped: store+RMSNorm+RoPE+decode+SwiGLU)"
WORKLOAD = ("cfg2 Qwen3-8B-shaped paged decode layer: batch 64/GPU, 32q/8kv heads, hd 128, page 16, ctx 4096, bf16; "
            "ResidualAddRMSNorm 64x4096 + RoPE + StorePagedKVCache + PagedDecodeGQA + SwiGLU 64x12288")


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at communicator
    creation, for one): point fd 1 at stderr for the duration of the run and keep the real stdout for the line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------------
def make_decode_inputs(cfg, batch, layers, seed, device, dtype=torch.bfloat16):
    """Host-side (pinned) per-step inputs + device-resident KV caches of `layers` layers."""
    g = torch.Generator().manual_seed(seed)
    B, Hq, Hkv, D, bs, ctx = batch, cfg["hq"], cfg["hkv"], cfg["d"], cfg["bs"], cfg["ctx"]
    blocks_per_seq = ctx // bs
    nb = B * blocks_per_seq + 10

    def pin(t):
        return t.pin_memory() if device != "cpu" else t

    host = dict(
        hidden=pin(torch.randn(B, cfg["hidden"], generator=g).to(dtype)),
        residual=pin(torch.randn(B, cfg["hidden"], generator=g).to(dtype)),
        q=pin(torch.randn(B, Hq, D, generator=g).to(dtype)),
        k=pin(torch.randn(B, Hkv, D, generator=g).to(dtype)),
        v=pin(torch.randn(B, Hkv, D, generator=g).to(dtype)),
        gate=pin(torch.randn(B, cfg["inter"], generator=g).to(dtype)),
        up=pin(torch.randn(B, cfg["inter"], generator=g).to(dtype)),
        total_seq_lens=pin(torch.full((B,), ctx, dtype=torch.int32)),
        context_lens=pin(torch.full((B,), ctx - 1, dtype=torch.int32)),
    )
    inv_freq = 1.0 / (cfg["theta"] ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    ang = torch.full((B, 1), float(ctx - 1)) * inv_freq[None, :]
    emb = torch.cat((ang, ang), dim=-1)
    host["cos"], host["sin"] = pin(emb.cos()), pin(emb.sin())
    host["norm_weight"] = torch.randn(cfg["hidden"], generator=g).to(dtype)
    tables, metas, caches = [], [], []
    for _ in range(layers):
        perm = torch.randperm(nb, generator=g)[: B * blocks_per_seq].view(B, blocks_per_seq).to(torch.int32)
        tables.append(pin(perm.contiguous()))
        # store plan of the new token at position ctx-1: (src token, block, offset, len)
        meta = torch.stack((torch.arange(B, dtype=torch.int32), perm[:, (ctx - 1) // bs],
                            torch.full((B,), (ctx - 1) % bs, dtype=torch.int32), torch.ones(B, dtype=torch.int32)), -1)
        metas.append(pin(meta.contiguous()))
        if device != "cpu":
            kc = torch.empty(nb, Hkv, bs, D, dtype=dtype, device=device).normal_()
            vc = torch.empty(nb, Hkv, bs, D, dtype=dtype, device=device).normal_()
        else:
            kc = torch.randn(nb, Hkv, bs, D, generator=g).to(dtype)
            vc = torch.randn(nb, Hkv, bs, D, generator=g).to(dtype)
        caches.append((kc, vc))
    host["tables"], host["metas"] = tables, metas
    return host, caches


def decode_bytes(cfg, batch, dtype_bytes=2):
    """Algorithmic bytes of one paged-decode launch."""
    kv = 2 * batch * cfg["ctx"] * cfg["hkv"] * cfg["d"] * dtype_bytes
    qo = 2 * batch * cfg["hq"] * cfg["d"] * dtype_bytes
    table = batch * (cfg["ctx"] // cfg["bs"]) * 4 + batch * 4
    return kv + qo + table


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py measures the CUDA path: no GPU visible"
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    # N ranks on one box: keep every rank - and the pinned host buffers it is about to allocate (first touch) - on its
    # GPU's NUMA node, so the e2e leg's H2D / D2H copies do not cross the socket interconnect (MOJO_BENCH_NUMA=0: off)
    numa = "unbound"
    if world > 1 and os.environ.get("MOJO_BENCH_NUMA", "1") != "0":
        try:
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = f"rank bound to the {len(os.sched_getaffinity(0))} CPUs nearest GPU {local_rank}"
        except Exception as e:  # noqa: BLE001 - best effort: an unbound rank is still a valid measurement
            numa = f"unbound ({type(e).__name__})"

    os.environ["MOJO_BACKEND"] = "b200"
    import mojo_opset_b200 as m

    cfg, B = CFG2, CFG2["batch"]
    host, caches = make_decode_inputs(cfg, B, args.layers, 20260716 + 2 + rank, dev)
    norm = m.MojoResidualAddRMSNorm(cfg["hidden"], eps=cfg["eps"], device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        norm.weight.copy_(host["norm_weight"])
    rope, store, decode, swiglu = m.MojoApplyRoPE(), m.MojoStorePagedKVCache(), m.MojoPagedDecodeGQA(), m.MojoSwiGLU()
    assert type(decode).__name__ == "B200PagedDecodeGQA", "b200 backend not active"

    step_keys = ["hidden", "residual", "q", "k", "v", "gate", "up", "cos", "sin", "total_seq_lens"]
    res = {k: host[k].to(dev) for k in step_keys}
    res_tables = [t.to(dev) for t in host["tables"]]
    res_metas = [t.to(dev) for t in host["metas"]]
    decode_events = []

    def layer_step(d, table, meta, kc, vc, timed_decode=None):
        y, r = norm(d["hidden"], d["residual"])
        q_rot, k_rot = rope(d["q"], d["k"], d["cos"], d["sin"], head_first=False)
        store(k_rot, d["v"], kc, vc, chunk_metadata=meta)
        if timed_decode is not None:
            timed_decode[0].record()
        o = decode(q_rot, kc, vc, d["total_seq_lens"], table, max_total_seq_len=cfg["ctx"])
        if timed_decode is not None:
            timed_decode[1].record()
        a = swiglu(d["gate"], d["up"])
        return y, r, o, a

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i, False)
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(steps):
            fn(warmup + i, True)
        stop.record()
        barrier()
        ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # ---- value: inputs resident in HBM
    def resident_step(i, is_timed):
        L = i % args.layers
        ev = None
        if is_timed:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            decode_events.append(ev)
        layer_step(res, res_tables[L], res_metas[L], caches[L][0], caches[L][1], ev)

    # every number is the MEDIAN of `--repeats` timed regions of exactly K steps each (a 4 ms region on a fresh box is
    # at the mercy of one host hiccup; all samples are reported)
    def median(xs):
        xs = sorted(xs)
        return xs[len(xs) // 2]

    # The step is launched the way a serving loop launches it: one captured device graph per rotating KV cache
    # (mojo_opset_b200.runtime.DeviceGraphRunner, mirror of the reference's compile/device_graph.py).  `value` is the
    # graph-replayed step; the same K steps are ALSO timed launched one by one with CUDA events around every decode
    # launch - that pass feeds `roofline` and is reported as `eager`.  --eager makes it the `value`.
    res_graphs = []
    if not args.eager:
        for L in range(args.layers):
            layer_step(res, res_tables[L], res_metas[L], caches[L][0], caches[L][1])  # warm-up outside the capture
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                layer_step(res, res_tables[L], res_metas[L], caches[L][0], caches[L][1])
            res_graphs.append(graph)
        torch.cuda.synchronize()

    def graph_step(i, is_timed):
        res_graphs[i % args.layers].replay()

    with ClockSampler(local_rank) as clocks:
        eager_samples = []
        for _ in range(args.repeats):
            eager_samples.append(timed(resident_step, args.steps, args.warmup))
        ms_eager = median(eager_samples)
        resident_samples = eager_samples
        if not args.eager:
            resident_samples = [timed(graph_step, args.steps, args.warmup) for _ in range(args.repeats)]
        ms_resident = median(resident_samples)
        # keep the GPU busy long enough for the sampler to see clocks under load on short runs
        extra_rounds = 0
        while len(clocks.rows) < 3 and extra_rounds < 20:
            for i in range(args.steps):
                resident_step(i, False)
            torch.cuda.synchronize()
            extra_rounds += 1
    torch.cuda.synchronize()
    decode_ms = median([a.elapsed_time(b) for a, b in decode_events])  # per-launch CUDA-event time, all repeats

    # ---- e2e: host buffers in, host results out, every step.  A serving loop overlaps PCIe with compute, so the step
    # is a 2-deep pipeline over three streams: step i+1's inputs travel host->device (ONE packed pinned buffer: all
    # per-step tensors, the block table and the store plan) while step i computes and step i-1's results travel back.
    # Every step's copies are inside the timed region; the host "consumes" a result buffer before it is reused.
    names = step_keys + ["table", "meta"]
    specs, offset = {}, 0
    for k in names:
        t = host[k] if k in host else (host["tables"][0] if k == "table" else host["metas"][0])
        specs[k] = (offset, t.shape, t.dtype, t.numel() * t.element_size())
        offset += (specs[k][3] + 255) // 256 * 256
    h2d_bytes = sum(v[3] for v in specs.values())

    def views(buf):
        return {k: buf[o:o + n].view(dt).view(shape) for k, (o, shape, dt, n) in specs.items()}

    host_packed = [torch.empty(offset, dtype=torch.uint8).pin_memory() for _ in range(args.layers)]
    for L in range(args.layers):  # one packed host buffer per rotating layer (they differ in table / plan)
        hv = views(host_packed[L])
        for k in step_keys:
            hv[k].copy_(host[k])
        hv["table"].copy_(host["tables"][L])
        hv["meta"].copy_(host["metas"][L])
    dev_packed = [torch.empty(offset, dtype=torch.uint8, device=dev) for _ in range(2)]
    dev_views = [views(b) for b in dev_packed]
    out_host = [dict(
        y=torch.empty(B, cfg["hidden"], dtype=torch.bfloat16).pin_memory(),
        o=torch.empty(B, cfg["hq"], cfg["d"], dtype=torch.bfloat16).pin_memory(),
        a=torch.empty(B, cfg["inter"], dtype=torch.bfloat16).pin_memory()) for _ in range(2)]
    d2h_bytes = sum(t.numel() * t.element_size() for t in out_host[0].values())
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_ready = [torch.cuda.Event() for _ in range(2)]
    compute_done = [torch.cuda.Event() for _ in range(2)]
    out_done = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(i, is_timed):
        j, L = i % 2, i % args.layers
        main = torch.cuda.current_stream()
        with torch.cuda.stream(s_in):
            s_in.wait_event(compute_done[j])           # buffer j was last read by step i-2's kernels
            dev_packed[j].copy_(host_packed[L], non_blocking=True)
            in_ready[j].record(s_in)
        main.wait_event(in_ready[j])
        if e2e_graphs is not None:  # one graph launch instead of five kernel launches from Python
            graph, (y, _, o, a) = e2e_graphs[(j, L)]
            graph.replay()
        else:
            d = dev_views[j]
            y, _, o, a = layer_step(d, d["table"], d["meta"], caches[L][0], caches[L][1])
        compute_done[j].record(main)
        out_done[j].synchronize()                       # the caller has read step i-2's results from out_host[j]
        with torch.cuda.stream(s_out):
            s_out.wait_event(compute_done[j])
            for k, t in (("y", y), ("o", o), ("a", a)):
                out_host[j][k].copy_(t, non_blocking=True)
                t.record_stream(s_out)
            out_done[j].record(s_out)

    # The user-facing way to run a decode step is a captured device graph (mojo_opset_b200.runtime.DeviceGraphRunner,
    # mirror of the reference's compile/device_graph.py): the five kernels of the layer step are captured once per
    # (device input buffer, KV cache) pair and replayed.  --e2e-eager launches them one by one from Python instead
    # (then the host, ~35 us per launch through ctypes, is as slow as the GPU step and the number jitters).
    e2e_graphs = None
    if not args.e2e_eager:
        e2e_graphs = {}
        for j in range(2):
            for L in range(args.layers):
                d = dev_views[j]
                dev_packed[j].copy_(host_packed[L])
                layer_step(d, d["table"], d["meta"], caches[L][0], caches[L][1])  # warm-up outside the capture
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    outs = layer_step(d, d["table"], d["meta"], caches[L][0], caches[L][1])
                e2e_graphs[(j, L)] = (graph, outs)
        torch.cuda.synchronize()

    def e2e_timed(steps, warmup):
        for ev in compute_done + out_done:
            ev.record()
        for i in range(warmup):
            e2e_step(i, False)
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        s_in.wait_event(start)
        s_out.wait_event(start)
        for i in range(steps):
            e2e_step(warmup + i, True)
        main = torch.cuda.current_stream()
        main.wait_event(out_done[0])
        main.wait_event(out_done[1])
        stop.record()
        barrier()
        ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    e2e_samples = [e2e_timed(args.steps, args.warmup) for _ in range(args.repeats)]
    ms_e2e = median(e2e_samples)

    peaks = load_peaks()
    algo_bytes = decode_bytes(cfg, B)
    achieved = algo_bytes / (decode_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "decode_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    splits = m._lib.load().mojo_b200_paged_decode_num_splits(B, cfg["hq"], cfg["hkv"], cfg["d"], cfg["bs"], cfg["ctx"], 0)
    launches_per_step = 1 + 1 + 1 + (2 if splits > 1 else 1) + 1
    line = {
        "metric": METRIC,
        "value": B * world * args.steps / (ms_resident * 1e-3),
        "unit": "tokens/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_resident / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "impl": "b200",
        "config": {
            "workload": WORKLOAD,
            "batch_per_gpu": B, "ctx": cfg["ctx"], "block_size": cfg["bs"], "parallelism": f"dp{world}",
            "l2": f"inputs larger than L2: {args.layers} rotating KV caches of {2 * B * cfg['ctx'] * cfg['hkv'] * cfg['d'] * 2 / 1e9:.2f} GB",
            "decode_splits": splits,
            "host_affinity": numa,
            "launch": "eager, one kernel launch at a time" if args.eager else
                      f"CUDA graph replay, one graph of 5 kernels per rotating KV cache ({args.layers} graphs)",
        },
        "eager": {"ms_per_step": ms_eager / args.steps, "value": B * world * args.steps / (ms_eager * 1e-3),
                  "ms_samples": [round(x, 4) for x in eager_samples],
                  "note": "the same K steps launched one by one from Python with CUDA events around every decode "
                          "launch: the pass `roofline` is measured in"},
        "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "tokens/s",
                "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "pipeline": "2-deep: packed pinned H2D | compute | D2H on three streams",
                "compute_launch": "eager (5 launches per step)" if args.e2e_eager else "CUDA graph replay per step"},
        "gpu_launches": launches_per_step * args.steps,
        "timing": {"repeats": args.repeats, "stat": "median over repeats of a timed region of exactly `steps` steps",
                   "resident_ms_samples": [round(x, 4) for x in resident_samples],
                   "e2e_ms_samples": [round(x, 4) for x in e2e_samples]},
        "roofline": {"kernel": "paged_decode_mma_kernel (+ reduce)", "bound": "hbm", "achieved": achieved,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "frac_of_8TBs_nominal": achieved / 8000.0, "peak_source": peaks["source"],
                     "algorithmic_bytes": algo_bytes, "us_per_launch": decode_ms * 1e3, "traffic": traffic},
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_extra:
        line["extra"] = run_extra(m, dev, peaks)
        # the metric names two rooflines: HBM for decode (`roofline`) and the bf16 tensor peak for prefill
        pf = line["extra"]["prefill_cfg3"]
        line["roofline_prefill"] = {
            "kernel": "attn_fwd_sm100_kernel (MojoPagedPrefillGQA, cfg3: T=8192 causal, 32q/8kv, hd 128, page 16)",
            "bound": "tensor", "achieved": pf["tflops"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": pf["frac_of_bf16_peak"], "frac_of_2250_nominal": pf["tflops"] / 2250.0,
            "peak_source": peaks["source"], "algorithmic_flops": pf["flops"], "us_per_launch": pf["ms"] * 1e3,
            "traffic": None}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_sample_batch)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _time_gpu(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def run_extra(m, dev, peaks):
    """cfg3 (prefill 8192, causal) and cfg5 (DiT SDPA, 2 of 16 batch elements = one GPU's share at 8 GPUs)."""
    out = {}
    g = torch.Generator().manual_seed(20260716 + 3)
    Hq, Hkv, D, bs, T = 32, 8, 128, 16, 8192
    nb = T // bs + 10
    kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=dev).normal_()
    vc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=dev).normal_()
    q = torch.empty(T, Hq, D, dtype=torch.bfloat16, device=dev).normal_()
    table = torch.randperm(nb, generator=g)[: T // bs].view(1, -1).to(torch.int32).to(dev)
    cu = torch.tensor([0, T], dtype=torch.int32, device=dev)
    prefill = m.MojoPagedPrefillGQA()
    ms = _time_gpu(lambda: prefill(q, kc, vc, cu, table, max_q_len=T, max_total_seq_len=T), 10)
    flops = 4 * Hq * D * (T * (T + 1) // 2)
    out["prefill_cfg3"] = {"workload": "MojoPagedPrefillGQA T=8192 causal 32q/8kv hd128 page16 bf16", "ms": ms,
                           "tflops": flops / ms / 1e9, "frac_of_bf16_peak": flops / ms / 1e9 / peaks["bf16_tflops"],
                           "tokens_per_s": T / (ms * 1e-3), "flops": flops}
    Bd, H, S = 2, 24, 4096
    qs, ks, vs = (torch.empty(Bd, S, H, D, dtype=torch.bfloat16, device=dev).normal_().transpose(1, 2) for _ in range(3))
    sdpa = m.MojoSdpa()
    ms = _time_gpu(lambda: sdpa(qs, ks, vs), 10)
    flops = 4 * Bd * H * S * S * D
    out["sdpa_cfg5_per_gpu"] = {"workload": "MojoSdpa DiT 24 heads hd128 S=4096 non-causal bf16, 2 of 16 batch "
                                            "elements (one GPU's share at 8 GPUs), transposed-BSHD views",
                                "ms": ms, "tflops": flops / ms / 1e9,
                                "frac_of_bf16_peak": flops / ms / 1e9 / peaks["bf16_tflops"], "flops": flops}
    return out


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------
def oracle_step(golden, host, caches, layer, cfg):
    kc, vc = caches[layer]
    y, r = golden.residual_add_rms_norm(host["hidden"], host["residual"], host["norm_weight"], cfg["eps"])
    q_rot, k_rot = golden.apply_rope(host["q"], host["k"], host["cos"], host["sin"], head_first=False)
    golden.store_paged_kv(k_rot, host["v"], kc, vc, host["metas"][layer])
    o = golden.paged_decode_gqa(q_rot, kc, vc, host["total_seq_lens"], host["tables"][layer])
    a = golden.swiglu(host["gate"], host["up"])
    return y, r, o, a


def cpu_baseline(sample_batch, steps=2, warmup=1):
    from oracle import golden

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    host, caches = make_decode_inputs(CFG2, sample_batch, 1, 20260716 + 2, "cpu")
    for _ in range(warmup):
        oracle_step(golden, host, caches, 0, CFG2)
    best = math.inf
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle_step(golden, host, caches, 0, CFG2)
        best = min(best, time.perf_counter() - t0)
    return {"value": sample_batch / best, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_batch} of 64 sequences of cfg2 (same per-sequence shapes, one layer step), "
                      f"best of {steps} after {warmup} warm-up; oracle/golden.py on CPU",
            "seconds_per_step": best}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import golden

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.cpu_sample_batch
    host, caches = make_decode_inputs(CFG2, B, 1, 20260716 + 2, "cpu")
    for _ in range(args.warmup):
        oracle_step(golden, host, caches, 0, CFG2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(golden, host, caches, 0, CFG2)
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    sample = (f"{B} of 64 sequences of cfg2 per step (same per-sequence shapes); the reference's torch-native "
              f"algorithm restated in oracle/golden.py, CPU, {torch.get_num_threads()} threads")
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "reference",
        # the b200 arm's workload; each CPU step is a bounded sample of it (cpu_baseline.sample), tokens/s is per token
        "config": {"workload": WORKLOAD, "batch_per_gpu": CFG2["batch"], "ctx": CFG2["ctx"], "block_size": CFG2["bs"],
                   "parallelism": "host cores (rank 0 only)", "sample_batch_per_step": B},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=4, help="distinct KV caches rotated across steps")
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of K steps each; the median is reported")
    ap.add_argument("--cpu-sample-batch", type=int, default=8)
    ap.add_argument("--e2e-eager", action="store_true", help="e2e leg: launch the step's kernels one by one")
    ap.add_argument("--eager", action="store_true", help="value = the eagerly launched step instead of graph replay")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        args.steps = 5 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 20 if args.steps is None else args.steps
    args.warmup = 5 if args.warmup is None else max(args.warmup, 3)
    run_b200(args)


if __name__ == "__main__":
    main()
