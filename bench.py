#!/usr/bin/env python
"""Headline benchmark of the paged-attention decoder hot path (BASELINE.json: metric / configs[1]).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this repo's CUDA path (impl b200)
    torchrun ... bench.py --gpus N --steps K --warmup W             # N ranks, one per GPU
    python bench.py --impl reference --steps 20 --warmup 5          # the UNMODIFIED reference (baseline/_ref) on host cores

A "step" is one pass of the hot path over one batch of the Qwen3-8B-shaped decode workload (cfg2: batch 64,
32 q / 8 kv heads, head_dim 128, page 16, context 4096, bf16): ResidualAdd+RMSNorm -> RoPE -> StorePagedKVCache
-> PagedDecodeGQA -> SwiGLU, exactly one decoder layer's share of these ops.  Consecutive steps rotate over
`--layers` distinct KV caches (1.07 GB each) so every step streams cold KV (inputs larger than L2).

One JSON line is printed by rank 0:
  value      tokens/s (= batch * ranks / step time) with every input already resident in HBM (graph-replayed step);
  sustained  the same step replayed for >= 1 s (clocks under load, power cap);
  e2e        the same through the op modules with HOST (pinned) inputs: every step is ONE device graph holding the
             H2D copy of the step's packed inputs, the five kernels and the D2H copies of its results, launched
             round-robin on three streams so step i+1's copy-in and step i-1's copy-out overlap step i's kernels
             (the KV cache itself is device-resident state, as in serving);
  roofline   the dominant kernel (paged decode): algorithmic bytes / CUDA-event time vs the measured HBM peak;
  cpu_baseline  the reference's own torch-native ops (baseline/_ref; else the oracle port) on the host cores;
  extra      N=1: prefill (cfg3) TFLOP/s burst + sustained, DiT SDPA (cfg5) TFLOP/s vs the measured bf16 peak.
Multi-GPU (N > 1): `value` = cfg2 sequences sharded across ranks, no data-path collective (weak scaling, max-over-
ranks timing).  Two more legs put the rest of SURVEY 8(e) in the same line:
  tp_cfg4    Llama-3-70B-shaped decode layer step (batch 256, 64q/8kv, ctx 32k) with the KV heads sharded TP = N:
             StorePagedKVCache -> PagedDecodeGQA (local heads) -> MojoGemmAllReduce (o_proj GEMM + the path's ONLY
             collective in one kernel over NVLink peer memory), graph-replayed; cuBLAS + NCCL all-reduce beside it;
             the fused output is checked in-run against the oracle (non-zero exit on mismatch);
  dp_cfg5    DiT MojoSdpa (24 heads, 4096 tokens, batch 16) sharded by batch: 16/N elements per rank.
"""

import argparse
import hashlib
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
CFG4 = dict(batch=256, hq=64, hkv=8, d=128, bs=16, ctx=32768, hidden=8192)
CFG5 = dict(batch=16, heads=24, seq=4096, d=128)
METRIC = "paged decode-GQA hot-path step throughput (Qwen3-8B-shaped: store+RMSNorm+RoPE+decode+SwiGLU)"
WORKLOAD = ("cfg2 Qwen3-8B-shaped paged decode layer: batch 64/GPU, 32q/8kv heads, hd 128, page 16, ctx 4096, bf16; "
            "ResidualAddRMSNorm 64x4096 + RoPE + StorePagedKVCache + PagedDecodeGQA + SwiGLU 64x12288")
DECODE_SOURCES = ("mojo_opset_b200/csrc/paged_decode.cu", "mojo_opset_b200/csrc/tma.cuh",
                  "mojo_opset_b200/csrc/common.cuh")


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
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained"), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


def sources_sha256(paths=DECODE_SOURCES):
    h = hashlib.sha256()
    for rel in paths:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def load_decode_traffic():
    """DRAM bytes per decode launch from the ncu `--set full` capture of THIS decode kernel: the capture records the
    hash of the kernel's sources (tools/decode_traffic.py writes it); a capture of other sources is refused."""
    path = os.path.join(ROOT, "profiles", "decode_traffic.json")
    if not os.path.exists(path):
        return None, "no capture"
    with open(path) as f:
        t = json.load(f)
    if t.get("source_sha256") == sources_sha256():
        return t.get("dram_bytes_per_launch"), t.get("source")
    # the sources changed: the capture still stands if the MACHINE CODE of the measured kernel did not (a change
    # elsewhere in the file, e.g. in the split-KV instantiation); mojo_opset_b200/build.py pins it at build time
    built = os.path.join(ROOT, "mojo_opset_b200", "kernel_sass.json")
    if t.get("kernel_sass_sha256") and os.path.exists(built):
        with open(built) as f:
            now = [v.get("sass_sha256") for v in json.load(f).values()]
        if t["kernel_sass_sha256"] in now:
            return t.get("dram_bytes_per_launch"), t.get("source") + "; sources changed since, the measured kernel's SASS is identical"
    return None, "stale capture refused (profiles/decode_traffic.json was taken on another decode kernel)"


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


def decode_bytes(batch, ctx, hq, hkv, d, bs, dtype_bytes=2):
    """Algorithmic bytes of one paged-decode launch."""
    kv = 2 * batch * ctx * hkv * d * dtype_bytes
    qo = 2 * batch * hq * d * dtype_bytes
    table = batch * (ctx // bs) * 4 + batch * 4
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


def bind_to_gpu_numa_node(local_rank):
    """Keep the rank - and the pinned host buffers it is about to allocate (first touch) - on the NUMA node its GPU's
    PCIe slot hangs off, read from sysfs (NVML's affinity reports node 0 for every GPU on this pool's hosts)."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return f"unbound (sysfs reports no NUMA node for {bdf})"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"unbound (no allowed CPU on node {node})"
        os.sched_setaffinity(0, cpus)
        return f"GPU {local_rank} ({bdf}) on NUMA node {node}: rank bound to its {len(cpus)} CPUs"
    except Exception as e:  # noqa: BLE001 - best effort: an unbound rank is still a valid measurement
        return f"unbound ({type(e).__name__}: {e})"


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


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
    numa = "unbound (single rank)"
    if world > 1 and os.environ.get("MOJO_BENCH_NUMA", "1") != "0":
        numa = bind_to_gpu_numa_node(local_rank)

    os.environ["MOJO_BACKEND"] = "b200"
    os.environ.setdefault("MOJO_B200_GAR_TIMEOUT_S", "60")
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

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

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
        return max_over_ranks(start.elapsed_time(stop))

    # ---- value: inputs resident in HBM
    def resident_step(i, is_timed):
        L = i % args.layers
        ev = None
        if is_timed:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            decode_events.append(ev)
        layer_step(res, res_tables[L], res_metas[L], caches[L][0], caches[L][1], ev)

    # The step is launched the way a serving loop launches it: one captured device graph per rotating KV cache
    # (mojo_opset_b200.runtime.DeviceGraphRunner, mirror of the reference's compile/device_graph.py).  `value` is the
    # graph-replayed step; the same K steps are ALSO timed launched one by one with CUDA events around every decode
    # launch - that pass feeds `roofline` and is reported as `eager`.  --eager makes it the `value`.
    # Every number is the MEDIAN of `--repeats` timed regions of exactly K steps each (all samples are reported).
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

    sustained = None
    with ClockSampler(local_rank) as clocks:
        eager_samples = [timed(resident_step, args.steps, args.warmup) for _ in range(args.repeats)]
        ms_eager = median(eager_samples)
        resident_samples = eager_samples
        if not args.eager:
            resident_samples = [timed(graph_step, args.steps, args.warmup) for _ in range(args.repeats)]
        ms_resident = median(resident_samples)
        # the same step for >= `--sustain-s` seconds: clocks under load, power cap, thermal state
        if args.sustain_s > 0:
            fn = (lambda i, t: resident_step(i, False)) if args.eager else graph_step
            n_sus = max(args.steps, int(args.sustain_s * 1e3 / (ms_resident / args.steps)))
            ms_sus = timed(fn, n_sus, args.warmup)
            sustained = {"value": B * world * n_sus / (ms_sus * 1e-3), "unit": "tokens/s", "steps": n_sus,
                         "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus}
    torch.cuda.synchronize()
    decode_ms = median([a.elapsed_time(b) for a, b in decode_events])  # per-launch CUDA-event time, all repeats

    # ---- e2e: host buffers in, host results out, every step, through ONE device graph per step.  The graph of
    # (slot j, layer L) holds: H2D of the layer's packed pinned input buffer (all per-step tensors, the block table
    # and the store plan) -> the five kernels -> D2H of the three results into slot j's pinned output buffers.  Steps
    # go round-robin over NSLOT streams (slot = stream), so step i+1's copy-in and step i-1's copy-out overlap step
    # i's kernels while the host issues one graph launch per step.  Before a slot is reused the host waits for the
    # slot's previous step (three steps back) - the point where a caller has consumed that step's results.
    NSLOT = 3
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
    dev_packed = [torch.empty(offset, dtype=torch.uint8, device=dev) for _ in range(NSLOT)]
    dev_views = [views(b) for b in dev_packed]
    out_host = [dict(
        y=torch.empty(B, cfg["hidden"], dtype=torch.bfloat16).pin_memory(),
        o=torch.empty(B, cfg["hq"], cfg["d"], dtype=torch.bfloat16).pin_memory(),
        a=torch.empty(B, cfg["inter"], dtype=torch.bfloat16).pin_memory()) for _ in range(NSLOT)]
    d2h_bytes = sum(t.numel() * t.element_size() for t in out_host[0].values())
    slot_streams = [torch.cuda.Stream() for _ in range(NSLOT)]
    slot_done = [torch.cuda.Event() for _ in range(NSLOT)]

    def e2e_body(j, L):
        d = dev_views[j]
        dev_packed[j].copy_(host_packed[L], non_blocking=True)
        y, _, o, a = layer_step(d, d["table"], d["meta"], caches[L][0], caches[L][1])
        out_host[j]["y"].copy_(y, non_blocking=True)
        out_host[j]["o"].copy_(o, non_blocking=True)
        out_host[j]["a"].copy_(a, non_blocking=True)

    e2e_graphs = None
    if not args.e2e_eager:
        e2e_graphs = {}
        for j in range(NSLOT):
            for L in range(args.layers):
                with torch.cuda.stream(slot_streams[j]):
                    e2e_body(j, L)  # warm-up outside the capture
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=slot_streams[j]):
                    e2e_body(j, L)
                e2e_graphs[(j, L)] = graph
        torch.cuda.synchronize()

    def e2e_step(i):
        j, L = i % NSLOT, i % args.layers
        slot_done[j].synchronize()  # the caller has consumed the results of the step that last used this slot
        with torch.cuda.stream(slot_streams[j]):
            if e2e_graphs is not None:
                e2e_graphs[(j, L)].replay()
            else:
                e2e_body(j, L)
            slot_done[j].record()

    def e2e_timed(steps, warmup):
        for i in range(warmup):
            e2e_step(i)
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        start.record()
        for s in slot_streams:
            s.wait_event(start)
        for i in range(steps):
            e2e_step(warmup + i)
        for ev in slot_done:
            main.wait_event(ev)
        stop.record()
        barrier()
        return max_over_ranks(start.elapsed_time(stop))

    for ev in slot_done:
        ev.record()
    e2e_samples = [e2e_timed(args.steps, args.warmup) for _ in range(args.repeats)]
    ms_e2e = median(e2e_samples)
    e2e_sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / (ms_e2e / args.steps)))
        ms_sus = e2e_timed(n_sus, args.warmup)
        e2e_sustained = {"value": B * world * n_sus / (ms_sus * 1e-3), "unit": "tokens/s", "steps": n_sus,
                         "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus}
    # the copies alone (no kernels): what the host link gives this rank while all ranks copy at once
    copy_only = measure_copy_only(host_packed, dev_packed, out_host, slot_streams, args.steps * 4, barrier,
                                  max_over_ranks, h2d_bytes, d2h_bytes)

    peaks = load_peaks()
    algo_bytes = decode_bytes(B, cfg["ctx"], cfg["hq"], cfg["hkv"], cfg["d"], cfg["bs"])
    achieved = algo_bytes / (decode_ms * 1e-3) / 1e9
    traffic, traffic_source = load_decode_traffic()

    tp_leg = dp5_leg = None
    ok = True
    if world > 1 and not args.no_extra:
        dp5_leg = run_dp_cfg5(m, dev, rank, world, peaks, barrier, max_over_ranks)
        torch.cuda.empty_cache()
        tp_leg, ok = run_tp_cfg4(m, dev, rank, world, peaks, args, barrier, max_over_ranks)

    if rank != 0:
        if world > 1:
            dist.barrier()
            from mojo_opset_b200.backends.b200.operators.compute_with_comm import release_workspaces

            release_workspaces()
            dist.destroy_process_group()
        if not ok:
            sys.exit(3)
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
        "sustained": sustained,
        "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "tokens/s",
                "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "pipeline": f"one device graph per step (H2D of the packed pinned inputs -> 5 kernels -> D2H of the "
                            f"results), round-robin over {NSLOT} streams; the host waits for a slot's previous step "
                            "before reusing it",
                "compute_launch": "eager (5 launches per step)" if args.e2e_eager else "CUDA graph replay per step",
                "sustained": e2e_sustained, "copies_only": copy_only},
        "gpu_launches": launches_per_step * args.steps,
        "timing": {"repeats": args.repeats, "stat": "median over repeats of a timed region of exactly `steps` steps",
                   "resident_ms_samples": [round(x, 4) for x in resident_samples],
                   "e2e_ms_samples": [round(x, 4) for x in e2e_samples]},
        "roofline": {"kernel": "paged_decode_mma_kernel (+ reduce)", "bound": "hbm", "achieved": achieved,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "frac_of_8TBs_nominal": achieved / 8000.0, "peak_source": peaks["source"],
                     "algorithmic_bytes": algo_bytes, "us_per_launch": decode_ms * 1e3, "traffic": traffic,
                     "traffic_source": traffic_source},
        "clocks": clocks.summary(),
    }
    if tp_leg is not None:
        line["tp_cfg4"] = tp_leg
    if dp5_leg is not None:
        line["dp_cfg5"] = dp5_leg
    if world == 1 and not args.no_extra:
        line["extra"] = run_extra(m, dev, peaks, args)
        # the metric names two rooflines: HBM for decode (`roofline`) and the bf16 tensor peak for prefill
        pf = line["extra"]["prefill_cfg3"]
        line["roofline_prefill"] = {
            "kernel": "attn_fwd_sm100_kernel (MojoPagedPrefillGQA, cfg3: T=8192 causal, 32q/8kv, hd 128, page 16)",
            "bound": "tensor", "achieved": pf["tflops"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": pf["frac_of_bf16_peak"], "frac_of_2250_nominal": pf["tflops"] / 2250.0,
            "peak_source": peaks["source"], "algorithmic_flops": pf["flops"], "us_per_launch": pf["ms"] * 1e3,
            "sustained": pf.get("sustained"), "traffic": None}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    if world > 1:
        dist.barrier()
        from mojo_opset_b200.backends.b200.operators.compute_with_comm import release_workspaces

        release_workspaces()
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


def measure_copy_only(host_packed, dev_packed, out_host, streams, steps, barrier, max_over_ranks, h2d_bytes, d2h_bytes):
    """The e2e leg's copies without its kernels: packed pinned H2D on one stream, the D2H of a result-sized buffer on
    another, `steps` times each, concurrently (and concurrently on every rank).  The per-step time is the floor the
    host link puts under an e2e step on this box."""
    dev = dev_packed[0].device
    d2h_src = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    d2h_dst = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    out = {}
    for mode in ("h2d", "d2h", "both"):
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        start.record()
        for s in streams[:2]:
            s.wait_event(start)
        for i in range(steps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(streams[0]):
                    dev_packed[i % len(dev_packed)].copy_(host_packed[i % len(host_packed)], non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(streams[1]):
                    d2h_dst.copy_(d2h_src, non_blocking=True)
        for s in streams[:2]:
            main.wait_stream(s)
        stop.record()
        barrier()
        ms = max_over_ranks(start.elapsed_time(stop)) / steps
        out[mode] = {"ms_per_step": ms,
                     "h2d_gbs_per_gpu": h2d_bytes / ms / 1e6 if mode != "d2h" else None,
                     "d2h_gbs_per_gpu": d2h_bytes / ms / 1e6 if mode != "h2d" else None}
    out["note"] = ("pinned-memory copies of the e2e step's byte counts alone, all ranks at once, max over ranks: the "
                   "host-link floor under an e2e step")
    return out


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


def _time_gpu_burst(fn, iters=10, ramp_ms=30.0):
    """Burst timing of ONE kernel the way MEASURED_PEAKS.json takes its burst figures: the GPU is first brought out
    of its idle clocks with `ramp_ms` of the same work, then `iters` back-to-back launches are timed one by one with
    CUDA events; returns (median, best) in ms."""
    fn()
    torch.cuda.synchronize()
    time.sleep(0.5)  # let the power / thermal state of whatever ran before decay: this is the kernel "timed alone"
    t0 = time.perf_counter()
    while (time.perf_counter() - t0) * 1e3 < ramp_ms:
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    evs[0].record()
    for i in range(iters):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    times = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(iters))
    return times[len(times) // 2], times[0]


def run_extra(m, dev, peaks, args):
    """cfg3 (prefill 8192, causal; burst of 10 launches and a >= 2 s sustained loop) and cfg5 (DiT SDPA: one GPU's
    share at 8 GPUs = 2 of 16 batch elements, and the whole batch)."""
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
    run_prefill = lambda: prefill(q, kc, vc, cu, table, max_q_len=T, max_total_seq_len=T)  # noqa: E731
    ms, ms_best = _time_gpu_burst(run_prefill)
    flops = 4 * Hq * D * (T * (T + 1) // 2)
    out["prefill_cfg3"] = {"workload": "MojoPagedPrefillGQA T=8192 causal 32q/8kv hd128 page16 bf16", "ms": ms,
                           "ms_best": ms_best, "tflops_best": flops / ms_best / 1e9,
                           "timing": "30 ms ramp out of idle clocks, then 10 launches timed one by one: median (best)",
                           "tflops": flops / ms / 1e9, "frac_of_bf16_peak": flops / ms / 1e9 / peaks["bf16_tflops"],
                           "tokens_per_s": T / (ms * 1e-3), "flops": flops}
    H, S = CFG5["heads"], CFG5["seq"]
    sdpa = m.MojoSdpa()
    for key, Bd, note in (("sdpa_cfg5_per_gpu", 2, "2 of 16 batch elements (one GPU's share at 8 GPUs)"),
                          ("sdpa_cfg5_whole", 16, "all 16 batch elements on one GPU")):
        qs, ks, vs = (torch.empty(Bd, S, H, D, dtype=torch.bfloat16, device=dev).normal_().transpose(1, 2)
                      for _ in range(3))
        ms, ms_best = _time_gpu_burst(lambda: sdpa(qs, ks, vs), 10 if Bd == 2 else 6)
        flops = 4 * Bd * H * S * S * D
        out[key] = {"workload": f"MojoSdpa DiT 24 heads hd128 S=4096 non-causal bf16, {note}, transposed-BSHD views",
                    "ms": ms, "ms_best": ms_best, "tflops": flops / ms / 1e9, "tflops_best": flops / ms_best / 1e9,
                    "frac_of_bf16_peak": flops / ms / 1e9 / peaks["bf16_tflops"], "flops": flops}
        del qs, ks, vs
    # head_dim 64 (reference SDPA accepts {64, 128}) on the same tcgen05 kernel: exponentials per FLOP double
    qs, ks, vs = (torch.empty(2, S, H, 64, dtype=torch.bfloat16, device=dev).normal_().transpose(1, 2) for _ in range(3))
    ms, ms_best = _time_gpu_burst(lambda: sdpa(qs, ks, vs), 10)
    flops = 4 * 2 * H * S * S * 64
    out["sdpa_d64"] = {"workload": "MojoSdpa 24 heads hd64 S=4096 non-causal bf16, batch 2", "ms": ms, "ms_best": ms_best,
                       "tflops": flops / ms / 1e9, "tflops_best": flops / ms_best / 1e9,
                       "frac_of_bf16_peak": flops / ms / 1e9 / peaks["bf16_tflops"], "flops": flops}
    del qs, ks, vs
    out["decode_small"] = run_decode_small(m, dev, peaks)
    if args.sustain_s > 0:
        ms, flops = out["prefill_cfg3"]["ms"], out["prefill_cfg3"]["flops"]
        iters = max(10, int(2.0 * args.sustain_s * 1e3 / ms))
        ms_s = _time_gpu(run_prefill, iters, warmup=1)
        sus_peak = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
        out["prefill_cfg3"]["sustained"] = {"seconds": ms_s * iters * 1e-3, "launches": iters, "ms": ms_s,
                                            "tflops": flops / ms_s / 1e9, "peak_sustained": sus_peak,
                                            "frac_of_bf16_sustained_peak": flops / ms_s / 1e9 / sus_peak}
    return out


def run_decode_small(m, dev, peaks):
    """Small-batch split-KV decode (the batch-1 / batch-4 serving cases): 20 launches replayed from one CUDA graph."""
    out = {}
    decode = m.MojoPagedDecodeGQA()
    Hq, Hkv, D, bs = 32, 8, 128, 16
    for name, B, ctx in (("b1_ctx32k", 1, 32768), ("b4_ctx8k", 4, 8192)):
        nb = B * ctx // bs + 10
        kc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=dev).normal_()
        vc = torch.empty(nb, Hkv, bs, D, dtype=torch.bfloat16, device=dev).normal_()
        q = torch.empty(B, Hq, D, dtype=torch.bfloat16, device=dev).normal_()
        table = torch.randperm(nb)[: B * ctx // bs].view(B, -1).to(torch.int32).to(dev)
        lens = torch.full((B,), ctx, dtype=torch.int32, device=dev)
        fn = lambda: decode(q, kc, vc, lens, table, max_total_seq_len=ctx)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 20)
        nbytes = decode_bytes(B, ctx, Hq, Hkv, D, bs)
        gbs = nbytes / best / 1e6
        out[name] = {"workload": f"MojoPagedDecodeGQA batch {B}, 32q/8kv hd128 page16 ctx {ctx} bf16 (split-KV)",
                     "us": best * 1e3, "gbs": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"],
                     "algorithmic_bytes": nbytes}
        del kc, vc
    return out


# ------------------------------------------------------------------------------------------------------
# multi-GPU legs: cfg5 sharded by batch, cfg4 tensor parallel with the fused o_proj GEMM + all-reduce
# ------------------------------------------------------------------------------------------------------
def run_dp_cfg5(m, dev, rank, world, peaks, barrier, max_over_ranks):
    from mojo_opset_b200.parallel import shard_range

    Bt, H, S, D = CFG5["batch"], CFG5["heads"], CFG5["seq"], CFG5["d"]
    lo, hi = shard_range(Bt, world, rank)
    Bl = hi - lo
    gen = torch.Generator(device=dev).manual_seed(20260716 + 5 + rank)
    qs, ks, vs = (torch.empty(Bl, S, H, D, dtype=torch.bfloat16, device=dev).normal_(generator=gen).transpose(1, 2)
                  for _ in range(3))
    sdpa = m.MojoSdpa()
    for _ in range(3):
        sdpa(qs, ks, vs)
    iters = 10
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        sdpa(qs, ks, vs)
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b)) / iters
    flops_total = 4 * Bt * H * S * S * D
    return {"workload": f"cfg5 DiT MojoSdpa 24 heads hd128 S=4096 non-causal bf16, batch 16 sharded by batch: "
                        f"{Bl} elements per rank on {world} GPUs, no collective",
            "ms_per_pass": ms, "tflops_total": flops_total / ms / 1e9,
            "tflops_per_gpu": flops_total / world / ms / 1e9,
            "frac_of_bf16_peak_per_gpu": flops_total / world / ms / 1e9 / peaks["bf16_tflops"],
            "batch_elements_per_s": Bt / (ms * 1e-3), "scaling": "strong", "timing": "CUDA events, max over ranks"}


def run_tp_cfg4(m, dev, rank, world, peaks, args, barrier, max_over_ranks):
    """One Llama-3-70B-shaped decode layer step per rank of a TP group of `world` ranks (reference sharding rule:
    q Shard(-2), caches Shard(-3), o_proj RowwiseParallel - distributed/parallel/partitions.py:42-47,85-89; fused op:
    core/operators/compute_with_comm.py:57-117).  Returns (json leg, parity ok)."""
    import torch.distributed as dist

    from mojo_opset_b200.parallel import shard_heads
    from oracle import golden  # the checker of the in-run parity test (test infrastructure, not the measured path)

    c = CFG4
    B, D, bs, ctx, hidden = c["batch"], c["d"], c["bs"], c["ctx"], c["hidden"]
    shard = shard_heads(c["hq"], c["hkv"], world, rank)
    hq_l, hkv_l = shard.q_end - shard.q_begin, shard.kv_end - shard.kv_begin
    blocks_per_seq = ctx // bs
    nb = B * blocks_per_seq + 8
    g = torch.Generator().manual_seed(20260716 + 4)       # same table / plan on every rank (one batch, sharded heads)
    gen = torch.Generator(device=dev).manual_seed(20260716 + 40 + rank)
    kc = torch.empty(nb, hkv_l, bs, D, dtype=torch.bfloat16, device=dev).normal_(generator=gen)
    vc = torch.empty(nb, hkv_l, bs, D, dtype=torch.bfloat16, device=dev).normal_(generator=gen)
    perm = torch.randperm(nb, generator=g)[: B * blocks_per_seq].view(B, blocks_per_seq).to(torch.int32)
    meta = torch.stack((torch.arange(B, dtype=torch.int32), perm[:, (ctx - 1) // bs],
                        torch.full((B,), (ctx - 1) % bs, dtype=torch.int32), torch.ones(B, dtype=torch.int32)), -1)
    table, meta = perm.to(dev), meta.contiguous().to(dev)
    rnd = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=dev).normal_(generator=gen)  # noqa: E731
    q, k_new, v_new = rnd(B, hq_l, D), rnd(B, hkv_l, D), rnd(B, hkv_l, D)
    lens = torch.full((B,), ctx, dtype=torch.int32, device=dev)
    w_o = rnd(hidden, hq_l * D) / math.sqrt(c["hq"] * D)          # this rank's column shard of o_proj, K-major
    store, decode = m.MojoStorePagedKVCache(), m.MojoPagedDecodeGQA()
    gar = m.MojoGemmAllReduce(w_o, None, trans_weight=False)
    assert type(gar).__name__ == "B200GemmAllReduce"

    def attn():
        store(k_new, v_new, kc, vc, chunk_metadata=meta)
        return decode(q, kc, vc, lens, table, max_total_seq_len=ctx)

    def step_fused():
        return gar(attn().view(B, hq_l * D))

    def step_nocomm():
        return torch.nn.functional.linear(attn().view(B, hq_l * D), w_o)

    def step_unfused():
        y = torch.nn.functional.linear(attn().view(B, hq_l * D), w_o)
        dist.all_reduce(y)
        return y

    # ---- in-run parity: decode of two sequences and the fused projection against the oracle
    o = attn()
    y = step_fused()
    torch.cuda.synchronize()
    ok, notes = True, []
    o_ref = golden.paged_decode_gqa(q[:2], kc, vc, lens[:2], table[:2])
    err_o = (o[:2].float() - o_ref.float()).norm().item() / max(o_ref.float().norm().item(), 1e-30)
    part = golden.gemm(o.view(B, hq_l * D), w_o).float()          # this rank's projection, rounded to bf16 like F.linear
    dist.all_reduce(part)                                         # NCCL fp32 sum of the ranks' bf16 partials (plumbing)
    y_ref = part.to(torch.bfloat16)
    # relative measures: the outputs are O(1e-2) (a 32k-key softmax averages thousands of values), so an absolute
    # 2e-2 alone would accept anything
    tol = 1e-2
    err_y = (y.float() - y_ref.float()).norm().item() / max(y_ref.float().norm().item(), 1e-30)
    y_hi, y_lo = y.float().clone(), y.float().clone()
    dist.all_reduce(y_hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(y_lo, op=dist.ReduceOp.MIN)
    same = bool(torch.equal(y_hi, y_lo))
    bad = torch.tensor([int(err_o > 1.5e-2) + 2 * int(err_y > tol) + 4 * int(not same)], device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX)
    if bad.item():
        ok = False
        notes.append(f"PARITY FAILURE code {int(bad.item())}: decode rel err {err_o:.3e}, fused rel err {err_y:.3e} "
                     f"(tol {tol:.3e}), ranks identical {same}")
        print(notes[-1], file=sys.stderr, flush=True)

    def timed_loop(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / steps

    def graphed(fn):
        fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        return graph.replay

    K, W = args.steps, args.warmup
    # the four variants are timed INTERLEAVED (fused, no-collective, attention only, cuBLAS + NCCL; five rounds, median
    # per variant): the differences between them are tens of microseconds of a multi-millisecond step, less than the
    # drift of a GPU warming up over back-to-back regions of one variant each
    variants = {"fused": graphed(step_fused), "nocomm": graphed(step_nocomm), "attn": graphed(attn),
                "unfused": step_unfused}      # NCCL launched eagerly
    samples = {name: [] for name in variants}
    for _ in range(5):
        for name, fn in variants.items():
            samples[name].append(timed_loop(fn, K, W if not samples[name] else 2))
    ms_fused, ms_nocomm, ms_attn, ms_unfused = (median(samples[n]) for n in ("fused", "nocomm", "attn", "unfused"))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K)]

    def attn_timed_once(i):
        store(k_new, v_new, kc, vc, chunk_metadata=meta)
        ev[2 * i].record()
        decode(q, kc, vc, lens, table, max_total_seq_len=ctx)
        ev[2 * i + 1].record()

    for i in range(K):
        attn_timed_once(i)
    torch.cuda.synchronize()
    dec_ms = max_over_ranks(median([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(K)]))
    kv_bytes = decode_bytes(B, ctx, hq_l, hkv_l, D, bs)
    gbs = kv_bytes / (dec_ms * 1e-3) / 1e9
    leg = {
        "workload": f"cfg4 Llama-3-70B-shaped decode layer step: batch {B}, {c['hq']}q/{c['hkv']}kv heads, hd 128, "
                    f"ctx {ctx}, page 16, bf16; KV heads sharded TP={world} (local {hq_l}q/{hkv_l}kv): "
                    f"StorePagedKVCache -> PagedDecodeGQA -> MojoGemmAllReduce (o_proj [{B},{hq_l * D}]x[{hidden}]"
                    " + all-reduce, one kernel over NVLink peer memory)",
        "tp": world, "steps": K, "warmup": W, "launch": "one CUDA graph per step (fused); NCCL baseline eager",
        "ms_per_step": ms_fused, "tokens_per_s": B / (ms_fused * 1e-3),
        "ms_per_step_cublas_nccl": ms_unfused, "tokens_per_s_cublas_nccl": B / (ms_unfused * 1e-3),
        "ms_per_step_no_collective": ms_nocomm, "ms_attention_only": ms_attn,
        "allreduce_exposed_us": (ms_fused - ms_nocomm) * 1e3,
        "allreduce_exposed_us_cublas_nccl": (ms_unfused - ms_nocomm) * 1e3,
        "gemm_allreduce_us": (ms_fused - ms_attn) * 1e3, "cublas_plus_nccl_us": (ms_unfused - ms_attn) * 1e3,
        "decode_us_max_rank": dec_ms * 1e3, "decode_bytes_per_rank": kv_bytes, "decode_gbs_per_rank": gbs,
        "decode_frac_of_measured_hbm": gbs / peaks["hbm_gbs"],
        "allreduce_bytes": B * hidden * 2,
        # roofline of the fused compute + collective kernel (B200_PROFILING.md): the slower of FLOPs / measured GEMM peak
        # and the bytes that must cross NVLink (two-shot all-reduce payload per GPU and direction) / 770 GB/s measured
        "fused_roofline": (lambda flops, nvl: {
            "gemm_flops": flops, "gemm_floor_us": flops / (peaks["bf16_tflops"] * 1e6),
            "nvlink_bytes_per_gpu_per_direction": nvl, "nvlink_gbs": 770.0, "nvlink_floor_us": nvl / 770e3,
            "floor_us": max(flops / (peaks["bf16_tflops"] * 1e6), nvl / 770e3),
            "fused_us_in_step": (ms_fused - ms_attn) * 1e3,
            "frac": max(flops / (peaks["bf16_tflops"] * 1e6), nvl / 770e3) / max((ms_fused - ms_attn) * 1e3, 1e-9),
            "frac_cublas_nccl": max(flops / (peaks["bf16_tflops"] * 1e6), nvl / 770e3) / max((ms_unfused - ms_attn) * 1e3, 1e-9),
        })(2.0 * B * hidden * hq_l * D, 2.0 * (world - 1) / world * B * hidden * 2),
        "parity": {"ok": ok, "decode_rel_err_vs_oracle": err_o, "fused_rel_err_vs_oracle": err_y,
                   "fused_rel_tol": tol, "decode_rel_tol": 1.5e-2, "ranks_bit_identical": same, "notes": notes},
        "timing": "CUDA events, the four variants interleaved, median of 5 regions of K steps each, max over ranks",
        "ms_samples": {k: [round(x, 4) for x in v] for k, v in samples.items()},
    }
    return leg, ok


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own torch-native ops (baseline/_ref) - else the oracle port - on the host cores
# ------------------------------------------------------------------------------------------------------
def load_reference_ops(cfg):
    """(kind, step(host, caches, layer)) - kind "reference": the UNMODIFIED reference package installed in
    baseline/_ref (MOJO_BACKEND=torch: its torch-native golden backend); "port": oracle/golden.py."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_root, "mojo_opset")) and os.environ.get("MOJO_BENCH_FORCE_PORT") != "1":
        sys.path.insert(0, ref_root)
        os.environ["MOJO_BACKEND"] = "torch"
        os.environ["MOJO_OPSET_PLUGIN_AUTOLOAD"] = "0"
        os.environ.pop("MOJO_DISABLE_ASSERTION_REWRITE", None)
        import mojo_opset as ref

        norm = ref.MojoResidualAddRMSNorm(cfg["hidden"], eps=cfg["eps"], dtype=torch.bfloat16)
        rope, store, decode, swiglu = (ref.MojoApplyRoPE(), ref.MojoStorePagedKVCache(), ref.MojoPagedDecodeGQA(),
                                       ref.MojoSwiGLU())
        assert type(decode).__name__ == "TorchPagedDecodeGQA" and ref.__file__.startswith(ref_root)
        state = {"init": False}

        def step(host, caches, layer):
            if not state["init"]:
                with torch.no_grad():
                    norm.weight.copy_(host["norm_weight"])
                state["init"] = True
            kc, vc = caches[layer]
            y, r = norm(host["hidden"], host["residual"])
            q_rot, k_rot = rope(host["q"], host["k"], host["cos"], host["sin"], head_first=False)
            store(k_rot, host["v"], kc, vc, chunk_metadata=host["metas"][layer])
            o = decode(q_rot, kc, vc, host["total_seq_lens"], host["tables"][layer], max_total_seq_len=cfg["ctx"])
            a = swiglu(host["gate"], host["up"])
            return y, r, o, a

        return "reference", step, f"unmodified reference (baseline/_ref, byted-mojo-opset), backend torch, CPU"

    from oracle import golden

    def step(host, caches, layer):
        kc, vc = caches[layer]
        y, r = golden.residual_add_rms_norm(host["hidden"], host["residual"], host["norm_weight"], cfg["eps"])
        q_rot, k_rot = golden.apply_rope(host["q"], host["k"], host["cos"], host["sin"], head_first=False)
        golden.store_paged_kv(k_rot, host["v"], kc, vc, host["metas"][layer])
        o = golden.paged_decode_gqa(q_rot, kc, vc, host["total_seq_lens"], host["tables"][layer])
        a = golden.swiglu(host["gate"], host["up"])
        return y, r, o, a

    return "port", step, "oracle/golden.py (restatement of the reference's torch-native ops; baseline/_ref absent), CPU"


def cpu_baseline():
    """The reference arm run as a subprocess on a bounded number of steps (1 warm-up + 2 timed steps of the FULL cfg2
    batch: ~10-20 s of CPU work), so the reference package and its import hooks stay out of this process."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MOJO_BACKEND")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        line = json.loads(res.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "unavailable",
                "sample": f"reference arm subprocess failed: {type(e).__name__}: {e}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    torch.set_num_threads(cores)
    kind, step, what = load_reference_ops(CFG2)
    B = args.cpu_sample_batch or CFG2["batch"]
    host, caches = make_decode_inputs(CFG2, B, 1, 20260716 + 2, "cpu")
    for _ in range(args.warmup):
        step(host, caches, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(host, caches, 0)
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    full = B == CFG2["batch"]
    sample = (f"{'the full cfg2 batch (64 sequences)' if full else f'{B} of 64 sequences of cfg2'} per step, "
              f"{args.steps} steps after {args.warmup} warm-up; {what}, {torch.get_num_threads()} threads")
    config = {"workload": WORKLOAD, "batch_per_gpu": CFG2["batch"], "ctx": CFG2["ctx"], "block_size": CFG2["bs"],
              "parallelism": "host cores (rank 0 only)"}
    if not full:
        config["sample_batch_per_step"] = B
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "reference",
        "config": config,
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample, "seconds_per_step": dt / args.steps},
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
    ap.add_argument("--sustain-s", type=float, default=1.0, help="length of the sustained legs in seconds (0: skip)")
    ap.add_argument("--cpu-sample-batch", type=int, default=0,
                    help="reference arm: sequences per step (default: the full cfg2 batch of 64)")
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
