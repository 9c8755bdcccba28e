#!/usr/bin/env python
"""A Qwen3-shaped decoder driven end to end through the b200 backend: the drop-in story of the reference's
``examples/llm_inference.py`` / ``modeling/qwen3/mojo_qwen3_dense.py`` without a checkpoint (the GPU box has neither
network nor weights - SURVEY.md appendix B - so weights are random bf16 of the real shapes).

Per decode step, all inside ONE CUDA graph with no host read (``mojo_opset_b200.runtime``):

    PagedAttentionRuntimeState.prepare_decode_inputs  (device-side block allocator)
    embedding -> L x [ MojoResidualAddRMSNorm -> qkv projection (cuBLAS) -> MojoNormRoPEStoreKV (q/k-norm + RoPE +
    paged KV store, one kernel) -> MojoPagedDecodeGQA -> o_proj -> MojoResidualAddRMSNorm -> gate/up projection ->
    MojoSwiGLU -> down projection ] -> MojoResidualAddRMSNorm -> lm_head -> argmax

The GEMMs are library calls (cuBLAS through ``torch.nn.functional.linear``): plumbing around the hand-written path.
``torch`` golden twins of the same arithmetic live in ``reference_forward`` for the parity test
(``tests/test_gpu_runtime.py``).

    python examples/qwen3_synthetic.py --layers 36 --batch 64 --context 4096 --steps 32
"""

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Qwen3Config:
    def __init__(self, hidden_size=4096, num_layers=36, num_heads=32, num_kv_heads=8, head_dim=128,
                 intermediate_size=12288, vocab_size=151936, rms_norm_eps=1e-6, rope_theta=1e6,
                 max_position_embeddings=8192):
        self.hidden_size, self.num_layers, self.num_heads, self.num_kv_heads = hidden_size, num_layers, num_heads, num_kv_heads
        self.head_dim, self.intermediate_size, self.vocab_size = head_dim, intermediate_size, vocab_size
        self.rms_norm_eps, self.rope_theta, self.max_position_embeddings = rms_norm_eps, rope_theta, max_position_embeddings


class Qwen3Layer(torch.nn.Module):
    def __init__(self, cfg, ops, device, dtype, gen):
        super().__init__()
        H, D, Hq, Hkv, I = cfg.hidden_size, cfg.head_dim, cfg.num_heads, cfg.num_kv_heads, cfg.intermediate_size

        def w(out_f, in_f):
            return (torch.randn(out_f, in_f, generator=gen, dtype=torch.float32) / in_f ** 0.5).to(dtype).to(device)

        self.w_qkv, self.w_o = w((Hq + 2 * Hkv) * D, H), w(H, Hq * D)
        self.w_gate_up, self.w_down = w(2 * I, H), w(H, I)
        self.input_norm = ops.MojoResidualAddRMSNorm(H, eps=cfg.rms_norm_eps, device=device, dtype=dtype)
        self.post_norm = ops.MojoResidualAddRMSNorm(H, eps=cfg.rms_norm_eps, device=device, dtype=dtype)
        self.qk_rope_store = ops.MojoNormRoPEStoreKV(D, eps=cfg.rms_norm_eps, device=device, dtype=dtype)
        self.decode_attn, self.prefill_attn = ops.MojoPagedDecodeGQA(), ops.MojoPagedPrefillGQA()
        self.swiglu = ops.MojoSwiGLU()
        with torch.no_grad():
            for p in (self.input_norm.weight, self.post_norm.weight, self.qk_rope_store.q_weight,
                      self.qk_rope_store.k_weight):
                p.copy_((1 + 0.1 * torch.randn(p.shape, generator=gen)).to(dtype))
        self.cfg = cfg


class Qwen3Synthetic(torch.nn.Module):
    def __init__(self, cfg: Qwen3Config, device="cuda", dtype=torch.bfloat16, seed=0):
        super().__init__()
        os.environ.setdefault("MOJO_BACKEND", "b200")
        import mojo_opset_b200 as ops

        gen = torch.Generator().manual_seed(seed)
        self.cfg, self.device, self.dtype = cfg, device, dtype
        self.embed = (torch.randn(cfg.vocab_size, cfg.hidden_size, generator=gen) * 0.5).to(dtype).to(device)
        self.lm_head = (torch.randn(cfg.vocab_size, cfg.hidden_size, generator=gen) / cfg.hidden_size ** 0.5).to(dtype).to(device)
        self.layers = torch.nn.ModuleList(Qwen3Layer(cfg, ops, device, dtype, gen) for _ in range(cfg.num_layers))
        self.final_norm = ops.MojoResidualAddRMSNorm(cfg.hidden_size, eps=cfg.rms_norm_eps, device=device, dtype=dtype)
        with torch.no_grad():
            self.final_norm.weight.copy_((1 + 0.1 * torch.randn(cfg.hidden_size, generator=gen)).to(dtype))
        self.rotary = ops.MojoRotaryEmbedding(cfg.rope_theta, cfg.head_dim, device=device)
        self.scale = cfg.head_dim ** -0.5

    @torch.inference_mode()
    def forward(self, input_ids, positions, meta):
        """``input_ids [T]``, ``positions [T]``, ``meta`` from the runtime state; returns logits ``[T, vocab]``."""
        cfg = self.cfg
        Hq, Hkv, D, I = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.intermediate_size
        F = torch.nn.functional
        hidden = self.embed[input_ids]
        residual = torch.zeros_like(hidden)
        cos, sin = self.rotary(hidden, position_ids=positions.to(torch.int32))
        T = hidden.shape[0]
        max_q = T  # grid-sizing hint only
        for li, layer in enumerate(self.layers):
            x, residual = layer.input_norm(hidden, residual)
            qkv = F.linear(x, layer.w_qkv).view(T, Hq + 2 * Hkv, D)
            q_rot = layer.qk_rope_store(qkv[:, :Hq], qkv[:, Hq:Hq + Hkv], qkv[:, Hq + Hkv:], cos, sin,
                                        meta.key_caches[li], meta.value_caches[li], meta.block_tables,
                                        meta.cu_q_lens, meta.context_kv_lens)
            if meta.is_prefill:
                cu_total = F.pad(meta.total_seq_lens.cumsum(-1, dtype=torch.int32), (1, 0))
                attn = layer.prefill_attn(q_rot, meta.key_caches[li], meta.value_caches[li], meta.cu_q_lens,
                                          meta.block_tables, self.scale, cu_total_seq_lens=cu_total, max_q_len=max_q)
            else:
                attn = layer.decode_attn(q_rot, meta.key_caches[li], meta.value_caches[li], meta.total_seq_lens,
                                         meta.block_tables, self.scale,
                                         max_total_seq_len=cfg.max_position_embeddings)
            hidden = F.linear(attn.view(T, Hq * D), layer.w_o)
            x, residual = layer.post_norm(hidden, residual)
            gate_up = F.linear(x, layer.w_gate_up)
            hidden = F.linear(layer.swiglu(gate_up[:, :I], gate_up[:, I:]), layer.w_down)
        final, _ = self.final_norm(hidden, residual)
        return F.linear(final, self.lm_head)


def reference_forward(model: Qwen3Synthetic, golden, input_ids, positions, q_lens, context_lens, block_tables,
                      key_caches, value_caches, is_prefill):
    """The same network with the ORACLE's ops on CPU tensors (test infrastructure: called by the parity test only)."""
    cfg = model.cfg
    Hq, Hkv, D, I = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.intermediate_size
    F = torch.nn.functional
    cpu = lambda t: t.detach().cpu()  # noqa: E731
    hidden = cpu(model.embed)[input_ids]
    residual = torch.zeros_like(hidden)
    inv_freq = cpu(model.rotary.inv_freq)
    cos, sin = golden.rotary_cos_sin(positions.to(torch.float32), inv_freq)
    T = hidden.shape[0]
    cu_q = None
    if is_prefill:
        cu_q = F.pad(q_lens.cumsum(-1, dtype=torch.int32), (1, 0))
    total = context_lens + q_lens
    for li, layer in enumerate(model.layers):
        x, residual = golden.residual_add_rms_norm(hidden, residual, cpu(layer.input_norm.weight), cfg.rms_norm_eps)
        qkv = F.linear(x, cpu(layer.w_qkv)).view(T, Hq + 2 * Hkv, D)
        q_rot, _ = golden.norm_rope_store_kv(qkv[:, :Hq], qkv[:, Hq:Hq + Hkv], qkv[:, Hq + Hkv:], cos, sin,
                                             key_caches[li], value_caches[li], block_tables, cu_q, context_lens,
                                             cpu(layer.qk_rope_store.q_weight), cpu(layer.qk_rope_store.k_weight),
                                             cfg.rms_norm_eps)
        if is_prefill:
            cu_total = F.pad(total.cumsum(-1, dtype=torch.int32), (1, 0))
            attn = golden.paged_prefill_gqa(q_rot, key_caches[li], value_caches[li], cu_q, block_tables, model.scale,
                                            cu_total)
        else:
            attn = golden.paged_decode_gqa(q_rot, key_caches[li], value_caches[li], total, block_tables, model.scale)
        hidden = F.linear(attn.reshape(T, Hq * D), cpu(layer.w_o))
        x, residual = golden.residual_add_rms_norm(hidden, residual, cpu(layer.post_norm.weight), cfg.rms_norm_eps)
        gate_up = F.linear(x, cpu(layer.w_gate_up))
        hidden = F.linear(golden.swiglu(gate_up[:, :I], gate_up[:, I:]), cpu(layer.w_down))
    final, _ = golden.residual_add_rms_norm(hidden, residual, cpu(model.final_norm.weight), cfg.rms_norm_eps)
    return F.linear(final, cpu(model.lm_head))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=36)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--context", type=int, default=4096, help="tokens already in the KV cache when decoding starts")
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--block-size", type=int, default=16)
    ap.add_argument("--vocab", type=int, default=151936)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    from mojo_opset_b200.runtime import DeviceGraphRunner
    from mojo_opset_b200.runtime import PagedAttentionRuntimeState

    dev, dtype = "cuda", torch.bfloat16
    cfg = Qwen3Config(num_layers=args.layers, vocab_size=args.vocab,
                      max_position_embeddings=args.context + args.steps + 64)
    model = Qwen3Synthetic(cfg, dev, dtype)
    state = PagedAttentionRuntimeState.from_config(cfg, args.batch, dev, dtype, block_size=args.block_size)
    # synthetic context: reserve `context` tokens per sequence and fill their pages with noise (no prompt to read)
    state._reserve(torch.full((args.batch,), args.context, dtype=torch.int32))
    for kc, vc in zip(state.key_caches, state.value_caches):
        kc.normal_()
        vc.normal_()
    state.check()
    ids = torch.randint(0, cfg.vocab_size, (args.batch,), device=dev)

    def step(input_ids):
        input_ids, positions, meta = state.prepare_decode_inputs(input_ids)
        return model(input_ids, positions, meta).argmax(-1)

    if args.no_graph:
        run = step
    else:
        runner = DeviceGraphRunner(step)
        runner.capture(ids, session=state)
        run = runner.replay
    for _ in range(3):
        ids = run(ids).clone()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(args.steps):
        ids = run(ids).clone()
    b.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    state.check()
    ms = a.elapsed_time(b) / args.steps
    weights_gb = sum(p.numel() * p.element_size() for l in model.layers for p in (l.w_qkv, l.w_o, l.w_gate_up, l.w_down)) / 1e9
    kv_gb = 2 * args.batch * args.context * cfg.num_kv_heads * cfg.head_dim * 2 * cfg.num_layers / 1e9
    print(json.dumps({
        "workload": f"Qwen3-8B-shaped synthetic decode: {args.layers} layers, batch {args.batch}, context {args.context}, "
                    f"page {args.block_size}, bf16, random weights, CUDA graph={'no' if args.no_graph else 'yes'}",
        "ms_per_step": ms, "tokens_per_s": args.batch / (ms * 1e-3), "wall_ms_per_step": wall / args.steps * 1e3,
        "weights_gb": weights_gb + 2 * cfg.vocab_size * cfg.hidden_size * 2 / 1e9, "kv_read_gb_per_step": kv_gb,
        "hbm_gbs_weights_plus_kv": (weights_gb + cfg.vocab_size * cfg.hidden_size * 2 / 1e9 + kv_gb) / (ms * 1e-3),
    }), flush=True)


if __name__ == "__main__":
    main()
