#!/usr/bin/env python
"""One Wan2.2-shaped DiT attention block driven through the Mojo ops: the drop-in story of the reference's
``examples/dit_inference.py`` without the external Wan2.2 repo, ``diffusers`` or a checkpoint (none of which exist on
the GPU box - SURVEY.md appendix B).  The module structure follows ``modeling/wan2_2/mojo_wan_model.py:39-187``
(``WanSelfAttention`` / ``WanCrossAttention`` / ``WanAttentionBlock``): which Mojo op sits where, which tensors are
transposed views, the six-way modulation - with random weights of the TI2V-5B shape (dim 3072, 24 heads of 128,
ffn 14336, 4096 image tokens, 512 text tokens).

    LayerNorm (no affine) -> modulate -> q/k/v (cuBLAS) -> RMSNorm(q), RMSNorm(k) -> GridRoPE(q), GridRoPE(k)
    -> MojoSdpa on transposed [B,S,H,D] views -> o -> gated residual -> cross attention (MojoSdpa, Skv = 512)
    -> LayerNorm -> modulate -> Linear, MojoGelu, Linear -> gated residual

Which kernels run is decided by ``MOJO_BACKEND`` when the block is built (b200: the hand-written CUDA path; the test
builds the same block on the oracle's ops for parity).  The Linear layers are library GEMMs: plumbing.

    python examples/dit_block_synthetic.py --batch 2 --steps 20
"""

import argparse
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rope_phases(max_len, dim, theta=10000.0):
    """Complex unit phases ``[max_len, dim/2]`` (reference ``rope_params``, ``mojo_wan_model.py:29-36``)."""
    inv = 1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim)
    ang = torch.outer(torch.arange(max_len, dtype=torch.float64), inv)
    return torch.polar(torch.ones_like(ang), ang).to(torch.complex64)


def grid_phase_table(grid, head_dim, device):
    """Per-sample phase table ``[F*H*W, 1, head_dim/2]`` for a (F, H, W) token grid: the head dimension is split into a
    temporal and two spatial bands (reference ``WanModel.__init__`` ``freqs`` + the per-sample expansion its forward
    does before calling ``MojoGridRoPE``)."""
    f, h, w = grid
    c = head_dim // 2
    bands = [c - 2 * (c // 3), c // 3, c // 3]
    tabs = [rope_phases(1024, 2 * b) for b in bands]
    ph = torch.cat([
        tabs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        tabs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
        tabs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1),
    ], dim=-1)
    return ph.reshape(f * h * w, 1, c).to(device)


class DiTAttention(nn.Module):
    """Self- or cross-attention of the block (reference ``WanSelfAttention`` / ``WanCrossAttention``)."""

    def __init__(self, ops, dim, num_heads, eps, device, dtype):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        kw = dict(device=device, dtype=dtype)
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim, **kw) for _ in range(4))
        self.norm_q = ops.MojoRMSNorm(dim, eps=eps, **kw)
        self.norm_k = ops.MojoRMSNorm(dim, eps=eps, **kw)
        self.sdpa = ops.MojoSdpa()
        self.grid_rope = ops.MojoGridRoPE()

    def forward(self, x, context=None, grid_sizes=None, freqs=None):
        b, n, d = x.size(0), self.num_heads, self.head_dim
        src = x if context is None else context
        q = self.norm_q(self.q(x)).view(b, -1, n, d)
        k = self.norm_k(self.k(src)).view(b, -1, n, d)
        v = self.v(src).view(b, -1, n, d)
        if context is None:  # image tokens carry their 3-D grid position
            q = self.grid_rope(q, grid_sizes, freqs)
            k = self.grid_rope(k, grid_sizes, freqs)
        out = self.sdpa(query=q.transpose(1, 2), key=k.transpose(1, 2), value=v.transpose(1, 2))
        return self.o(out.transpose(1, 2).contiguous().flatten(2))


class DiTBlock(nn.Module):
    """Reference ``WanAttentionBlock`` (``mojo_wan_model.py:128-187``)."""

    def __init__(self, ops, dim=3072, ffn_dim=14336, num_heads=24, eps=1e-6, device="cuda", dtype=torch.bfloat16,
                 seed=0):
        super().__init__()
        torch.manual_seed(seed)
        kw = dict(device=device, dtype=dtype)
        self.norm1 = ops.MojoLayerNorm(dim, eps, elementwise_affine=False, **kw)
        self.self_attn = DiTAttention(ops, dim, num_heads, eps, device, dtype)
        self.norm3 = ops.MojoLayerNorm(dim, eps, elementwise_affine=True, **kw)
        self.cross_attn = DiTAttention(ops, dim, num_heads, eps, device, dtype)
        self.norm2 = ops.MojoLayerNorm(dim, eps, elementwise_affine=False, **kw)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim, **kw), ops.MojoGelu(), nn.Linear(ffn_dim, dim, **kw))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim, **kw) / dim ** 0.5)
        with torch.no_grad():  # the op constructors leave norm weights uninitialised (as the reference's do)
            for m in (self.self_attn, self.cross_attn):
                m.norm_q.weight.copy_(1 + 0.1 * torch.randn(dim, **kw))
                m.norm_k.weight.copy_(1 + 0.1 * torch.randn(dim, **kw))
            self.norm3.weight.copy_(1 + 0.1 * torch.randn(dim, **kw))
            self.norm3.bias.copy_(0.1 * torch.randn(dim, **kw))

    def forward(self, x, e, grid_sizes, freqs, context):
        e = (self.modulation.unsqueeze(0) + e).chunk(6, dim=2)
        y = self.self_attn(self.norm1(x) * (1 + e[1].squeeze(2)) + e[0].squeeze(2), None, grid_sizes, freqs)
        x = x + y * e[2].squeeze(2)
        x = x + self.cross_attn(self.norm3(x), context)
        y = self.ffn(self.norm2(x) * (1 + e[4].squeeze(2)) + e[3].squeeze(2))
        return x + y * e[5].squeeze(2)


def make_inputs(batch, grid, text_len, dim, head_dim, device, dtype, seed=1):
    g = torch.Generator().manual_seed(seed)
    L = grid[0] * grid[1] * grid[2]
    x = torch.randn(batch, L, dim, generator=g).to(dtype).to(device)
    e = (0.1 * torch.randn(batch, 1, 6, dim, generator=g)).to(dtype).to(device)
    context = torch.randn(batch, text_len, dim, generator=g).to(dtype).to(device)
    grid_sizes = torch.tensor([list(grid)] * batch, dtype=torch.long, device=device)
    freqs = [grid_phase_table(grid, head_dim, device) for _ in range(batch)]
    return x, e, grid_sizes, freqs, context


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    os.environ.setdefault("MOJO_BACKEND", "b200")
    import mojo_opset_b200 as ops

    dim, heads, grid = 3072, 24, (4, 32, 32)  # 4096 image tokens (cfg5)
    block = DiTBlock(ops, dim=dim, num_heads=heads).eval()
    inputs = make_inputs(args.batch, grid, 512, dim, dim // heads, "cuda", torch.bfloat16)
    with torch.inference_mode():
        for _ in range(3):
            out = block(*inputs)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            out = block(*inputs)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    L = grid[0] * grid[1] * grid[2]
    print(json.dumps({"workload": f"Wan2.2-5B-shaped DiT block, batch {args.batch}, {L} image tokens, 512 text tokens",
                      "ms_per_block": ms, "image_tokens_per_s": args.batch * L / ms * 1e3,
                      "finite": bool(torch.isfinite(out.float()).all())}))


if __name__ == "__main__":
    main()
