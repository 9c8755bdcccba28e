"""-m gpu: the UNMODIFIED reference package driven on the B200 through the plugin (VERDICT r1 missing #3, weak #9).

``baseline/_ref`` holds a pip install of the reference (``__graft_entry__.build()`` makes it where ``/root/reference``
exists; it is git-ignored but travels to the GPU box like the built ``.so``).  In a subprocess with that directory on
``PYTHONPATH`` this test

1. ``import mojo_opset`` (the reference), ``mojo_opset_b200.plugin.register()``;
2. for every hot-path op instantiates upstream's ``Mojo<Op>`` under ``MOJO_BACKEND=b200`` and runs the reference's
   own A/B harness ``forward_diff_with`` (``core/operator.py:81-129``) against upstream's ``Torch<Op>`` on CUDA
   tensors - the reference's decode list included (``tests/accuracy/operators/test_attention.py:86-92``);
3. builds the reference's own ``Qwen3ForCausalLM`` (``modeling/qwen3/mojo_qwen3_dense.py``) twice - backend torch and
   backend b200, same weights - and steps both through ``Qwen3Attention.paged_attention_forward`` token by token
   (``PagedDummyCache.update`` -> ``MojoStorePagedKVCache`` -> ``MojoPagedDecodeGQA``), comparing logits.  (The
   reference's multi-token prefill branch passes a non-cumulative ``cu_total_seq_lens`` and trips its own contract
   assertion on every backend - SURVEY.md 3.2 - so tokens enter one at a time.)
"""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in (os.environ.get("MOJO_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "mojo_opset")):
            return cand
    return None


SCRIPT = r"""
import math, os, torch
import mojo_opset                                   # the reference, unmodified
from mojo_opset_b200 import plugin
made = plugin.register()
assert sorted(made) == sorted(plugin.OPS), sorted(made)
import mojo_opset.experimental as experimental
DEV = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16

def pair(name, *args, **kw):
    core = getattr(mojo_opset, "Mojo" + name, None) or getattr(experimental, "Mojo" + name)
    os.environ["MOJO_BACKEND"] = "b200"
    ours = core(*args, **kw)
    assert type(ours).__name__ == "B200" + name and type(ours).__base__ is core, type(ours)
    ref = core._registry.get("torch")(*args, **kw)
    assert type(ref).__name__ == "Torch" + name
    return ours, ref

def paged(B, lens, Hkv, D, bs, dtype=bf):
    need = [(n + bs - 1) // bs for n in lens]
    nb = sum(need) + 10
    kc = torch.randn(nb, Hkv, bs, D, device=DEV).to(dtype)
    vc = torch.randn(nb, Hkv, bs, D, device=DEV).to(dtype)
    table = torch.full((B, max(max(need), 1)), -1, dtype=torch.int32)
    free = torch.randperm(nb).to(torch.int32)
    pos = 0
    for i, n in enumerate(need):
        table[i, :n] = free[pos:pos + n]
        pos += n
    return kc, vc, table.to(DEV)

done = []
# ---- MojoPagedDecodeGQA: the reference's decode list (test_attention.py:86-92), both layouts
for (B, Hq, Hkv, D, max_len, bs) in [(8, 16, 4, 128, 1024, 32), (8, 16, 4, 96, 1024, 128), (8, 8, 1, 128, 8192, 1024),
                                     (8, 8, 1, 128, 2048, 1024), (8, 8, 1, 128, 0, 1024)]:
    for layout in ("ABAB", "AABB"):
        lens = (torch.randint(0, max_len, (B,), dtype=torch.int32).clamp(min=1) if max_len
                else torch.randperm(B).to(torch.int32))
        kc, vc, table = paged(B, lens.tolist(), Hkv, D, bs)
        q = torch.randn(B, Hq, D, device=DEV).to(bf)
        ours, ref = pair("PagedDecodeGQA", is_causal=True, gqa_layout=layout)
        ours.forward_diff_with(ref, q, kc, vc, lens.to(DEV), table, softmax_scale=1.0 / math.sqrt(D),
                               max_total_seq_len=int(lens.max()), atol=2e-2, rtol=2e-2)
done.append("PagedDecodeGQA")

# ---- MojoPagedPrefillGQA (ragged, cached prefixes), MojoPagedPrefillSWA / MojoPagedDecodeSWA
q_lens, prefix = [300, 1, 513], [0, 260, 77]
kv = [a + b for a, b in zip(q_lens, prefix)]
kc, vc, table = paged(3, kv, 2, 128, 16)
q = torch.randn(sum(q_lens), 8, 128, device=DEV).to(bf)
cu_q = torch.tensor([0, 300, 301, 814], dtype=torch.int32, device=DEV)
cu_kv = torch.tensor([0] + torch.tensor(kv).cumsum(0).tolist(), dtype=torch.int32, device=DEV)
ours, ref = pair("PagedPrefillGQA", is_causal=True, gqa_layout="AABB")
ours.forward_diff_with(ref, q, kc, vc, cu_q, table, softmax_scale=None, cu_total_seq_lens=cu_kv, max_q_len=513,
                       max_total_seq_len=max(kv), atol=2e-2, rtol=2e-2)
done.append("PagedPrefillGQA")
ours, ref = pair("PagedPrefillSWA", is_causal=True, gqa_layout="AABB", global_window_size=16, local_window_size=100)
ours.forward_diff_with(ref, q, kc, vc, cu_q, table, softmax_scale=None, cu_total_seq_lens=cu_kv, atol=2e-2, rtol=2e-2)
done.append("PagedPrefillSWA")
lens = torch.tensor(kv, dtype=torch.int32, device=DEV)
qd = torch.randn(3, 8, 128, device=DEV).to(bf)
ours, ref = pair("PagedDecodeSWA", is_causal=True, gqa_layout="AABB", global_window_size=16, local_window_size=100)
ours.forward_diff_with(ref, qd, kc, vc, lens, table, softmax_scale=None, atol=2e-2, rtol=2e-2)
done.append("PagedDecodeSWA")

# ---- MojoSWA (non-paged, packed var-len key / value)
kpk, vpk = (torch.randn(sum(kv), 2, 128, device=DEV).to(bf) for _ in range(2))
ours, ref = pair("SWA", is_causal=True, gqa_layout="AABB", global_window_size=16, local_window_size=100)
ours.forward_diff_with(ref, q, kpk, vpk, cu_q, cu_kv, None, atol=2e-2, rtol=2e-2)
done.append("SWA")

# ---- MojoSdpa: DiT-shaped transposed views
qs, ks, vs = (torch.randn(2, 640, 6, 128, device=DEV).to(bf).transpose(1, 2) for _ in range(3))
ours, ref = pair("Sdpa")
ours.forward_diff_with(ref, qs, ks, vs, atol=1e-2, rtol=1e-2)
amask = torch.rand(640, 640, device=DEV) > 0.3
amask[:, 0] = True
ours.forward_diff_with(ref, qs, ks, vs, amask, atol=1e-2, rtol=1e-2)      # bool attn_mask
done.append("Sdpa")

# ---- MojoStorePagedKVCache: both signatures, bit exact
kc, vc, table = paged(3, [400, 90, 333], 4, 128, 16)
ctx = torch.tensor([100, 0, 301], dtype=torch.int32, device=DEV)
cu = torch.tensor([0, 40, 90, 122], dtype=torch.int32, device=DEV)
k_new, v_new = (torch.randn(122, 4, 128, device=DEV).to(bf) for _ in range(2))
ours, ref = pair("StorePagedKVCache")
ours.forward_diff_with(ref, k_new, v_new, kc, vc, table, cu, ctx, atol=0, rtol=0)
from mojo_opset.core.operators.kv_cache import build_paged_kv_chunk_metadata
meta = build_paged_kv_chunk_metadata(table, cu, ctx, 16)
ours.forward_diff_with(ref, k_new, v_new, kc, vc, chunk_metadata=meta, atol=0, rtol=0)
ours.forward_diff_with(ref, k_new[:3], v_new[:3], kc, vc, table, None, ctx, atol=0, rtol=0)   # decode mode
done.append("StorePagedKVCache")

# ---- norms, RoPE, activations
x, res = (torch.randn(70, 4096, device=DEV).to(bf) for _ in range(2))
for pos in ("pre", "post"):
    ours, ref = pair("ResidualAddRMSNorm", 4096, eps=1e-6, norm_pos=pos, device=DEV, dtype=bf)
    with torch.no_grad():
        ours.weight.normal_(); ref.weight.copy_(ours.weight)
    ours.forward_diff_with(ref, x, res, atol=5e-2, rtol=1e-2)
done.append("ResidualAddRMSNorm")
ours, ref = pair("RMSNorm", 128, eps=1e-6, device=DEV, dtype=bf)
with torch.no_grad():
    ours.weight.normal_(); ref.weight.copy_(ours.weight)
ours.forward_diff_with(ref, torch.randn(5, 9, 8, 128, device=DEV).to(bf), atol=3e-2, rtol=6e-3)
done.append("RMSNorm")
ours, ref = pair("LayerNorm", 3072, eps=1e-6, device=DEV, dtype=bf)
with torch.no_grad():
    ours.weight.normal_(); ours.bias.normal_(); ref.weight.copy_(ours.weight); ref.bias.copy_(ours.bias)
ours.forward_diff_with(ref, torch.randn(2, 100, 3072, device=DEV).to(bf), atol=5e-2, rtol=1e-2)
done.append("LayerNorm")
ours, ref = pair("RotaryEmbedding", 1e6, 128, device=DEV)
pos_ids = torch.randint(0, 4096, (33,), device=DEV, dtype=torch.int32)
cos, sin = ours.forward_diff_with(ref, torch.empty(33, 4096, device=DEV, dtype=bf), position_ids=pos_ids,
                                  atol=1e-5, rtol=1e-5)
done.append("RotaryEmbedding")
qr = torch.randn(33, 32, 128, device=DEV).to(bf)
kr = torch.randn(33, 8, 128, device=DEV).to(bf)
ours, ref = pair("ApplyRoPE")
ours.forward_diff_with(ref, qr, kr, cos, sin, head_first=False, atol=5e-2, rtol=5e-2)
ours.forward_diff_with(ref, qr.transpose(0, 1), kr.transpose(0, 1), cos, sin, head_first=True, atol=5e-2, rtol=5e-2)
done.append("ApplyRoPE")
g, u = (torch.randn(64, 12288, device=DEV).to(bf) for _ in range(2))
ours, ref = pair("SwiGLU")
ours.forward_diff_with(ref, g, u, atol=1e-2, rtol=1e-2)
done.append("SwiGLU")
for name in ("Silu", "Gelu"):
    ours, ref = pair(name)
    ours.forward_diff_with(ref, g, atol=1e-2, rtol=1e-2)
    done.append(name)
print("AB_OK", ",".join(done))

# ---- the reference's own Qwen3 model on the b200 backend vs the torch backend
from mojo_opset.modeling.qwen3 import mojo_qwen3_dense as qw
cfg = qw.Qwen3Config()
cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads, cfg.num_key_value_heads = 512, 1024, 4, 2
cfg.head_dim, cfg.num_hidden_layers, cfg.vocab_size, cfg.max_position_embeddings = 128, 2, 1000, 256
cfg.rope_theta = 1e6
def build(backend):
    os.environ["MOJO_BACKEND"] = backend
    torch.manual_seed(1)
    model = qw.Qwen3ForCausalLM(cfg).to(DEV).to(bf)
    for name, p in model.named_parameters():
        if p.dim() == 1:  # norm weights are created with torch.empty upstream
            with torch.no_grad():
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=torch.Generator().manual_seed(len(name))).to(DEV))
    return model
m_ref, m_b200 = build("torch"), build("b200")
m_b200.load_state_dict(m_ref.state_dict())
att = m_b200.model.layers[0].self_attn
assert type(att.attn_decode).__name__ == "B200PagedDecodeGQA" and type(att.q_norm).__name__ == "B200RMSNorm"
assert type(m_ref.model.layers[0].self_attn.attn_decode).__name__ == "TorchPagedDecodeGQA"
B, steps = 3, 20   # crosses a page boundary (block_size 16)
tokens = torch.randint(0, cfg.vocab_size, (B, steps), device=DEV)
cache_r = cache_b = None
worst = 0.0
with torch.no_grad():
    for t in range(steps):
        os.environ["MOJO_BACKEND"] = "torch"
        lr, cache_r = m_ref(tokens[:, t:t + 1], past_key_values=cache_r)
        os.environ["MOJO_BACKEND"] = "b200"
        lb, cache_b = m_b200(tokens[:, t:t + 1], past_key_values=cache_b)
        assert type(cache_b.store_paged_kv).__name__ == "B200StorePagedKVCache"
        torch.testing.assert_close(lb.float(), lr.float(), atol=6e-2, rtol=6e-2)
        worst = max(worst, (lb.float() - lr.float()).abs().max().item())
assert torch.equal(cache_b.block_tables, cache_r.block_tables) and torch.equal(cache_b.seq_lens, cache_r.seq_lens)
print("QWEN3_OK worst |dlogit| %.4f" % worst)
"""


def test_reference_ab_harness_and_qwen3_on_b200():
    ref = _reference_root()
    if ref is None:
        pytest.skip("no reference install (baseline/_ref): run __graft_entry__.build() where /root/reference exists")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ref, ROOT]), PYTHONDONTWRITEBYTECODE="1",
               MOJO_OPSET_PLUGIN_AUTOLOAD="0")
    env.pop("MOJO_DISABLE_ASSERTION_REWRITE", None)  # upstream's rewrite_assertion() breaks when it is set
    env.pop("MOJO_BACKEND", None)
    res = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=900)
    tail = res.stdout[-3000:] + res.stderr[-3000:]
    assert res.returncode == 0, tail
    assert "AB_OK" in res.stdout and "QWEN3_OK" in res.stdout, tail
