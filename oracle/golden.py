"""TEST INFRASTRUCTURE - CPU restatement of the reference's torch-native golden ops.

Nothing in the product (``mojo_opset_b200/``) imports this module.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may, and
only as the checker / reported CPU baseline.

Each function restates one reference op with the same rounding points, written independently
(tensor-indexed page gathers instead of per-page Python loops), and cites the lines it follows.  The
arithmetic itself is PyTorch ATen (``einsum`` / ``softmax`` / ``rsqrt`` / ``silu``), i.e. the same
third-party dependency the reference calls (torch 2.11.0+cu128 in this image; the reference declares no
pin, ``pyproject.toml:13-18``).

Parity pin: ``tests/golden/*.pt`` hold input/output vectors produced by importing the UNMODIFIED
reference from ``/root/reference`` (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py``
checks every function here against them bit-for-bit (SDPA: 1e-2, the reference delegates that op to
``F.scaled_dot_product_attention`` whose CPU kernel is not restated).  The reference itself ships no
stored golden vectors (its accuracy tests are comparative), so these fixtures are the pin.
"""

import math

from typing import Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# paged KV helpers
# --------------------------------------------------------------------------------------------------
def _kv_head_of_q_head(num_q_heads: int, num_kv_heads: int, gqa_layout: str, device) -> torch.Tensor:
    """AABB: kv = h // G (repeat_interleave); ABAB: kv = h % Hkv (repeat).  attention.py:209-215."""
    h = torch.arange(num_q_heads, device=device)
    group = num_q_heads // num_kv_heads
    return h // group if gqa_layout == "AABB" else h % num_kv_heads


def _gather_sequence_kv(cache: torch.Tensor, table_row, seq_len: int, dtype) -> torch.Tensor:
    """Pages of one sequence -> ``[seq_len, Hkv, D]``.

    The reference walks the logical blocks in order and stops at the first negative id, leaving
    zeros for every later position (attention.py:192-207, :399-419).
    """
    _, num_kv_heads, block_size, head_dim = cache.shape
    need = (seq_len + block_size - 1) // block_size
    ids = list(table_row[:need])
    for j, pid in enumerate(ids):
        if pid < 0:
            ids = ids[:j]
            break
    out = torch.zeros(seq_len, num_kv_heads, head_dim, dtype=dtype, device=cache.device)
    if ids:
        pages = cache[torch.tensor(ids, dtype=torch.long, device=cache.device)]  # [n, Hkv, bs, D]
        flat = pages.permute(0, 2, 1, 3).reshape(len(ids) * block_size, num_kv_heads, head_dim)
        n = min(seq_len, flat.shape[0])
        out[:n] = flat[:n]
    return out


# --------------------------------------------------------------------------------------------------
# a1  MojoPagedDecodeGQA.forward            reference core/operators/attention.py:141-229
# --------------------------------------------------------------------------------------------------
def paged_decode_gqa(query, key_cache, value_cache, total_seq_lens, block_tables, softmax_scale=None,
                     gqa_layout: str = "AABB"):
    batch, num_q_heads, head_dim = query.shape
    num_kv_heads = key_cache.shape[1]
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(head_dim)
    head_map = _kv_head_of_q_head(num_q_heads, num_kv_heads, gqa_layout, query.device)
    out = torch.zeros(batch, num_q_heads, head_dim, dtype=query.dtype, device=query.device)
    lens = total_seq_lens.tolist()
    tables = block_tables.tolist()
    for b, seq_len in enumerate(lens):
        if seq_len <= 0:
            continue
        if tables[b][0] < 0:
            raise ValueError("Paged decode requires a valid block table for rows with kv lens > 0.")
        k = _gather_sequence_kv(key_cache, tables[b], seq_len, query.dtype)
        v = _gather_sequence_kv(value_cache, tables[b], seq_len, query.dtype)
        if num_q_heads != num_kv_heads:
            k = k[:, head_map]
            v = v[:, head_map]
        # scores in the input dtype, THEN scaled (attention.py:217); softmax fp32 -> input dtype (:227)
        scores = torch.einsum("hd,khd->hk", query[b], k) * softmax_scale
        probs = torch.softmax(scores, dim=-1, dtype=torch.float32).to(query.dtype)
        out[b] = torch.einsum("hk,khd->hd", probs, v)
    return out


# --------------------------------------------------------------------------------------------------
# a2  MojoPagedPrefillGQA.forward           reference core/operators/attention.py:335-447
# --------------------------------------------------------------------------------------------------
def paged_prefill_gqa(query, key_cache, value_cache, cu_q_lens, block_tables, softmax_scale=None,
                      cu_total_seq_lens=None, gqa_layout: str = "AABB", is_causal: bool = True):
    total_q, num_q_heads, head_dim = query.shape
    num_kv_heads = key_cache.shape[1]
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(head_dim)
    head_map = _kv_head_of_q_head(num_q_heads, num_kv_heads, gqa_layout, query.device)
    out = torch.zeros(total_q, num_q_heads, head_dim, dtype=query.dtype, device=query.device)
    cu_q = cu_q_lens.tolist()
    cu_kv = cu_q if cu_total_seq_lens is None else cu_total_seq_lens.tolist()
    tables = block_tables.tolist()
    for b in range(len(cu_q) - 1):
        lo, hi = cu_q[b], cu_q[b + 1]
        q_len = hi - lo
        kv_len = cu_kv[b + 1] - cu_kv[b]
        if q_len == 0 or kv_len <= 0:
            continue
        if tables[b][0] < 0:
            raise ValueError("Paged prefill requires a valid block table for rows with kv lens > 0.")
        k = _gather_sequence_kv(key_cache, tables[b], kv_len, query.dtype)
        v = _gather_sequence_kv(value_cache, tables[b], kv_len, query.dtype)
        if num_q_heads != num_kv_heads:
            k = k[:, head_map]
            v = v[:, head_map]
        # scores in the input dtype, upcast, THEN scaled (attention.py:432)
        scores = torch.einsum("thd,khd->thk", query[lo:hi], k).float() * softmax_scale
        if is_causal:
            # query row t sees keys 0 .. kv_len - q_len + t   (attention.py:433-437)
            key_pos = torch.arange(kv_len, device=query.device).unsqueeze(0)
            last_visible = torch.arange(q_len, device=query.device).unsqueeze(1) + (kv_len - q_len)
            scores.masked_fill_((key_pos > last_visible).unsqueeze(1), -torch.inf)
        probs = torch.softmax(scores, dim=-1, dtype=torch.float32).to(query.dtype)
        out[lo:hi] = torch.einsum("thk,khd->thd", probs, v)
    return out


# --------------------------------------------------------------------------------------------------
# f4  MojoPagedPrefillSWA.forward / MojoPagedDecodeSWA.forward   reference core/operators/attention.py:533-745
# --------------------------------------------------------------------------------------------------
def window_mask(q_len: int, kv_len: int, local_window_size=None, global_window_size=None):
    """``[q_len, kv_len]`` bool, True = visible (reference ``_generate_window_mask``, attention.py:507-531): causal
    with offset ``kv_len - q_len`` AND (inside the local window OR inside the global prefix); a window that is ``None``
    contributes nothing, both ``None`` leaves the causal mask."""
    pos = torch.arange(q_len).unsqueeze(1) + (kv_len - q_len)
    key = torch.arange(kv_len).unsqueeze(0)
    visible = key <= pos
    if local_window_size is None and global_window_size is None:
        return visible
    windows = torch.zeros(q_len, kv_len, dtype=torch.bool)
    if local_window_size is not None:
        windows |= pos <= key + local_window_size
    if global_window_size is not None:
        windows |= key < global_window_size
    return visible & windows


def paged_prefill_swa(query, key_cache, value_cache, cu_q_lens, block_table, softmax_scale=None,
                      cu_total_seq_lens=None, gqa_layout: str = "AABB", is_causal: bool = True,
                      local_window_size=None, global_window_size=None):
    """Scores by a batched matmul in the input dtype, upcast, scaled (attention.py:611); window mask; softmax pieces in
    fp32 with P rounded to the input dtype before P V (:620-624); P V in the input dtype, upcast, divided by l (:635)."""
    total_q, num_q_heads, head_dim = query.shape
    num_kv_heads = key_cache.shape[1]
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(head_dim)
    head_map = _kv_head_of_q_head(num_q_heads, num_kv_heads, gqa_layout, query.device)
    out = torch.zeros(total_q, num_q_heads, head_dim, dtype=query.dtype, device=query.device)
    cu_q = cu_q_lens.tolist()
    cu_kv = cu_q if cu_total_seq_lens is None else cu_total_seq_lens.tolist()
    tables = block_table.tolist()
    for b in range(len(cu_q) - 1):
        lo, hi = cu_q[b], cu_q[b + 1]
        q_len, kv_len = hi - lo, cu_kv[b + 1] - cu_kv[b]
        if q_len == 0 or kv_len <= 0:
            continue
        if tables[b][0] < 0:
            raise ValueError("Paged prefill requires a valid block table for rows with kv lens > 0.")
        k = _gather_sequence_kv(key_cache, tables[b], kv_len, query.dtype)[:, head_map]    # [kv, Hq, D]
        v = _gather_sequence_kv(value_cache, tables[b], kv_len, query.dtype)[:, head_map]
        q = query[lo:hi].permute(1, 0, 2)                                                  # [Hq, q, D]
        scores = torch.bmm(q, k.permute(1, 2, 0)).float() * softmax_scale                  # [Hq, q, kv]
        if is_causal:
            vis = window_mask(q_len, kv_len, local_window_size, global_window_size).to(scores.device)
            scores = torch.where(vis, scores, float("-inf"))
        scores = scores - scores.max(dim=-1, keepdim=True).values
        p = torch.exp(scores)
        l = p.sum(dim=-1, keepdim=True)
        o = torch.bmm(p.to(query.dtype), v.permute(1, 0, 2)).float() / l
        out[lo:hi] = o.permute(1, 0, 2).to(out.dtype)
    return out


def paged_decode_swa(query, key_cache, value_cache, total_seq_lens, block_table, softmax_scale=None,
                     gqa_layout: str = "AABB", local_window_size=None, global_window_size=None):
    """One query token per sequence at position ``seq_len - 1`` (attention.py:672-742): the prefill rule with
    ``q_len = 1``; rows with ``seq_len <= 0`` stay zero."""
    lens = [max(int(n), 0) for n in total_seq_lens.tolist()]
    batch = query.shape[0]
    cu_q = torch.arange(batch + 1, dtype=torch.int32)
    cu_kv = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32)
    return paged_prefill_swa(query, key_cache, value_cache, cu_q, block_table, softmax_scale, cu_kv, gqa_layout, True,
                             local_window_size, global_window_size)


def swa(query, key, value, cu_q_lens, cu_total_seq_lens, softmax_scale=None, gqa_layout: str = "AABB",
        is_causal: bool = True, local_window_size=None, global_window_size=None):
    """Non-paged ``MojoSWA.forward`` (attention.py:776-834): the paged SWA arithmetic over packed ``key/value[Tk,Hkv,D]``
    rows ``cu_total_seq_lens[b] : cu_total_seq_lens[b+1]``.  Rows outside every sequence are left as zeros here (the
    reference leaves them uninitialised)."""
    total_q, num_q_heads, head_dim = query.shape
    num_kv_heads = key.shape[1]
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(head_dim)
    head_map = _kv_head_of_q_head(num_q_heads, num_kv_heads, gqa_layout, query.device)
    out = torch.zeros_like(query)
    cu_q, cu_kv = cu_q_lens.tolist(), cu_total_seq_lens.tolist()
    for b in range(len(cu_q) - 1):
        lo, hi = cu_q[b], cu_q[b + 1]
        q_len, kv_len = hi - lo, cu_kv[b + 1] - cu_kv[b]
        if q_len == 0 or kv_len <= 0:
            continue
        k = key[cu_kv[b]:cu_kv[b + 1]][:, head_map]                                    # [kv, Hq, D]
        v = value[cu_kv[b]:cu_kv[b + 1]][:, head_map]
        q = query[lo:hi].permute(1, 0, 2)                                              # [Hq, q, D]
        scores = torch.bmm(q, k.permute(1, 2, 0)).float() * softmax_scale              # [Hq, q, kv]
        if is_causal:
            vis = window_mask(q_len, kv_len, local_window_size, global_window_size).to(scores.device)
            scores = torch.where(vis, scores, float("-inf"))
        scores = scores - scores.max(dim=-1, keepdim=True).values
        p = torch.exp(scores)
        l = p.sum(dim=-1, keepdim=True)
        o = torch.bmm(p.to(value.dtype), v.permute(1, 0, 2)).float() / l
        out[lo:hi] = o.permute(1, 0, 2).to(out.dtype)
    return out


# --------------------------------------------------------------------------------------------------
# a3  MojoSdpa.forward                      reference core/operators/attention.py:466-501
# --------------------------------------------------------------------------------------------------
def sdpa(query, key, value, attn_mask=None, scale: Optional[float] = None, enable_gqa: bool = False):
    """Published definition of ``F.scaled_dot_product_attention(..., is_causal=False)`` evaluated in fp32:
    ``softmax(Q K^T * scale + mask) V`` with ``scale = 1/sqrt(D)`` by default, a bool mask meaning
    "True = take part", and GQA mapping q head ``h`` -> kv head ``h // G``.  Rounded once to the input dtype.
    """
    head_dim = query.shape[-1]
    if scale is None:
        scale = 1.0 / math.sqrt(head_dim)
    q, k, v = query.float(), key.float(), value.float()
    if enable_gqa and q.shape[-3] != k.shape[-3]:
        group = q.shape[-3] // k.shape[-3]
        k = k.repeat_interleave(group, dim=-3)
        v = v.repeat_interleave(group, dim=-3)
    scores = torch.matmul(q, k.transpose(-1, -2)) * scale
    if attn_mask is not None:
        if attn_mask.dtype == torch.bool:
            scores = scores.masked_fill(~attn_mask, -torch.inf)
        else:
            scores = scores + attn_mask.float()
    return torch.matmul(torch.softmax(scores, dim=-1), v).to(query.dtype)


def sdpa_aten(query, key, value, attn_mask=None, scale=None, enable_gqa=False):
    """The reference's literal call (attention.py:490-499) - the same ATen entry point."""
    return F.scaled_dot_product_attention(
        query, key, value, attn_mask=attn_mask, dropout_p=0.0, is_causal=False, scale=scale, enable_gqa=enable_gqa
    )


# --------------------------------------------------------------------------------------------------
# a4  MojoStorePagedKVCache.forward + build_paged_kv_chunk_metadata   reference core/operators/kv_cache.py:33-171
# --------------------------------------------------------------------------------------------------
def build_chunk_plan(block_table, cu_q_lens, context_kv_lens, block_size: int):
    """Scalar restatement of the plan builder (kv_cache.py:45-101): plain Python loops over sequences
    and logical blocks, output rows ordered by (sequence, logical block)."""
    table = block_table.tolist()
    ctx = context_kv_lens.tolist()
    width = block_table.shape[1]
    rows = []
    if cu_q_lens is None:
        for i, c in enumerate(ctx):
            if c < 0 or width == 0:
                continue
            logical = c // block_size
            if logical >= width or table[i][logical] < 0:
                continue
            rows.append((i, table[i][logical], c % block_size, 1))
    else:
        cu = cu_q_lens.tolist()
        for i, c in enumerate(ctx):
            q_len = cu[i + 1] - cu[i]
            if q_len <= 0 or c < 0:
                continue
            for j in range(width):
                lo = max(c, j * block_size)
                hi = min(c + q_len, (j + 1) * block_size)
                if hi - lo <= 0 or table[i][j] < 0:
                    continue
                rows.append((cu[i] + lo - c, table[i][j], lo - j * block_size, hi - lo))
    return torch.tensor(rows, dtype=torch.int32, device=block_table.device).reshape(-1, 4)


def store_paged_kv(key_states, value_states, key_cache, value_cache, chunk_plan):
    """In-place scatter, chunk by chunk in plan order (kv_cache.py:161-169).  Pure copy: bit-exact."""
    for src, blk, off, n in chunk_plan.tolist():
        key_cache[blk, :, off:off + n, :] = key_states[src:src + n].transpose(0, 1)
        value_cache[blk, :, off:off + n, :] = value_states[src:src + n].transpose(0, 1)
    return key_cache, value_cache


# --------------------------------------------------------------------------------------------------
# a5 / a5'  MojoResidualAddRMSNorm, MojoRMSNorm      reference core/operators/normalization.py:71-111,308-362
# --------------------------------------------------------------------------------------------------
def rms_norm(x, weight, eps: float):
    """``F.rms_norm`` as ATen evaluates it on CPU for fp32/bf16/fp16 (probed against torch 2.11):
    upcast, ``x * rsqrt(mean(x^2) + eps) * w`` in fp32, ONE rounding to the input dtype."""
    xf = x.float()
    inv = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (xf * inv * weight.float()).to(x.dtype)


def residual_add_rms_norm(hidden_state, residual, weight, eps: float, norm_pos: str = "pre"):
    summed = hidden_state + residual  # rounded to the input dtype before the norm (normalization.py:342,350)
    y = rms_norm(summed, weight, eps)
    return (y, summed) if norm_pos == "pre" else (y, y)


# --------------------------------------------------------------------------------------------------
# a6 / a6'  MojoApplyRoPE, MojoRotaryEmbedding       reference core/operators/position_embedding.py:9-175
# --------------------------------------------------------------------------------------------------
def apply_rope(q, k, cos, sin, head_first: bool = True):
    """Rotate-half on the LAST ``cos.shape[-1]`` features; math in promote(q.dtype, cos.dtype) with every
    product and the sum individually rounded (no FMA), result cast to the q/k dtype (:115-135,169-175)."""
    axis = -3 if head_first else -2
    cos, sin = cos.unsqueeze(axis), sin.unsqueeze(axis)
    rope_dim = cos.shape[-1]

    def rotate(x):
        keep, rot = x[..., : x.shape[-1] - rope_dim], x[..., x.shape[-1] - rope_dim:]
        half = rope_dim // 2
        swapped = torch.cat((-rot[..., half:], rot[..., :half]), dim=-1)
        turned = (rot * cos + swapped * sin).to(x.dtype)
        return torch.cat((keep, turned), dim=-1) if keep.shape[-1] else turned

    return rotate(q), rotate(k)


def rotary_positions(x, cu_q_lens=None, total_seq_lens=None, position_ids=None):
    """Position ids per token (position_embedding.py:68-86)."""
    if cu_q_lens is not None:
        pos = torch.full((x.shape[0],), -1, dtype=torch.int32, device=x.device)
        cu = cu_q_lens.tolist()
        for i in range(len(cu) - 1):
            q_len = cu[i + 1] - cu[i]
            ctx = 0 if total_seq_lens is None else int(total_seq_lens[i]) - q_len
            pos[cu[i]:cu[i + 1]] = torch.arange(ctx, ctx + q_len, dtype=torch.int32, device=x.device)
        return pos
    if position_ids is not None:
        return position_ids
    return torch.arange(x.shape[1], dtype=torch.int32, device=x.device)


def rotary_cos_sin(position_ids, inv_freq, attention_scaling: float = 1.0):
    """cos/sin rows: ``angle = pos * inv_freq`` in fp32, halves duplicated (position_embedding.py:88-92)."""
    freqs = position_ids[..., None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos() * attention_scaling, emb.sin() * attention_scaling


# --------------------------------------------------------------------------------------------------
# a7 / a7'  MojoSwiGLU, MojoSilu                     reference core/operators/activation.py:20-66
# --------------------------------------------------------------------------------------------------
def silu(x):
    return F.silu(x)


def swiglu(gate_out, up_out, swiglu_limit: float = 0.0):
    """``silu(gate)`` rounded to the input dtype, then ``* up`` rounded again (activation.py:60-63)."""
    if swiglu_limit > 0:
        up_out = torch.clamp(up_out, min=-swiglu_limit, max=swiglu_limit)
        gate_out = torch.clamp(gate_out, max=swiglu_limit)
    return F.silu(gate_out) * up_out


# --------------------------------------------------------------------------------------------------
# f1  MojoGemmAllReduce                          reference core/operators/compute_with_comm.py:12-24,57-117
# --------------------------------------------------------------------------------------------------
def gemm(input, weight, bias=None, trans_weight: bool = False):
    """The rank-local projection (compute_with_comm.py:12-24): ``input @ weight (+ bias)`` or ``F.linear``."""
    if trans_weight:
        output = input @ weight
        if bias is not None:
            output = output + bias
        return output
    return F.linear(input, weight, bias)


def gemm_allreduce(input, weight, bias=None, trans_weight: bool = False, process_group=None):
    """``all_reduce_sum(gemm(...))`` over ``process_group``; identity reduction without an initialised group
    (compute_with_comm.py:111-117)."""
    import torch.distributed as dist

    output = gemm(input, weight, bias, trans_weight)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(output, op=dist.ReduceOp.SUM, group=process_group)
    return output


def gemm_allreduce_emulated(inputs, weights, biases=None, trans_weight: bool = False):
    """Single-process statement of the same result for a list of per-rank shards: every rank's projection is
    materialised in the input dtype (as ``F.linear`` does), the all-reduce then sums those tensors."""
    parts = [gemm(x, w, None if biases is None else biases[r], trans_weight)
             for r, (x, w) in enumerate(zip(inputs, weights))]
    return torch.stack([p.float() for p in parts]).sum(0).to(parts[0].dtype)


# --------------------------------------------------------------------------------------------------
# f2  MojoRoPEStoreKV / MojoNormRoPEStoreKV: the composition the reference's Qwen3 block performs with four ops
#     (modeling/qwen3/mojo_qwen3_dense.py:229-234 + PagedDummyCache.update :99-109)
# --------------------------------------------------------------------------------------------------
def norm_rope_store_kv(q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens, context_kv_lens,
                       q_weight=None, k_weight=None, eps: float = 1e-6):
    """Token-major ``[T, heads, D]`` tensors.  Returns ``(q_rot, k_rot)``; the caches are updated in place."""
    if q_weight is not None:
        q, k = rms_norm(q, q_weight, eps), rms_norm(k, k_weight, eps)
    q_rot, k_rot = apply_rope(q, k, cos, sin, head_first=False)
    plan = build_chunk_plan(block_table, cu_q_lens, context_kv_lens, key_cache.shape[2])
    store_paged_kv(k_rot, v, key_cache, value_cache, plan)
    return q_rot, k_rot


# --------------------------------------------------------------------------------------------------
# f3  DiT block ops: MojoGelu (activation.py:6-17), MojoLayerNorm (normalization.py:19-66),
#     MojoGridRoPE (experimental/operators/position_embedding.py:80-118)
# --------------------------------------------------------------------------------------------------
def gelu(x):
    return F.gelu(x)


def layer_norm(x, weight, bias, eps: float):
    return F.layer_norm(x, [x.shape[-1]], weight=weight, bias=bias, eps=eps)


def grid_rope(x, grid_sizes, freqs_list):
    """Per sample: the first F*H*W tokens as complex pairs (fp32) times the phase table, the rest unchanged; the
    result is cast back to x's dtype (position_embedding.py:106-118)."""
    n = x.size(2)
    output = []
    for i, (f, h, w) in enumerate(grid_sizes.tolist()):
        seq_len = f * h * w
        x_i = torch.view_as_complex(x[i, :seq_len].to(torch.float32).reshape(seq_len, n, -1, 2))
        x_i = torch.view_as_real(x_i * freqs_list[i]).flatten(2)
        x_i = torch.cat([x_i, x[i, seq_len:]])
        output.append(x_i)
    return torch.stack(output).type_as(x)
