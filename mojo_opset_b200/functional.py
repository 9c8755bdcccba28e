"""Tensor-level entry points of the b200 backend: marshal torch tensors into the C ABI.

Every function enqueues CUDA work on ``torch.cuda.current_stream()`` and returns without synchronising or
reading device data on the host (CUDA-graph capturable).  torch is used for output / workspace allocation
and stream handles only.  Unsupported-but-valid requests raise ``NotImplementedError``; malformed ones
``ValueError``/``AssertionError``; nothing falls back to eager torch.
"""

import functools
import math

from typing import Optional
from typing import Tuple

import torch

from . import _lib


def _require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "mojo_opset_b200 kernels run on an sm_100 GPU only: got a tensor on "
                f"'{t.device}' (there is no CPU fallback)"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors live on different devices: {dev} vs {t.device}")
    return dev


def _on_tensor_device(fn):
    """Run ``fn`` with the CUDA device of its first CUDA tensor argument current.  Kernels launch on the process's
    current device while streams and pointers belong to the tensors' device: a model placed on ``cuda:1`` by HF
    ``device_map`` / accelerate while ``cuda:0`` is current would otherwise fail at launch (invalid resource handle).
    The switch costs nothing when the device already is current."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)

    return wrapper


def _inner_contiguous(t: torch.Tensor) -> torch.Tensor:
    return t if t.stride(-1) == 1 or t.shape[-1] == 1 else t.contiguous()


def _as_rows(t: torch.Tensor) -> torch.Tensor:
    """View ``[..., H]`` as ``[rows, H]`` with a single row stride (copy only if it cannot be a view)."""
    hidden = t.shape[-1]
    if t.dim() == 2 and t.stride(-1) == 1:
        return t
    if t.stride(-1) == 1 or hidden == 1:
        try:
            return t.view(-1, hidden)
        except RuntimeError:
            pass
    return t.contiguous().view(-1, hidden)


# ------------------------------------------------------------------------------------------------------
# MojoStorePagedKVCache
# ------------------------------------------------------------------------------------------------------
@_on_tensor_device
def store_paged_kv(
    key_states: torch.Tensor,
    value_states: torch.Tensor,
    key_cache: torch.Tensor,
    value_cache: torch.Tensor,
    *,
    chunk_metadata: Optional[torch.Tensor] = None,
    block_table: Optional[torch.Tensor] = None,
    cu_q_lens: Optional[torch.Tensor] = None,
    context_kv_lens: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """In-place scatter of new K/V tokens into the paged caches (bit exact).  Returns the caches."""
    dev = _require_cuda(key_states, value_states, key_cache, value_cache, chunk_metadata, block_table, cu_q_lens,
                        context_kv_lens)
    lib = _lib.load()
    if not (key_states.dtype == value_states.dtype == key_cache.dtype == value_cache.dtype):
        raise ValueError("store_paged_kv: states and caches must share one dtype")
    if key_cache.shape != value_cache.shape or key_cache.dim() != 4:
        raise ValueError("store_paged_kv: caches must be [num_blocks, kv_heads, block_size, head_dim]")
    num_blocks, num_kv_heads, block_size, head_dim = key_cache.shape
    if key_states.shape[1:] != (num_kv_heads, head_dim):
        raise ValueError(f"store_paged_kv: states {tuple(key_states.shape)} do not match cache heads/dim")
    if key_cache.stride(-1) != 1 or value_cache.stride(-1) != 1:
        raise NotImplementedError("store_paged_kv: cache head_dim must be contiguous (in-place op cannot copy)")
    ks, vs = _inner_contiguous(key_states), _inner_contiguous(value_states)
    tokens = ks.shape[0]
    common = (
        tokens, num_kv_heads, head_dim, num_blocks, block_size,
        ks.stride(0), ks.stride(1), vs.stride(0), vs.stride(1),
        key_cache.stride(0), key_cache.stride(1), key_cache.stride(2),
        value_cache.stride(0), value_cache.stride(1), value_cache.stride(2),
        _lib.dtype_id(ks.dtype), _lib.stream_ptr(dev),
    )
    if chunk_metadata is not None:
        plan = chunk_metadata.contiguous()
        rc = lib.mojo_b200_store_paged_kv_chunks(
            ks.data_ptr(), vs.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(), plan.data_ptr(),
            plan.shape[0], *common)
    else:
        table = block_table if block_table.stride(-1) == 1 else block_table.contiguous()
        ctx = context_kv_lens.contiguous()
        cu = None if cu_q_lens is None else cu_q_lens.contiguous()
        rc = lib.mojo_b200_store_paged_kv_table(
            ks.data_ptr(), vs.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(), table.data_ptr(),
            table.stride(0), table.shape[1], _lib.ptr(cu), ctx.data_ptr(), ctx.shape[0], *common)
    _lib.check(lib, rc, "store_paged_kv")
    return key_cache, value_cache


# ------------------------------------------------------------------------------------------------------
# MojoRMSNorm / MojoResidualAddRMSNorm
# ------------------------------------------------------------------------------------------------------
def _check_norm_weight(x, weight):
    if weight.dtype != x.dtype:
        raise NotImplementedError(f"rms_norm: weight dtype {weight.dtype} must equal input dtype {x.dtype}")
    if weight.dim() != 1 or weight.shape[0] != x.shape[-1]:
        raise ValueError(f"rms_norm: weight shape {tuple(weight.shape)} does not match hidden size {x.shape[-1]}")


@_on_tensor_device
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    dev = _require_cuda(x, weight)
    lib = _lib.load()
    _check_norm_weight(x, weight)
    xr = _as_rows(x)
    y = torch.empty(x.shape, dtype=x.dtype, device=dev)
    yr = y.view(-1, x.shape[-1])
    w = weight.detach().contiguous()
    rc = lib.mojo_b200_rms_norm(xr.data_ptr(), w.data_ptr(), yr.data_ptr(), xr.shape[0], xr.shape[1], xr.stride(0),
                                yr.stride(0), float(eps), _lib.dtype_id(x.dtype), _lib.stream_ptr(dev))
    _lib.check(lib, rc, "rms_norm")
    return y


@_on_tensor_device
def residual_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float,
                          want_sum: bool = True):
    """Returns ``(y, x + residual)``; the second item is ``None`` when ``want_sum`` is False."""
    dev = _require_cuda(x, residual, weight)
    lib = _lib.load()
    _check_norm_weight(x, weight)
    if residual.shape != x.shape or residual.dtype != x.dtype:
        raise NotImplementedError("residual_add_rms_norm: hidden_state and residual must share shape and dtype")
    xr, rr = _as_rows(x), _as_rows(residual)
    hidden = x.shape[-1]
    y = torch.empty(x.shape, dtype=x.dtype, device=dev)
    s = torch.empty(x.shape, dtype=x.dtype, device=dev) if want_sum else None
    w = weight.detach().contiguous()
    rc = lib.mojo_b200_residual_add_rms_norm(
        xr.data_ptr(), rr.data_ptr(), w.data_ptr(), y.data_ptr(), _lib.ptr(s), xr.shape[0], hidden, xr.stride(0),
        rr.stride(0), hidden, hidden, float(eps), _lib.dtype_id(x.dtype), _lib.stream_ptr(dev))
    _lib.check(lib, rc, "residual_add_rms_norm")
    return y, s


# ------------------------------------------------------------------------------------------------------
# MojoApplyRoPE / MojoRotaryEmbedding
# ------------------------------------------------------------------------------------------------------
def _bsh_strides(t: torch.Tensor, head_first: bool):
    """(batch, seq, heads, stride_b, stride_s, stride_h) of a 3-D / 4-D q or k tensor."""
    if t.dim() == 3:
        if head_first:  # [N, T, D]
            return 1, t.shape[1], t.shape[0], 0, t.stride(1), t.stride(0)
        return 1, t.shape[0], t.shape[1], 0, t.stride(0), t.stride(1)  # [T, N, D]
    if head_first:  # [B, N, S, D]
        return t.shape[0], t.shape[2], t.shape[1], t.stride(0), t.stride(2), t.stride(1)
    return t.shape[0], t.shape[1], t.shape[2], t.stride(0), t.stride(1), t.stride(2)  # [B, S, N, D]


@_on_tensor_device
def apply_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, head_first: bool = True):
    dev = _require_cuda(q, k, cos, sin)
    lib = _lib.load()
    if q.dtype != k.dtype:
        raise NotImplementedError("apply_rope: q and k must share a dtype")
    if q.shape[-1] != k.shape[-1]:
        raise ValueError("apply_rope: q and k head_dim differ")
    q, k = _inner_contiguous(q), _inner_contiguous(k)
    if cos.stride(-1) != 1 or sin.stride() != cos.stride() or sin.dtype != cos.dtype:
        cos, sin = cos.contiguous(), sin.contiguous().to(cos.dtype)
    head_dim, rope_dim = q.shape[-1], cos.shape[-1]
    if rope_dim > head_dim or rope_dim % 2:
        raise ValueError(f"apply_rope: rope_dim {rope_dim} must be even and <= head_dim {head_dim}")
    qb, qs, qh, q_sb, q_ss, q_sh = _bsh_strides(q, head_first)
    kb, ks_, kh, k_sb, k_ss, k_sh = _bsh_strides(k, head_first)
    if (qb, qs) != (kb, ks_):
        raise ValueError("apply_rope: q and k must agree on batch and sequence sizes")
    if cos.dim() == 2:
        if cos.shape[0] != qs:
            raise ValueError(f"apply_rope: cos has {cos.shape[0]} rows, q has {qs} positions")
        cos_sb, cos_ss = 0, cos.stride(0)
    elif cos.dim() == 3:
        if cos.shape[1] != qs or cos.shape[0] not in (1, qb):
            raise ValueError(f"apply_rope: cos shape {tuple(cos.shape)} does not broadcast to q {tuple(q.shape)}")
        cos_sb, cos_ss = (0 if cos.shape[0] == 1 else cos.stride(0)), cos.stride(1)
    else:
        raise ValueError("apply_rope: cos/sin must be [T, d], [S, d] or [B, S, d]")
    q_out, k_out = torch.empty_like(q), torch.empty_like(k)
    _, _, _, qo_sb, qo_ss, qo_sh = _bsh_strides(q_out, head_first)
    _, _, _, ko_sb, ko_ss, ko_sh = _bsh_strides(k_out, head_first)
    rc = lib.mojo_b200_apply_rope(
        q.data_ptr(), k.data_ptr(), cos.data_ptr(), sin.data_ptr(), q_out.data_ptr(), k_out.data_ptr(),
        qb, qs, qh, kh, head_dim, rope_dim,
        q_sb, q_ss, q_sh, k_sb, k_ss, k_sh, qo_sb, qo_ss, qo_sh, ko_sb, ko_ss, ko_sh, cos_sb, cos_ss,
        _lib.dtype_id(q.dtype), _lib.dtype_id(cos.dtype), _lib.stream_ptr(dev))
    _lib.check(lib, rc, "apply_rope")
    return q_out, k_out


def _rotary_on_device(fn):
    @functools.wraps(fn)
    def wrapper(num_tokens, out_shape, inv_freq, *args, **kwargs):
        if inv_freq.is_cuda and inv_freq.device.index != torch.cuda.current_device():
            with torch.cuda.device(inv_freq.device):
                return fn(num_tokens, out_shape, inv_freq, *args, **kwargs)
        return fn(num_tokens, out_shape, inv_freq, *args, **kwargs)

    return wrapper


@_rotary_on_device
def rotary_cos_sin(
    num_tokens: int,
    out_shape,
    inv_freq: torch.Tensor,
    attention_scaling: float = 1.0,
    *,
    position_ids: Optional[torch.Tensor] = None,
    cu_q_lens: Optional[torch.Tensor] = None,
    total_seq_lens: Optional[torch.Tensor] = None,
    period: int = 1,
    table_cos: Optional[torch.Tensor] = None,
    table_sin: Optional[torch.Tensor] = None,
):
    """cos/sin of shape ``out_shape + (rope_dim,)`` (fp32) for ``num_tokens`` positions."""
    dev = _require_cuda(inv_freq, position_ids, cu_q_lens, total_seq_lens, table_cos, table_sin)
    lib = _lib.load()
    rope_dim = 2 * inv_freq.shape[0]
    if inv_freq.dtype != torch.float32:  # model.to(bf16) / .half() cast the module's buffers: compute from an fp32 copy
        inv_freq = inv_freq.float()
    cos = torch.empty(*out_shape, rope_dim, dtype=torch.float32, device=dev)
    sin = torch.empty_like(cos)
    pos = None if position_ids is None else position_ids.contiguous()
    cu = None if cu_q_lens is None else cu_q_lens.contiguous()
    tot = None if total_seq_lens is None else total_seq_lens.contiguous()
    if table_cos is not None:
        table_cos, table_sin = table_cos.float().contiguous(), table_sin.float().contiguous()
    rc = lib.mojo_b200_rotary_cos_sin(
        cos.data_ptr(), sin.data_ptr(), num_tokens, rope_dim, inv_freq.contiguous().data_ptr(),
        float(attention_scaling), _lib.ptr(pos), _lib.ptr(cu), _lib.ptr(tot), 0 if cu is None else cu.shape[0] - 1,
        int(period), _lib.ptr(table_cos), _lib.ptr(table_sin), 0 if table_cos is None else table_cos.shape[0],
        _lib.stream_ptr(dev))
    _lib.check(lib, rc, "rotary_cos_sin")
    return cos, sin


# ------------------------------------------------------------------------------------------------------
# MojoSwiGLU / MojoSilu
# ------------------------------------------------------------------------------------------------------
def _rows_cols(t: torch.Tensor):
    """(tensor, rows, cols, row_stride): flat when contiguous, else [rows, last_dim] rows."""
    if t.is_contiguous():
        return t, 1, t.numel(), t.numel()
    r = _as_rows(t)
    return r, r.shape[0], r.shape[1], r.stride(0)


@_on_tensor_device
def swiglu(gate: torch.Tensor, up: torch.Tensor, swiglu_limit: float = 0.0) -> torch.Tensor:
    dev = _require_cuda(gate, up)
    lib = _lib.load()
    if gate.shape != up.shape or gate.dtype != up.dtype:
        raise NotImplementedError("swiglu: gate and up must share shape and dtype")
    out = torch.empty(gate.shape, dtype=gate.dtype, device=dev)
    if gate.numel() == 0:
        return out
    if gate.is_contiguous() and up.is_contiguous():
        g, u, rows, cols, g_rs, u_rs, o_rs = gate, up, 1, gate.numel(), 0, 0, 0
    else:
        g, u = _as_rows(gate), _as_rows(up)
        rows, cols, g_rs, u_rs, o_rs = g.shape[0], g.shape[1], g.stride(0), u.stride(0), g.shape[1]
    rc = lib.mojo_b200_swiglu(g.data_ptr(), u.data_ptr(), out.data_ptr(), rows, cols, g_rs, u_rs, o_rs,
                              float(swiglu_limit), _lib.dtype_id(gate.dtype), _lib.stream_ptr(dev))
    _lib.check(lib, rc, "swiglu")
    return out


@_on_tensor_device
def silu(x: torch.Tensor) -> torch.Tensor:
    dev = _require_cuda(x)
    lib = _lib.load()
    out = torch.empty(x.shape, dtype=x.dtype, device=dev)
    if x.numel() == 0:
        return out
    xr, rows, cols, x_rs = _rows_cols(x)
    rc = lib.mojo_b200_silu(xr.data_ptr(), out.data_ptr(), rows, cols, x_rs, cols, _lib.dtype_id(x.dtype),
                            _lib.stream_ptr(dev))
    _lib.check(lib, rc, "silu")
    return out


# ------------------------------------------------------------------------------------------------------
# MojoPagedDecodeGQA
# ------------------------------------------------------------------------------------------------------
def _check_paged_caches(query, key_cache, value_cache):
    if key_cache.dim() != 4 or key_cache.shape != value_cache.shape:
        raise ValueError("paged attention: caches must be [num_blocks, kv_heads, block_size, head_dim] and equal")
    if not (query.dtype == key_cache.dtype == value_cache.dtype):
        raise NotImplementedError("paged attention: query and caches must share one dtype")
    if key_cache.shape[-1] != query.shape[-1]:
        raise ValueError("paged attention: head_dim of query and cache differ")
    if query.shape[-2] % key_cache.shape[1]:
        raise ValueError("paged attention: num_q_heads must be a multiple of num_kv_heads")


@_on_tensor_device
def paged_decode_gqa(
    query: torch.Tensor,
    key_cache: torch.Tensor,
    value_cache: torch.Tensor,
    total_seq_lens: torch.Tensor,
    block_tables: torch.Tensor,
    softmax_scale: Optional[float] = None,
    gqa_layout: str = "AABB",
    max_total_seq_len: Optional[int] = None,
    num_splits: Optional[int] = None,
    local_window_size: Optional[int] = None,
    global_window_size: Optional[int] = None,
) -> torch.Tensor:
    """With a window (``MojoPagedDecodeSWA``) only the KV tiles of the global prefix and of the local window are read;
    raises ``NotImplementedError`` for shapes the tensor-tile kernel does not cover (``paged_decode_swa`` falls back)."""
    dev = _require_cuda(query, key_cache, value_cache, total_seq_lens, block_tables)
    lib = _lib.load()
    if query.dim() != 3:
        raise ValueError("paged_decode_gqa: query must be [batch, num_q_heads, head_dim]")
    _check_paged_caches(query, key_cache, value_cache)
    batch, num_q_heads, head_dim = query.shape
    num_blocks, num_kv_heads, block_size, _ = key_cache.shape
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(head_dim)
    q = _inner_contiguous(query)
    kc, vc = _inner_contiguous(key_cache), _inner_contiguous(value_cache)
    tables = block_tables if block_tables.stride(-1) == 1 or block_tables.shape[1] <= 1 else block_tables.contiguous()
    lens = total_seq_lens.contiguous()
    _lib.error_word(dev)  # unmapped first blocks are flagged on the device (mojo_opset_b200.check_device_errors)
    out = torch.empty((batch, num_q_heads, head_dim), dtype=query.dtype, device=dev)
    if batch == 0:
        return out
    dt = _lib.dtype_id(query.dtype)
    max_blocks = tables.shape[1]
    hint = max_blocks * block_size if max_total_seq_len is None else min(int(max_total_seq_len), max_blocks * block_size)
    windowed = local_window_size is not None or global_window_size is not None
    split_hint = hint
    if windowed:  # the splits divide the visible tiles (global prefix + local window + their edge tiles)
        for name, w in (("local_window_size", local_window_size), ("global_window_size", global_window_size)):
            if w is not None and int(w) < 0:
                raise ValueError(f"paged_decode_swa: {name} must be >= 0 or None")
        visible = (0 if local_window_size is None else int(local_window_size) + 1) + int(global_window_size or 0) + 128
        split_hint = max(1, min(hint, visible))
    if num_splits is None or num_splits <= 0:
        num_splits = lib.mojo_b200_paged_decode_num_splits(batch, num_q_heads, num_kv_heads, head_dim, block_size,
                                                           split_hint, dt)
    ws_bytes = lib.mojo_b200_paged_decode_workspace_bytes(batch, num_q_heads, head_dim, num_splits)
    workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    common = (
        q.data_ptr(), kc.data_ptr(), vc.data_ptr(), lens.data_ptr(), tables.data_ptr(), out.data_ptr(),
        _lib.ptr(workspace), ws_bytes, batch, num_q_heads, num_kv_heads, head_dim, num_blocks, block_size,
        max_blocks, tables.stride(0) if max_blocks else 0, hint,
        q.stride(0), q.stride(1), out.stride(0), out.stride(1),
        kc.stride(0), kc.stride(1), kc.stride(2), vc.stride(0), vc.stride(1), vc.stride(2),
        float(softmax_scale), 1 if gqa_layout == "ABAB" else 0, num_splits)
    if windowed:
        rc = lib.mojo_b200_paged_decode_swa(
            *common, -1 if local_window_size is None else int(local_window_size),
            -1 if global_window_size is None else int(global_window_size), dt, _lib.stream_ptr(dev))
        _lib.check(lib, rc, "paged_decode_swa")
    else:
        rc = lib.mojo_b200_paged_decode_gqa(*common, dt, _lib.stream_ptr(dev))
        _lib.check(lib, rc, "paged_decode_gqa")
    return out


# ------------------------------------------------------------------------------------------------------
# MojoPagedPrefillGQA / MojoSdpa
# ------------------------------------------------------------------------------------------------------
@_on_tensor_device
def paged_prefill_gqa(
    query: torch.Tensor,
    key_cache: torch.Tensor,
    value_cache: torch.Tensor,
    cu_q_lens: torch.Tensor,
    block_tables: torch.Tensor,
    softmax_scale: Optional[float] = None,
    cu_total_seq_lens: Optional[torch.Tensor] = None,
    gqa_layout: str = "AABB",
    max_q_len: Optional[int] = None,
    max_total_seq_len: Optional[int] = None,
    is_causal: bool = True,
    local_window_size: Optional[int] = None,
    global_window_size: Optional[int] = None,
) -> torch.Tensor:
    """``local_window_size`` / ``global_window_size`` (``MojoPagedPrefillSWA``): on top of the causal limit a key is
    visible iff ``key + local >= position`` or ``key < global``; both ``None`` = plain causal attention."""
    dev = _require_cuda(query, key_cache, value_cache, cu_q_lens, block_tables, cu_total_seq_lens)
    lib = _lib.load()
    if query.dim() != 3:
        raise ValueError("paged_prefill_gqa: query must be [total_q_tokens, num_q_heads, head_dim]")
    _check_paged_caches(query, key_cache, value_cache)
    total_q, num_q_heads, head_dim = query.shape
    num_blocks, num_kv_heads, block_size, _ = key_cache.shape
    batch = cu_q_lens.shape[0] - 1
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(head_dim)
    q = _inner_contiguous(query)
    kc, vc = _inner_contiguous(key_cache), _inner_contiguous(value_cache)
    tables = block_tables if block_tables.stride(-1) == 1 or block_tables.shape[1] <= 1 else block_tables.contiguous()
    cu_q = cu_q_lens.contiguous()
    cu_kv = None if cu_total_seq_lens is None else cu_total_seq_lens.contiguous()
    _lib.error_word(dev)  # unmapped first blocks are flagged on the device (mojo_opset_b200.check_device_errors)
    # rows no query block covers (tokens past cu_q_lens[-1], sequences without keys) are zero-filled by the
    # library itself, as in the golden: no host-side memset of the whole output
    out = torch.empty((total_q, num_q_heads, head_dim), dtype=query.dtype, device=dev)
    if total_q == 0:
        return out
    max_blocks = tables.shape[1]
    q_hint = total_q if max_q_len is None else min(int(max_q_len), total_q)
    kv_cap = max_blocks * block_size
    kv_hint = kv_cap if max_total_seq_len is None else min(int(max_total_seq_len), kv_cap)
    common = (
        q.data_ptr(), kc.data_ptr(), vc.data_ptr(), cu_q.data_ptr(), _lib.ptr(cu_kv), tables.data_ptr(),
        out.data_ptr(), total_q, batch, num_q_heads, num_kv_heads, head_dim, num_blocks, block_size, max_blocks,
        tables.stride(0) if max_blocks else 0, q_hint, kv_hint,
        q.stride(0), q.stride(1), out.stride(0), out.stride(1),
        kc.stride(0), kc.stride(1), kc.stride(2), vc.stride(0), vc.stride(1), vc.stride(2),
        float(softmax_scale), 1 if gqa_layout == "ABAB" else 0, 1 if is_causal else 0)
    if local_window_size is None and global_window_size is None:
        rc = lib.mojo_b200_paged_prefill_gqa(*common, _lib.dtype_id(query.dtype), _lib.stream_ptr(dev))
        _lib.check(lib, rc, "paged_prefill_gqa")
    else:
        for name, w in (("local_window_size", local_window_size), ("global_window_size", global_window_size)):
            if w is not None and int(w) < 0:
                raise ValueError(f"paged_prefill_swa: {name} must be >= 0 or None")
        rc = lib.mojo_b200_paged_prefill_swa(
            *common, -1 if local_window_size is None else int(local_window_size),
            -1 if global_window_size is None else int(global_window_size), _lib.dtype_id(query.dtype),
            _lib.stream_ptr(dev))
        _lib.check(lib, rc, "paged_prefill_swa")
    return out


@_on_tensor_device
def paged_decode_swa(
    query: torch.Tensor,
    key_cache: torch.Tensor,
    value_cache: torch.Tensor,
    total_seq_lens: torch.Tensor,
    block_tables: torch.Tensor,
    softmax_scale: Optional[float] = None,
    gqa_layout: str = "AABB",
    max_total_seq_len: Optional[int] = None,
    local_window_size: Optional[int] = None,
    global_window_size: Optional[int] = None,
) -> torch.Tensor:
    """``MojoPagedDecodeSWA``: one query token per sequence = the windowed prefill with ``q_len = 1`` rows
    (``cu_q_lens = 0..B``, cumulative ``total_seq_lens``; both built on the device, no host read) for shapes outside the
    tensor-tile decode kernel; inside it, the split-KV streaming kernel reads only the visible KV tiles.  Without any
    window this is ``paged_decode_gqa``."""
    if local_window_size is None and global_window_size is None:
        return paged_decode_gqa(query, key_cache, value_cache, total_seq_lens, block_tables, softmax_scale, gqa_layout,
                                max_total_seq_len)
    try:  # the split-KV decode kernel restricted to the visible KV tiles (bf16/fp16, head_dim 64/128)
        return paged_decode_gqa(query, key_cache, value_cache, total_seq_lens, block_tables, softmax_scale, gqa_layout,
                                max_total_seq_len, None, local_window_size, global_window_size)
    except NotImplementedError:
        pass  # other shapes: the general attention kernel with one query row per sequence (still CUDA, no fallback)
    dev = _require_cuda(query, key_cache, value_cache, total_seq_lens, block_tables)
    if query.dim() != 3:
        raise ValueError("paged_decode_swa: query must be [batch, num_q_heads, head_dim]")
    batch = query.shape[0]
    cu_q = torch.arange(batch + 1, dtype=torch.int32, device=dev)
    cu_kv = torch.zeros(batch + 1, dtype=torch.int32, device=dev)
    torch.cumsum(total_seq_lens.clamp_min(0), 0, out=cu_kv[1:])
    return paged_prefill_gqa(query, key_cache, value_cache, cu_q, block_tables, softmax_scale, cu_kv, gqa_layout, 1,
                             max_total_seq_len, True, local_window_size, global_window_size)


@_on_tensor_device
def swa(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, cu_q_lens: torch.Tensor,
        cu_total_seq_lens: torch.Tensor, softmax_scale: Optional[float] = None, gqa_layout: str = "AABB",
        local_window_size: Optional[int] = None, global_window_size: Optional[int] = None,
        max_q_len: Optional[int] = None, max_total_seq_len: Optional[int] = None) -> torch.Tensor:
    """``MojoSWA`` (non-paged): causal sliding-window attention over packed ``query[Tq,Hq,D]`` / ``key, value[Tk,Hkv,D]``
    with ``cu_q_lens`` / ``cu_total_seq_lens`` marking the sequences; the paged prefill kernels read the packed key /
    value rows directly (no copy into pages).  Rows outside every sequence read as zeros."""
    dev = _require_cuda(query, key, value, cu_q_lens, cu_total_seq_lens)
    lib = _lib.load()
    if query.dim() != 3 or key.dim() != 3 or key.shape != value.shape or key.shape[-1] != query.shape[-1]:
        raise ValueError("swa: query [Tq, Hq, D], key / value [Tk, Hkv, D] expected")
    if not (query.dtype == key.dtype == value.dtype):
        raise NotImplementedError("swa: query, key and value must share one dtype")
    total_q, num_q_heads, head_dim = query.shape
    total_kv, num_kv_heads, _ = key.shape
    if num_q_heads % num_kv_heads:
        raise ValueError("swa: num_q_heads must be a multiple of num_kv_heads")
    for name, w in (("local_window_size", local_window_size), ("global_window_size", global_window_size)):
        if w is not None and int(w) < 0:
            raise ValueError(f"swa: {name} must be >= 0 or None")
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(head_dim)
    q, k, v = _inner_contiguous(query), _inner_contiguous(key), _inner_contiguous(value)
    out = torch.empty((total_q, num_q_heads, head_dim), dtype=query.dtype, device=dev)
    if total_q == 0:
        return out
    if total_kv == 0:
        return out.zero_()
    batch = cu_q_lens.shape[0] - 1
    q_hint = total_q if max_q_len is None else min(int(max_q_len), total_q)
    kv_hint = total_kv if max_total_seq_len is None else min(int(max_total_seq_len), total_kv)
    rc = lib.mojo_b200_swa(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), cu_q_lens.contiguous().data_ptr(),
        cu_total_seq_lens.contiguous().data_ptr(), out.data_ptr(), total_q, total_kv, batch, num_q_heads, num_kv_heads,
        head_dim, q_hint, kv_hint, q.stride(0), q.stride(1), out.stride(0), out.stride(1), k.stride(0), k.stride(1),
        v.stride(0), v.stride(1), float(softmax_scale), 1 if gqa_layout == "ABAB" else 0, 1,
        -1 if local_window_size is None else int(local_window_size),
        -1 if global_window_size is None else int(global_window_size), _lib.dtype_id(query.dtype), _lib.stream_ptr(dev))
    _lib.check(lib, rc, "swa")
    return out


def _bool_mask_strides(attn_mask: torch.Tensor, batch: int, heads: int, q_len: int, kv_len: int):
    """A bool ``attn_mask`` broadcastable to ``[B, Hq, Sq, Skv]`` as (uint8 view, byte strides b / h / q) with 0 for
    broadcast dimensions; the key dimension must be dense (copied if it is not)."""
    m = attn_mask
    if m.dim() > 4:
        raise ValueError(f"sdpa: attn_mask of rank {m.dim()} does not broadcast to [B, H, Sq, Skv]")
    while m.dim() < 4:
        m = m.unsqueeze(0)
    want = (batch, heads, q_len, kv_len)
    for have, full in zip(m.shape, want):
        if have not in (1, full):
            raise ValueError(f"sdpa: attn_mask shape {tuple(attn_mask.shape)} does not broadcast to {want}")
    if m.shape[-1] != kv_len or (kv_len > 1 and m.stride(-1) != 1):
        m = m.expand(m.shape[0], m.shape[1], m.shape[2], kv_len).contiguous()
    strides = [0 if m.shape[i] == 1 else m.stride(i) for i in range(3)]
    return m.view(torch.uint8), strides


@_on_tensor_device
def sdpa(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, scale: Optional[float] = None,
         enable_gqa: bool = False, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Non-causal dense attention.  ``[B,H,S,D]`` inputs may be transposed views of ``[B,S,H,D]`` memory (the
    DiT call site); the output is written as ``[B,S,H,D]`` memory and returned as the ``[B,H,S,D]`` view, which
    is what the golden returns for such inputs and makes the caller's ``.transpose(1,2).contiguous()`` free.
    ``attn_mask``: bool, broadcastable to ``[B, Hq, Sq, Skv]``, True = the key takes part (reference
    ``attention.py:490-499``); additive float masks are not built."""
    dev = _require_cuda(query, key, value, attn_mask)
    lib = _lib.load()
    if query.dim() != 4 or key.dim() != 4 or value.dim() != 4:
        raise NotImplementedError("sdpa: only [batch, heads, seq, head_dim] inputs are supported")
    if not (query.dtype == key.dtype == value.dtype):
        raise NotImplementedError("sdpa: query, key and value must share one dtype")
    batch, num_q_heads, q_len, head_dim = query.shape
    _, num_kv_heads, kv_len, _ = key.shape
    if key.shape != value.shape or key.shape[0] != batch or key.shape[-1] != head_dim:
        raise ValueError("sdpa: key/value shapes do not match query")
    if num_q_heads != num_kv_heads and not (enable_gqa and num_q_heads % num_kv_heads == 0):
        raise ValueError("sdpa: head counts differ; pass enable_gqa=True with Hq a multiple of Hkv")
    if attn_mask is not None and attn_mask.dtype != torch.bool:
        raise NotImplementedError("sdpa: only bool attn_mask is built (additive float masks are not)")
    if scale is None:
        scale = 1.0 / math.sqrt(head_dim)
    q, k, v = _inner_contiguous(query), _inner_contiguous(key), _inner_contiguous(value)
    out = torch.empty((batch, q_len, num_q_heads, head_dim), dtype=query.dtype, device=dev).transpose(1, 2)
    if out.numel() == 0:
        return out
    if kv_len == 0:
        return out.zero_()
    common = (
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), batch, num_q_heads, num_kv_heads, q_len, kv_len,
        head_dim, q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2),
        v.stride(0), v.stride(1), v.stride(2), out.stride(0), out.stride(1), out.stride(2), float(scale))
    if attn_mask is None:
        rc = lib.mojo_b200_sdpa(*common, _lib.dtype_id(query.dtype), _lib.stream_ptr(dev))
    else:
        mask, (m_sb, m_sh, m_sq) = _bool_mask_strides(attn_mask, batch, num_q_heads, q_len, kv_len)
        rc = lib.mojo_b200_sdpa_masked(*common, mask.data_ptr(), m_sb, m_sh, m_sq, _lib.dtype_id(query.dtype),
                                       _lib.stream_ptr(dev))
    _lib.check(lib, rc, "sdpa")
    return out


# ------------------------------------------------------------------------------------------------------
# MojoGemmAllReduce
# ------------------------------------------------------------------------------------------------------
def gemm_allreduce_workspace_bytes(max_m: int, n: int, world: int) -> int:
    return int(_lib.load().mojo_b200_gemm_allreduce_workspace_bytes(int(max_m), int(n), int(world)))


@_on_tensor_device
def gemm_allreduce(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, workspace=None,
                   workspace_max_m: int = 0) -> torch.Tensor:
    """``all_reduce_sum(x @ weight.T + bias)`` in one kernel.  ``weight`` is ``[out_features, in_features_local]``;
    ``workspace`` is a ``comm.SymmetricWorkspace`` (or a ``comm.LocalRanks`` view) sized by
    ``gemm_allreduce_workspace_bytes(workspace_max_m, out_features, world)``; ``None`` = single rank, plain GEMM."""
    dev = _require_cuda(x, weight, bias)
    lib = _lib.load()
    if x.dtype != weight.dtype or (bias is not None and bias.dtype != x.dtype):
        raise NotImplementedError("gemm_allreduce: input, weight and bias must share one dtype")
    if weight.dim() != 2 or x.shape[-1] != weight.shape[1]:
        raise ValueError(f"gemm_allreduce: input {tuple(x.shape)} does not match weight {tuple(weight.shape)}")
    n, k = weight.shape
    if bias is not None and tuple(bias.shape) != (n,):
        raise ValueError(f"gemm_allreduce: bias shape {tuple(bias.shape)} != ({n},)")
    x2 = _as_rows(x)
    w = weight if weight.stride(1) == 1 else weight.contiguous()
    if x2.stride(0) % 8 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    if (w.stride(0) % 8 or w.data_ptr() % 16) and k % 8 == 0:
        w = w.contiguous()
    m = x2.shape[0]
    out = torch.empty((m, n), dtype=x.dtype, device=dev)
    b = None if bias is None else bias.contiguous()
    world = 1 if workspace is None else workspace.world
    if world > 1:
        table, ws_bytes, rank = workspace.table, workspace.nbytes, workspace.rank
    else:
        table, ws_bytes, rank = None, 0, 0
    rc = lib.mojo_b200_gemm_allreduce(
        x2.data_ptr(), w.data_ptr(), _lib.ptr(b), out.data_ptr(), m, n, k, x2.stride(0), w.stride(0), out.stride(0),
        table, ws_bytes, int(workspace_max_m), world, rank, _lib.dtype_id(x.dtype), _lib.stream_ptr(dev))
    _lib.check(lib, rc, "gemm_allreduce")
    return out.view(*x.shape[:-1], n)


# ------------------------------------------------------------------------------------------------------
# MojoNormRoPEStoreKV / MojoRoPEStoreKV
# ------------------------------------------------------------------------------------------------------
@_on_tensor_device
def norm_rope_store_kv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
                       key_cache: torch.Tensor, value_cache: torch.Tensor, block_table: torch.Tensor,
                       cu_q_lens: Optional[torch.Tensor], context_kv_lens: torch.Tensor,
                       q_norm_weight: Optional[torch.Tensor] = None, k_norm_weight: Optional[torch.Tensor] = None,
                       eps: float = 1e-6, want_k: bool = False):
    """One pass: optional per-head RMSNorm of q and k, RoPE on both, k / v scattered into their page slots.
    ``q, k, v`` are token-major ``[T, heads, D]`` (any token / head strides); returns ``q_rot`` (and ``k_rot`` when
    ``want_k``); the caches are updated in place."""
    dev = _require_cuda(q, k, v, cos, sin, key_cache, value_cache, block_table, cu_q_lens, context_kv_lens,
                        q_norm_weight, k_norm_weight)
    lib = _lib.load()
    if q.dim() != 3 or k.dim() != 3 or v.dim() != 3 or k.shape != v.shape or q.shape[0] != k.shape[0]:
        raise ValueError("norm_rope_store_kv: q [T,Hq,D], k/v [T,Hkv,D] expected")
    if not (q.dtype == k.dtype == v.dtype == key_cache.dtype == value_cache.dtype):
        raise NotImplementedError("norm_rope_store_kv: q, k, v and the caches must share one dtype")
    if key_cache.dim() != 4 or key_cache.shape != value_cache.shape:
        raise ValueError("norm_rope_store_kv: caches must be [num_blocks, kv_heads, block_size, head_dim]")
    tokens, hq, d = q.shape
    hkv = k.shape[1]
    num_blocks, c_heads, block_size, c_d = key_cache.shape
    if (c_heads, c_d) != (hkv, d) or q.shape[2] != d:
        raise ValueError("norm_rope_store_kv: head counts / head_dim of states and caches differ")
    if key_cache.stride(-1) != 1 or value_cache.stride(-1) != 1:
        raise NotImplementedError("norm_rope_store_kv: cache head_dim must be contiguous (in-place op)")
    if (q_norm_weight is None) != (k_norm_weight is None):
        raise ValueError("norm_rope_store_kv: give both norm weights or neither")
    q, k, v = _inner_contiguous(q), _inner_contiguous(k), _inner_contiguous(v)
    if cos.dim() != 2 or cos.shape[0] != tokens or sin.shape != cos.shape:
        raise ValueError("norm_rope_store_kv: cos/sin must be [T, rope_dim]")
    if cos.stride(-1) != 1 or sin.stride() != cos.stride() or sin.dtype != cos.dtype:
        cos, sin = cos.contiguous(), sin.contiguous().to(cos.dtype)
    table = block_table if block_table.stride(-1) == 1 or block_table.shape[1] <= 1 else block_table.contiguous()
    ctx = context_kv_lens.contiguous()
    cu = None if cu_q_lens is None else cu_q_lens.contiguous()
    wq = None if q_norm_weight is None else q_norm_weight.detach().contiguous()
    wk = None if k_norm_weight is None else k_norm_weight.detach().contiguous()
    q_out = torch.empty((tokens, hq, d), dtype=q.dtype, device=dev)
    k_out = torch.empty((tokens, hkv, d), dtype=q.dtype, device=dev) if want_k else None
    rc = lib.mojo_b200_norm_rope_store_kv(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), _lib.ptr(wq), _lib.ptr(wk), float(eps), cos.data_ptr(),
        sin.data_ptr(), q_out.data_ptr(), _lib.ptr(k_out), key_cache.data_ptr(), value_cache.data_ptr(),
        table.data_ptr(), table.stride(0) if table.shape[1] else 0, table.shape[1], _lib.ptr(cu), ctx.data_ptr(),
        ctx.shape[0], tokens, hq, hkv, d, cos.shape[-1], num_blocks, block_size,
        q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), q_out.stride(0), q_out.stride(1),
        k_out.stride(0) if want_k else 0, k_out.stride(1) if want_k else 0, cos.stride(0),
        key_cache.stride(0), key_cache.stride(1), key_cache.stride(2),
        value_cache.stride(0), value_cache.stride(1), value_cache.stride(2),
        _lib.dtype_id(q.dtype), _lib.dtype_id(cos.dtype), _lib.stream_ptr(dev))
    _lib.check(lib, rc, "norm_rope_store_kv")
    return (q_out, k_out) if want_k else q_out


# ------------------------------------------------------------------------------------------------------
# DiT block: MojoGelu / MojoLayerNorm / MojoGridRoPE
# ------------------------------------------------------------------------------------------------------
@_on_tensor_device
def gelu(x: torch.Tensor) -> torch.Tensor:
    dev = _require_cuda(x)
    lib = _lib.load()
    out = torch.empty(x.shape, dtype=x.dtype, device=dev)
    if x.numel() == 0:
        return out
    xr, rows, cols, x_rs = _rows_cols(x)
    rc = lib.mojo_b200_gelu(xr.data_ptr(), out.data_ptr(), rows, cols, x_rs, cols, _lib.dtype_id(x.dtype),
                            _lib.stream_ptr(dev))
    _lib.check(lib, rc, "gelu")
    return out


@_on_tensor_device
def layer_norm(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    dev = _require_cuda(x, weight, bias)
    lib = _lib.load()
    hidden = x.shape[-1]
    for name, t in (("weight", weight), ("bias", bias)):
        if t is not None and (t.dtype != x.dtype or tuple(t.shape) != (hidden,)):
            raise NotImplementedError(f"layer_norm: {name} must be [{hidden}] of dtype {x.dtype}")
    xr = _as_rows(x)
    y = torch.empty(x.shape, dtype=x.dtype, device=dev)
    if x.numel() == 0:
        return y
    w = None if weight is None else weight.detach().contiguous()
    b = None if bias is None else bias.detach().contiguous()
    rc = lib.mojo_b200_layer_norm(xr.data_ptr(), _lib.ptr(w), _lib.ptr(b), y.data_ptr(), xr.shape[0], hidden,
                                  xr.stride(0), hidden, float(eps), _lib.dtype_id(x.dtype), _lib.stream_ptr(dev))
    _lib.check(lib, rc, "layer_norm")
    return y


@_on_tensor_device
def grid_rope(x: torch.Tensor, freqs_list) -> torch.Tensor:
    """``x [B, L, N, D]`` (interleaved complex pairs); ``freqs_list[b]``: complex ``[seq_len_b, 1, D/2]``.  The sequence
    length of a sample is the phase table's (no host read of ``grid_sizes``); one launch per sample."""
    dev = _require_cuda(x, *freqs_list)
    lib = _lib.load()
    if x.dim() != 4 or x.shape[-1] % 2:
        raise ValueError("grid_rope: x must be [B, L, N, D] with even D")
    if len(freqs_list) != x.shape[0]:
        raise ValueError("grid_rope: one phase table per sample expected")
    x = _inner_contiguous(x)
    batch, tokens, heads, d = x.shape
    out = torch.empty((batch, tokens, heads, d), dtype=x.dtype, device=dev)
    for i, fr in enumerate(freqs_list):
        if not fr.is_complex() or fr.shape[-1] != d // 2 or fr.numel() != fr.shape[0] * (d // 2):
            raise ValueError(f"grid_rope: freqs_list[{i}] must be complex [seq_len, 1, {d // 2}]")
        ph = torch.view_as_real(fr.to(torch.complex64).reshape(fr.shape[0], d // 2).contiguous())
        seq_len = min(fr.shape[0], tokens)
        rc = lib.mojo_b200_grid_rope(x[i].data_ptr(), ph.data_ptr(), out[i].data_ptr(), seq_len, tokens, heads, d,
                                     x.stride(1), x.stride(2), out.stride(1), out.stride(2), d,
                                     _lib.dtype_id(x.dtype), _lib.stream_ptr(dev))
        _lib.check(lib, rc, "grid_rope")
    return out
