"""ctypes binding of ``libmojo_b200.so`` (C ABI: ``include/mojo_b200.h``).

The library is loaded lazily on first use and the load FAILS LOUDLY: a missing ``.so`` (not built), a
missing symbol or a non-sm_100 device raise ``RuntimeError`` - there is no CPU or eager fallback.
"""

import ctypes
import os
import threading

from ctypes import c_char_p
from ctypes import c_float
from ctypes import c_int
from ctypes import c_int64
from ctypes import c_size_t
from ctypes import c_void_p

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# MOJO_B200_LIB: developer override (A/B builds of the same ABI, tools/build_variant.sh); the default is the in-tree build
LIB_PATH = os.environ.get("MOJO_B200_LIB") or os.path.join(_PKG_DIR, "libmojo_b200.so")

EINVAL, EUNSUPPORTED, EWORKSPACE = -1, -2, -3
BF16, F16, F32 = 0, 1, 2
_DTYPE_IDS = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}

P, I, L, F, Z = c_void_p, c_int, c_int64, c_float, c_size_t

# name -> (restype, argtypes); must list every function include/mojo_b200.h declares
SIGNATURES = {
    "mojo_b200_abi_version": (I, []),
    "mojo_b200_last_error": (c_char_p, []),
    "mojo_b200_device_ok": (I, []),
    "mojo_b200_set_error_word": (I, [P]),
    "mojo_b200_set_decode_tickets": (I, [P, L]),
    "mojo_b200_store_paged_kv_chunks": (I, [P, P, P, P, P, L, L, I, I, L, I] + [L] * 10 + [I, P]),
    "mojo_b200_store_paged_kv_table": (I, [P, P, P, P, P, L, I, P, P, I, L, I, I, L, I] + [L] * 10 + [I, P]),
    "mojo_b200_rms_norm": (I, [P, P, P, L, I, L, L, F, I, P]),
    "mojo_b200_residual_add_rms_norm": (I, [P, P, P, P, P, L, I, L, L, L, L, F, I, P]),
    "mojo_b200_apply_rope": (I, [P, P, P, P, P, P, L, L, I, I, I, I] + [L] * 14 + [I, I, P]),
    "mojo_b200_rotary_cos_sin": (I, [P, P, L, I, P, F, P, P, P, I, L, P, P, L, P]),
    "mojo_b200_swiglu": (I, [P, P, P, L, L, L, L, L, F, I, P]),
    "mojo_b200_silu": (I, [P, P, L, L, L, L, I, P]),
    "mojo_b200_paged_decode_num_splits": (I, [I, I, I, I, I, L, I]),
    "mojo_b200_paged_decode_workspace_bytes": (Z, [I, I, I, I]),
    "mojo_b200_paged_decode_gqa": (I, [P, P, P, P, P, P, P, Z, I, I, I, I, L, I, I, L, L] + [L] * 10 + [F, I, I, I, P]),
    "mojo_b200_paged_decode_swa": (I, [P, P, P, P, P, P, P, Z, I, I, I, I, L, I, I, L, L] + [L] * 10
                                   + [F, I, I, I, I, I, P]),
    "mojo_b200_paged_prefill_gqa": (I, [P, P, P, P, P, P, P, L, I, I, I, I, L, I, I, L, L, L] + [L] * 10
                                    + [F, I, I, I, P]),
    "mojo_b200_paged_prefill_swa": (I, [P, P, P, P, P, P, P, L, I, I, I, I, L, I, I, L, L, L] + [L] * 10
                                    + [F, I, I, I, I, I, P]),
    "mojo_b200_swa": (I, [P, P, P, P, P, P, L, L, I, I, I, I] + [L] * 10 + [F, I, I, I, I, I, P]),
    "mojo_b200_sdpa": (I, [P, P, P, P, I, I, I, L, L, I] + [L] * 12 + [F, I, P]),
    "mojo_b200_sdpa_masked": (I, [P, P, P, P, I, I, I, L, L, I] + [L] * 12 + [F, P, L, L, L, I, P]),
    "mojo_b200_norm_rope_store_kv": (I, [P, P, P, P, P, F, P, P, P, P, P, P, P, L, I, P, P, I, L, I, I, I, I, L, I]
                                     + [L] * 17 + [I, I, P]),
    "mojo_b200_gelu": (I, [P, P, L, L, L, L, I, P]),
    "mojo_b200_layer_norm": (I, [P, P, P, P, L, I, L, L, F, I, P]),
    "mojo_b200_grid_rope": (I, [P, P, P, L, L, I, I, L, L, L, L, L, I, P]),
    "mojo_b200_paged_reserve": (I, [P, L, I, P, P, P, P, P, I, I, P, P]),
    "mojo_b200_paged_positions": (I, [P, P, P, I, L, P]),
    "mojo_b200_symm_alloc": (I, [Z, ctypes.POINTER(c_void_p)]),
    "mojo_b200_symm_free": (I, [P]),
    "mojo_b200_symm_export": (I, [P, P]),
    "mojo_b200_symm_open": (I, [P, ctypes.POINTER(c_void_p)]),
    "mojo_b200_symm_close": (I, [P]),
    "mojo_b200_gemm_allreduce_workspace_bytes": (Z, [L, L, I]),
    "mojo_b200_gemm_allreduce": (I, [P, P, P, P, L, L, L, L, L, L, ctypes.POINTER(c_void_p), Z, L, I, I, I, P]),
}

_lock = threading.Lock()
_lib = None


def dtype_id(dtype: torch.dtype) -> int:
    try:
        return _DTYPE_IDS[dtype]
    except KeyError:
        raise NotImplementedError(f"b200 backend supports bf16/fp16/fp32 tensors, got {dtype}") from None


def load(check_device: bool = True):
    """Return the loaded library (cached).  Raises RuntimeError when it cannot serve requests."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m mojo_opset_b200.build` "
                "(needs nvcc; the b200 backend has no CPU fallback)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                raise RuntimeError(f"{LIB_PATH} does not export {name}: stale build?") from None
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.mojo_b200_abi_version() != 1:
            raise RuntimeError(f"{LIB_PATH}: unexpected ABI version {lib.mojo_b200_abi_version()}")
        if check_device:
            if not torch.cuda.is_available():
                raise RuntimeError("mojo_opset_b200 needs an sm_100 (B200) GPU: no CUDA device is visible")
            if lib.mojo_b200_device_ok() != 1:
                raise RuntimeError("mojo_opset_b200 kernels are built for sm_100a only; current device is not sm_100")
        _lib = lib
        return lib


def last_error(lib) -> str:
    msg = lib.mojo_b200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(lib, rc: int, what: str) -> None:
    """Map a C-ABI return code to the reference's error conventions."""
    if rc == 0:
        return
    msg = f"{what}: {last_error(lib)}"
    if rc == EUNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == EINVAL:
        raise ValueError(msg)
    raise RuntimeError(f"{msg} (code {rc})")


_error_words = {}

ERR_DECODE_UNMAPPED_BLOCK, ERR_PREFILL_UNMAPPED_BLOCK = 1, 2
DECODE_TICKETS = 1 << 16  # ints per device (256 KiB): 16 slices of 4096 (sequence, kv head) groups
_decode_tickets = {}


def error_word(device) -> torch.Tensor:
    """The int32 device word the kernels OR error bits into (one per device, registered with the library on first
    use).  Allocated outside any graph capture by ``functional`` entry points that can raise data errors."""
    device = torch.device(device)
    index = torch.cuda.current_device() if device.index is None else device.index
    word = _error_words.get(index)
    if word is None:
        if torch.cuda.is_current_stream_capturing():
            return None  # registered by the first eager call; kernels skip the check while there is no word
        word = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", index))
        # ... and the split-KV decode's arrival counters (zero between launches: the kernels reset what they count)
        tickets = torch.zeros(DECODE_TICKETS, dtype=torch.int32, device=torch.device("cuda", index))
        with torch.cuda.device(index):
            check(load(), load().mojo_b200_set_error_word(word.data_ptr()), "set_error_word")
            check(load(), load().mojo_b200_set_decode_tickets(tickets.data_ptr(), DECODE_TICKETS), "set_decode_tickets")
        _error_words[index] = word
        _decode_tickets[index] = tickets
    return word


def check_device_errors(device=None) -> None:
    """Read (synchronises) and clear the device error word; raise what the reference raises on the host:
    ``ValueError`` for a sequence with keys whose first block is unmapped (reference ``attention.py:186-187, 396-397``)."""
    devices = list(_error_words) if device is None else [torch.device(device).index or 0]
    for index in devices:
        word = _error_words.get(index)
        if word is None:
            continue
        bits = int(word.item())
        if bits:
            word.zero_()
        if bits & ERR_DECODE_UNMAPPED_BLOCK:
            raise ValueError("Paged decode requires a valid block table for rows with kv lens > 0.")
        if bits & ERR_PREFILL_UNMAPPED_BLOCK:
            raise ValueError("Paged prefill requires a valid block table for rows with kv lens > 0.")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
