"""TEST INFRASTRUCTURE - ctypes access to the plain-C KV-store oracle (oracle/store_kv.c)."""

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libstore_kv_oracle.so")


def _load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "store_kv.c")):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    lib = ctypes.CDLL(_LIB)
    lib.oracle_build_chunk_plan.restype = ctypes.c_int64
    lib.oracle_build_chunk_plan.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]
    lib.oracle_store_paged_kv.restype = None
    lib.oracle_store_paged_kv.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 5
    return lib


def build_chunk_plan(block_table, cu_q_lens, context_kv_lens, block_size):
    lib = _load()
    table = block_table.contiguous()
    ctx = context_kv_lens.contiguous()
    cu = None if cu_q_lens is None else cu_q_lens.contiguous()
    num_seqs, width = table.shape
    cap = max(1, num_seqs * max(width, 1))
    plan = torch.empty(cap, 4, dtype=torch.int32)
    n = lib.oracle_build_chunk_plan(table.data_ptr(), num_seqs, width, None if cu is None else cu.data_ptr(),
                                    ctx.data_ptr(), block_size, plan.data_ptr(), cap)
    return plan[:n].clone()


def store_paged_kv(states, cache, plan):
    """In place on a contiguous CPU cache; returns it."""
    lib = _load()
    assert states.is_contiguous() and cache.is_contiguous() and plan.is_contiguous()
    _, heads, block_size, dim = cache.shape
    lib.oracle_store_paged_kv(states.data_ptr(), cache.data_ptr(), plan.data_ptr(), plan.shape[0], heads, dim,
                              block_size, states.element_size())
    return cache
