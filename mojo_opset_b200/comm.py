"""Symmetric (peer-mapped) device workspaces for the fused compute + collective kernels.

One process per GPU.  Every rank allocates the same number of bytes through the C ABI (``cudaMalloc``, zero
filled), exports a CUDA IPC handle, the handles travel through ``torch.distributed`` (plumbing) and every rank maps
its peers' allocations, so a kernel can store into / load from any rank's workspace over NVLink.  The reference's
analogue is ``MojoSymmetricMemoryManager`` (``mojo_opset/backends/ttx/operators/compute_with_comm.py:102-133``).

``LocalRanks`` builds the same structure for several *virtual* ranks inside one process on one GPU (plain
pointers, no IPC): the single-GPU tests drive the full protocol that way, one stream per virtual rank.
"""

import ctypes
import os

from typing import List
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib


class SymmetricWorkspace:
    """``bytes`` of zero-initialised device memory on every rank of ``group``, mapped into every rank."""

    def __init__(self, nbytes: int, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("SymmetricWorkspace needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nbytes = int(nbytes)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = lib = _lib.load()
        self._local = ctypes.c_void_p()
        self._peers: List[Optional[int]] = [None] * self.world
        with torch.cuda.device(self.device):
            _lib.check(lib, lib.mojo_b200_symm_alloc(self.nbytes, ctypes.byref(self._local)), "symm_alloc")
            handle = ctypes.create_string_buffer(64)
            _lib.check(lib, lib.mojo_b200_symm_export(self._local, handle), "symm_export")
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            for r, h in enumerate(handles):
                if r == self.rank:
                    self._peers[r] = self._local.value
                    continue
                p = ctypes.c_void_p()
                _lib.check(lib, lib.mojo_b200_symm_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)),
                           f"symm_open(rank {r})")
                self._peers[r] = p.value
        self.table = (ctypes.c_void_p * self.world)(*self._peers)
        dist.barrier(group=group)  # nobody launches before every mapping exists

    def close(self):
        if self._lib is None:
            return
        lib, self._lib = self._lib, None
        torch.cuda.synchronize(self.device)
        for r, p in enumerate(self._peers):
            if p is not None and r != self.rank:
                lib.mojo_b200_symm_close(ctypes.c_void_p(p))
        if dist.is_initialized():
            dist.barrier(group=self.group)  # peers have unmapped before the owner frees
        lib.mojo_b200_symm_free(self._local)

    def __del__(self):
        try:
            if self._lib is not None and not dist.is_initialized():
                self._lib.mojo_b200_symm_free(self._local)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class LocalRanks:
    """``world`` virtual ranks on ONE device (tests / single-GPU demos): rank r's workspace is a plain allocation,
    its "peer" pointers are the other allocations, each rank launches on its own stream."""

    class _View:
        def __init__(self, parent, rank):
            self.world, self.rank, self.nbytes, self.table = parent.world, rank, parent.nbytes, parent.table

    def __init__(self, world: int, nbytes: int):
        self.world, self.nbytes = world, int(nbytes)
        self._lib = lib = _lib.load()
        self._ptrs = []
        for _ in range(world):
            p = ctypes.c_void_p()
            _lib.check(lib, lib.mojo_b200_symm_alloc(self.nbytes, ctypes.byref(p)), "symm_alloc")
            self._ptrs.append(p)
        self.table = (ctypes.c_void_p * world)(*[p.value for p in self._ptrs])
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        # the ranks' persistent kernels wait for one another: all of them must be resident on this one GPU at once
        # (one CTA per SM each, ~194 KB of shared memory), so each gets at most SMs / world CTAs
        self._saved_cap = os.environ.get("MOJO_B200_GAR_MAX_CTAS")
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        os.environ["MOJO_B200_GAR_MAX_CTAS"] = str(max(1, sms // world))

    def view(self, rank: int) -> "_View":
        return LocalRanks._View(self, rank)

    def close(self):
        torch.cuda.synchronize()
        for p in self._ptrs:
            self._lib.mojo_b200_symm_free(p)
        self._ptrs = []
        if self._saved_cap is None:
            os.environ.pop("MOJO_B200_GAR_MAX_CTAS", None)
        else:
            os.environ["MOJO_B200_GAR_MAX_CTAS"] = self._saved_cap
