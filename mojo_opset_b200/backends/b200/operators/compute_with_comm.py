import os
import weakref

import torch
import torch.distributed as dist

from mojo_opset_b200 import functional as F
from mojo_opset_b200.comm import SymmetricWorkspace
from mojo_opset_b200.core import MojoGemmAllReduce

# One peer-mapped workspace per (process group, device, out_features), shared by every ``B200GemmAllReduce`` of a
# model (an 80-layer stack has 80 ``o_proj`` ops of one shape: one workspace, not 80).  The kernel's epoch / parity
# protocol already serialises back-to-back calls on one workspace (``csrc/gemm_allreduce.cu``); calls on the same
# stream are ordered anyway.
_WORKSPACES = {}


def _group_key(group):
    return id(group) if group is not None else 0


def shared_workspace(group, device: torch.device, n: int, m: int):
    """The shared workspace for ``m`` rows of ``n`` output features; created (collectively!) on first use, sized for
    ``max(m, MOJO_B200_GAR_MAX_TOKENS, 256)`` rows and regrown (collectively, geometric) when a larger ``m`` arrives.
    Creation exchanges IPC handles and synchronises, so it cannot happen while a CUDA graph is being captured:
    call the op once eagerly with the largest row count (or set ``MOJO_B200_GAR_MAX_TOKENS``) before capturing."""
    world = dist.get_world_size(group)
    key = (_group_key(group), device.index, int(n))
    entry = _WORKSPACES.get(key)
    if entry is not None and m <= entry[1]:
        return entry
    if torch.cuda.is_current_stream_capturing():
        have = 0 if entry is None else entry[1]
        raise RuntimeError(
            f"B200GemmAllReduce: the peer-mapped workspace holds {have} rows and {m} are needed, but it cannot be "
            "(re)allocated during CUDA-graph capture; run the op once eagerly at the largest token count or set "
            "MOJO_B200_GAR_MAX_TOKENS before capturing")
    want = max(int(m), int(os.environ.get("MOJO_B200_GAR_MAX_TOKENS", "0")), 256, 2 * (entry[1] if entry else 0))
    if entry is not None:
        entry[0].close()
    ws = SymmetricWorkspace(F.gemm_allreduce_workspace_bytes(want, n, world), group, device)
    entry = (ws, want)
    _WORKSPACES[key] = entry
    if group is not None:  # drop the cache entry with the group (ids are reused)
        try:
            weakref.finalize(group, _WORKSPACES.pop, key, None)
        except TypeError:
            pass
    return entry


def release_workspaces():
    """Unmap and free every shared workspace (collective; call before ``destroy_process_group``)."""
    for key in list(_WORKSPACES):
        ws, _ = _WORKSPACES.pop(key)
        ws.close()


class B200GemmAllReduce(MojoGemmAllReduce):
    """GEMM and all-reduce in ONE persistent sm_100a kernel (tcgen05 GEMM, partial tiles pushed over NVLink,
    owner-side fp32 reduction, broadcast) - ``csrc/gemm_allreduce.cu``.  The peer-mapped workspace is shared by all
    instances of one (process group, out_features) and is created on the first distributed call (see
    ``shared_workspace``); all ranks must call with the same number of rows, as with any all-reduce.  The constructor
    is the core op's: all backend state is created lazily in ``forward``."""

    supported_platforms_list = ["b200"]

    def _kmajor_weight(self) -> torch.Tensor:
        # the kernel reads the weight K-major ([out_features, in_features_local]); a transposed layout is re-packed
        # once (and again only if the weight tensor is replaced, e.g. by weight loading), not per call
        if not self.trans_weight:
            return self.weight
        cached = getattr(self, "_b200_w", None)
        if cached is None or cached[0] is not self.weight or cached[1] != self.weight._version:
            cached = (self.weight, self.weight._version, self.weight.t().contiguous())
            self._b200_w = cached
        return cached[2]

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        w = self._kmajor_weight()
        m = input.numel() // max(input.shape[-1], 1)
        ws, max_m = None, 0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            ws, max_m = shared_workspace(self.process_group, input.device, w.shape[0], m)
        return F.gemm_allreduce(input, w, self.bias, ws, max_m)
