import os

import torch
import torch.distributed as dist

from mojo_opset_b200 import functional as F
from mojo_opset_b200.comm import SymmetricWorkspace
from mojo_opset_b200.core import MojoGemmAllReduce


class B200GemmAllReduce(MojoGemmAllReduce):
    """GEMM and all-reduce in ONE persistent sm_100a kernel (tcgen05 GEMM, partial tiles pushed over NVLink,
    owner-side fp32 reduction, broadcast) - ``csrc/gemm_allreduce.cu``.  The peer-mapped workspace is created on
    the first distributed call and grows when a larger token count arrives (``MOJO_B200_GAR_MAX_TOKENS`` presets
    it); all ranks must call with the same number of rows, as with any all-reduce.  The constructor is the core
    op's: all backend state is created lazily in ``forward``."""

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

    def _workspace(self, m: int, n: int):
        if not (dist.is_available() and dist.is_initialized()):
            return None, 0
        world = dist.get_world_size(self.process_group)
        if world == 1:
            return None, 0
        ws, max_m = getattr(self, "_b200_ws", None), getattr(self, "_b200_ws_max_m", 0)
        if ws is None or m > max_m:
            want = max(m, int(os.environ.get("MOJO_B200_GAR_MAX_TOKENS", "0")), 2 * max_m)
            if ws is not None:
                ws.close()
            ws = SymmetricWorkspace(F.gemm_allreduce_workspace_bytes(want, n, world), self.process_group)
            self._b200_ws, self._b200_ws_max_m, max_m = ws, want, want
        return ws, max_m

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        w = self._kmajor_weight()
        m = input.numel() // max(input.shape[-1], 1)
        ws, max_m = self._workspace(m, w.shape[0])
        return F.gemm_allreduce(input, w, self.bias, ws, max_m)
