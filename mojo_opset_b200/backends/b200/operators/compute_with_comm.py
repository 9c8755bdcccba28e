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
    it); all ranks must call with the same number of rows, as with any all-reduce."""

    supported_platforms_list = ["b200"]

    def __init__(self, weight, bias=None, trans_weight: bool = False, process_group=None):
        super().__init__(weight, bias, trans_weight, process_group)
        # the kernel reads the weight K-major ([out_features, in_features_local]); a transposed layout is
        # re-packed once here, not per call
        self._w = weight.t().contiguous() if trans_weight else weight
        self._ws = None
        self._ws_max_m = 0

    def _workspace(self, m: int, n: int):
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(self.process_group)
        if world == 1:
            return None
        if self._ws is None or m > self._ws_max_m:
            want = max(m, int(os.environ.get("MOJO_B200_GAR_MAX_TOKENS", "0")), 2 * self._ws_max_m)
            if self._ws is not None:
                self._ws.close()
            self._ws_max_m = want
            self._ws = SymmetricWorkspace(F.gemm_allreduce_workspace_bytes(want, n, world), self.process_group)
        return self._ws

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if self._w.data_ptr() != self.weight.data_ptr() and not self.trans_weight:
            self._w = self.weight  # the parameter was re-assigned (weight loading)
        m = input.numel() // max(input.shape[-1], 1)
        ws = self._workspace(m, self._w.shape[0])
        return F.gemm_allreduce(input, self._w, self.bias, ws, self._ws_max_m)
