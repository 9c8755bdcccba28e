import torch

from mojo_opset_b200 import functional as F
from mojo_opset_b200.core import MojoLayerNorm
from mojo_opset_b200.core import MojoResidualAddRMSNorm
from mojo_opset_b200.core import MojoRMSNorm


class B200RMSNorm(MojoRMSNorm):
    supported_platforms_list = ["b200"]

    def forward(self, hidden_state: torch.Tensor) -> torch.Tensor:
        return F.rms_norm(hidden_state, self.weight, self.variance_epsilon)


class B200ResidualAddRMSNorm(MojoResidualAddRMSNorm):
    supported_platforms_list = ["b200"]

    def forward(self, hidden_state: torch.Tensor, residual: torch.Tensor):
        pre = self.norm_pos == "pre"
        y, summed = F.residual_add_rms_norm(hidden_state, residual, self.weight, self.variance_epsilon, want_sum=pre)
        # "post" returns the normalised tensor twice (reference normalization.py:350-359)
        return (y, summed) if pre else (y, y)


class B200LayerNorm(MojoLayerNorm):
    supported_platforms_list = ["b200"]

    def forward(self, hidden_state: torch.Tensor) -> torch.Tensor:
        return F.layer_norm(hidden_state, self.weight, self.bias, self.variance_epsilon)
