import torch

from mojo_opset_b200 import functional as F
from mojo_opset_b200.core import MojoGelu
from mojo_opset_b200.core import MojoSilu
from mojo_opset_b200.core import MojoSwiGLU


class B200Silu(MojoSilu):
    supported_platforms_list = ["b200"]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.silu(x)


class B200SwiGLU(MojoSwiGLU):
    supported_platforms_list = ["b200"]

    def forward(self, gate_out: torch.Tensor, up_out: torch.Tensor) -> torch.Tensor:
        return F.swiglu(gate_out, up_out, self.swiglu_limit)


class B200Gelu(MojoGelu):
    supported_platforms_list = ["b200"]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.gelu(x)
