"""Swap Mojo ops into HuggingFace Qwen3 (the ``examples/qwen3_patch.py`` drop-in).

Mirror of the reference's ``mojo_opset/utils/patching.py:4-59`` (``apply_mojo_to_qwen3``): same name, same keyword
arguments, same three static replacements made BEFORE the model is instantiated -

* ``modeling_qwen3.apply_rotary_pos_emb`` -> a ``MojoApplyRoPE`` instance (HF calls it as ``(q, k, cos, sin)`` with
  ``q/k [B, H, S, D]`` transposed views and ``cos/sin [B, S, D]``; ``head_first=True`` is the op's default),
* ``modeling_qwen3.Qwen3RMSNorm``        -> ``MojoRMSNorm`` (``(hidden_size, eps)`` ctor, ``weight`` parameter),
* ``modeling_qwen3.Qwen3MLP``            -> an MLP whose activation is ``MojoSwiGLU`` (same parameter names, so a
  HF checkpoint / state dict loads unchanged).

Which kernels run is decided by ``MOJO_BACKEND`` at op construction, exactly as everywhere else; with the b200
backend there is no CPU fallback.  ``revert_mojo_from_qwen3`` restores the originals (the reference has no such
helper; the tests need it to compare a patched and an unpatched model in one process).
"""

from typing import Dict

_ORIGINALS: Dict[str, object] = {}


def _remember(module, name: str) -> None:
    if name not in _ORIGINALS:
        _ORIGINALS[name] = getattr(module, name)


def apply_mojo_to_qwen3(
    rope: bool = True,
    cross_entropy: bool = False,
    fused_linear_cross_entropy: bool = True,
    rms_norm: bool = True,
    swiglu: bool = True,
    model=None,
) -> None:
    import torch.nn as nn

    from transformers.models.qwen3 import modeling_qwen3

    from mojo_opset_b200 import MojoApplyRoPE
    from mojo_opset_b200 import MojoRMSNorm
    from mojo_opset_b200 import MojoSwiGLU

    # the reference only validates these two (its loss functions are training-side, outside this backend's path)
    assert not (cross_entropy and fused_linear_cross_entropy), (
        "cross_entropy and fused_linear_cross_entropy cannot both be True."
    )
    if model is not None:
        # reference patching.py:61-79 keeps instance patching commented out: static replacement only
        raise NotImplementedError("apply_mojo_to_qwen3: patch before the model is instantiated (model=None)")

    if rope:
        _remember(modeling_qwen3, "apply_rotary_pos_emb")
        modeling_qwen3.apply_rotary_pos_emb = MojoApplyRoPE()

    if rms_norm:
        _remember(modeling_qwen3, "Qwen3RMSNorm")
        modeling_qwen3.Qwen3RMSNorm = MojoRMSNorm

    if swiglu:
        _remember(modeling_qwen3, "Qwen3MLP")

        class MojoSwiGLUMLP(nn.Module):
            """gate/up/down projections (cuBLAS, plumbing) around one fused ``silu(gate) * up`` pass."""

            def __init__(self, config):
                super().__init__()
                if config.hidden_act != "silu":
                    raise ValueError(f"MojoSwiGLUMLP requires 'silu' activation, but got {config.hidden_act}")
                self.config = config
                self.hidden_size = config.hidden_size
                self.intermediate_size = config.intermediate_size
                self.gate_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
                self.up_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
                self.down_proj = nn.Linear(self.intermediate_size, self.hidden_size, bias=False)

            def forward(self, x):
                # the op is resolved per call like the reference does (patching.py:50-57): MOJO_BACKEND may change
                # between model construction and the first forward, and the ctor is a dictionary lookup
                return self.down_proj(MojoSwiGLU()(self.gate_proj(x), self.up_proj(x)))

        modeling_qwen3.Qwen3MLP = MojoSwiGLUMLP


def revert_mojo_from_qwen3() -> None:
    """Undo ``apply_mojo_to_qwen3`` (models built while patched keep their Mojo modules)."""
    from transformers.models.qwen3 import modeling_qwen3

    for name, original in _ORIGINALS.items():
        setattr(modeling_qwen3, name, original)
    _ORIGINALS.clear()
