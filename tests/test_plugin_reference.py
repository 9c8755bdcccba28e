"""The out-of-tree registration path against the UNMODIFIED reference (build container only: the reference does not
travel to the GPU box, so this test skips there).  INTEGRATION.md section 1 is what is being checked."""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("MOJO_REFERENCE_ROOT", "/root/reference")

SCRIPT = r"""
import os, torch
import mojo_opset                                   # the reference
from mojo_opset_b200 import plugin
made = plugin.register()
assert sorted(made) == sorted(plugin.OPS), sorted(made)
assert plugin.register() == {}                      # idempotent
import mojo_opset.experimental as experimental       # MojoGridRoPE still lives there upstream
for op in plugin.OPS:
    core = getattr(mojo_opset, "Mojo" + op, None) or getattr(experimental, "Mojo" + op)
    assert core.get_registered_backends()[:2] == ("b200", "torch"), (op, core.get_registered_backends())
    impl = core.get_backend_impl("b200", strict=True)
    assert impl.__name__ == "B200" + op and impl.__base__ is core           # direct subclass (TP wrapper lookup)
    assert core.get_backend_impl("torch").__name__ == "Torch" + op          # the golden is still there
os.environ["MOJO_BACKEND"] = "b200"
dec = mojo_opset.MojoPagedDecodeGQA(is_causal=True, gqa_layout="ABAB")
assert type(dec).__name__ == "B200PagedDecodeGQA" and dec.gqa_layout == "ABAB"
norm = mojo_opset.MojoResidualAddRMSNorm(64, eps=1e-6, norm_pos="post")
assert type(norm).__name__ == "B200ResidualAddRMSNorm" and norm.weight.shape == (64,) and norm.norm_pos == "post"
os.environ["MOJO_BACKEND"] = "torch"
assert type(mojo_opset.MojoSwiGLU()).__name__ == "TorchSwiGLU"
os.environ["MOJO_BACKEND"] = "b200"
if not torch.cuda.is_available():                   # no CPU fallback: the op refuses CPU tensors
    try:
        mojo_opset.MojoSwiGLU()(torch.randn(2, 8), torch.randn(2, 8))
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("B200SwiGLU computed on the CPU")
print("PLUGIN_OK")
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "mojo_opset")), reason="reference tree not present")
def test_plugin_registers_b200_backend_in_the_reference():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REFERENCE, ROOT]), PYTHONDONTWRITEBYTECODE="1",
               MOJO_OPSET_PLUGIN_AUTOLOAD="0")
    env.pop("MOJO_BACKEND", None)
    res = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "PLUGIN_OK" in res.stdout, res.stdout + res.stderr
