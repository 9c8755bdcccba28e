"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the registry
dispatches like the reference's, the b200 ops refuse to run without a GPU, host-side argument logic."""

import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mojo_b200.h")).read()
    return sorted(set(re.findall(r"MOJO_B200_API\s+[\w\s\*]+?\b(mojo_b200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mojo_opset_b200 import _lib
    from mojo_opset_b200.build import build

    build()  # no-op when up to date; nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mojo_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(declared)
    lib.mojo_b200_abi_version.restype = ctypes.c_int
    assert lib.mojo_b200_abi_version() == 1


def test_no_cpu_fallback():
    """On a box without an sm_100 GPU the product refuses to compute instead of falling back."""
    if torch.cuda.is_available():
        pytest.skip("GPU box")
    import mojo_opset_b200 as m
    from mojo_opset_b200 import functional as F
    from mojo_opset_b200.backends.b200 import B200PagedDecodeGQA
    from mojo_opset_b200.backends.b200 import B200SwiGLU

    assert "b200" not in m.MojoSwiGLU.get_registered_backends()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200SwiGLU()(torch.randn(8), torch.randn(8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.rms_norm(torch.randn(2, 8), torch.randn(8), 1e-6)
    q = torch.randn(1, 4, 64)
    kc = torch.randn(2, 1, 16, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200PagedDecodeGQA()(q, kc, kc, torch.tensor([3], dtype=torch.int32), torch.zeros(1, 1, dtype=torch.int32))


def test_product_does_not_import_oracle():
    code = (
        "import sys; import mojo_opset_b200, mojo_opset_b200.functional, mojo_opset_b200.backends.b200;"
        "assert not any(n == 'oracle' or n.startswith('oracle.') for n in sys.modules), 'oracle imported by product'"
    )
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mojo_opset_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"


def test_registry_dispatch(torch_backend):
    """Reference tests/base/test_backend_dispatch.py:12-44, restated for the b200/torch pair."""
    import mojo_opset_b200 as m

    reg = m.MojoSilu.get_registry()
    assert reg is m.MojoSilu._registry
    torch_cls = m.MojoSilu.get_backend_impl("torch")
    assert torch_cls is reg.get("torch") and torch_cls.__name__ == "TorchSilu"
    assert reg.get(" Torch ") is torch_cls
    assert type(m.MojoSilu()) is m.MojoSilu.get_backend_impl()
    with pytest.raises(KeyError):
        m.MojoSilu.get_backend_impl("ttx", strict=True)
    assert m.MojoSilu.get_backend_impl("ttx") is m.MojoSilu.get_backend_impl()  # silent fallback, like upstream
    op = m.MojoPagedDecodeGQA(gqa_layout="ABAB")
    assert op.gqa_layout == "ABAB" and op.is_causal
    with pytest.raises(ValueError):
        m.MojoPagedDecodeGQA(gqa_layout="ABBA")
    with pytest.raises(NotImplementedError):
        op.forward_diff_with(m.MojoPagedDecodeGQA(gqa_layout="ABAB"))


def test_registration_rules():
    import mojo_opset_b200 as m

    with pytest.raises(AssertionError):

        class TTXSilu(m.MojoSilu):  # unknown backend prefix
            pass

    with pytest.raises(NameError):

        class B200xSilu(m.MojoSilu):  # typo of a known prefix
            pass


def test_b200_registers_on_platform():
    code = (
        "import os; os.environ['MOJO_BACKEND']='b200';"
        "import mojo_opset_b200 as m;"
        "ops=[m.MojoPagedDecodeGQA,m.MojoPagedPrefillGQA,m.MojoSdpa,m.MojoStorePagedKVCache,m.MojoResidualAddRMSNorm,"
        "m.MojoRMSNorm,m.MojoApplyRoPE,m.MojoRotaryEmbedding,m.MojoSwiGLU,m.MojoSilu];"
        "assert all(o.get_registered_backends()==('b200',) for o in ops);"
        "assert type(m.MojoSdpa()).__name__=='B200Sdpa' and type(m.MojoSdpa()).__base__ is m.MojoSdpa;"
        "import oracle.torch_backend;"
        "assert m.MojoSdpa.get_registered_backends()==('b200','torch');"
        "os.environ['MOJO_BACKEND']='torch'; assert type(m.MojoSdpa()).__name__=='TorchSdpa'"
    )
    env = dict(os.environ, MOJO_PLATFORM="b200")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT, env=env)


def test_contracts_raise_like_reference():
    from mojo_opset_b200.backends.b200 import B200ApplyRoPE
    from mojo_opset_b200.backends.b200 import B200PagedDecodeGQA
    from mojo_opset_b200.backends.b200 import B200PagedPrefillGQA
    from mojo_opset_b200.backends.b200 import B200Sdpa
    from mojo_opset_b200.backends.b200 import B200StorePagedKVCache

    q = torch.randn(2, 4, 64)
    kc = torch.randn(4, 1, 16, 64)
    lens64 = torch.tensor([3, 4])  # int64: contract violation
    tables = torch.zeros(2, 1, dtype=torch.int32)
    with pytest.raises(AssertionError):
        B200PagedDecodeGQA()(q, kc, kc, lens64, tables)
    with pytest.raises(NotImplementedError):
        B200PagedDecodeGQA()(q, kc, kc, lens64.int(), tables, mask=torch.ones(4, 4, dtype=torch.bool))
    with pytest.raises(NotImplementedError):
        B200PagedDecodeGQA(is_causal=False)(q, kc, kc, lens64.int(), tables)
    with pytest.raises(AssertionError):  # cu_total_seq_lens must be cumulative [B+1]
        B200PagedPrefillGQA()(q, kc, kc, torch.tensor([0, 1, 2], dtype=torch.int32), tables, None,
                              torch.tensor([3, 4], dtype=torch.int32))
    with pytest.raises(AssertionError):  # mixing the plan with the legacy triple
        B200StorePagedKVCache()(q, q, kc, kc, tables, chunk_metadata=torch.zeros(1, 4, dtype=torch.int32))
    with pytest.raises(AssertionError):
        B200ApplyRoPE()(q, q[None], torch.randn(2, 64), torch.randn(2, 64))
    with pytest.raises(AssertionError):
        B200ApplyRoPE(interleaved=True)


def test_qwen3_patch_swaps_and_reverts():
    """``apply_mojo_to_qwen3`` (reference utils/patching.py:4-59) replaces HF's rotary fn / RMSNorm / MLP classes
    statically and ``revert_mojo_from_qwen3`` restores them; no compute."""
    pytest.importorskip("transformers")
    from transformers.models.qwen3 import modeling_qwen3

    import mojo_opset_b200 as m
    from mojo_opset_b200.utils.patching import apply_mojo_to_qwen3
    from mojo_opset_b200.utils.patching import revert_mojo_from_qwen3

    orig = (modeling_qwen3.apply_rotary_pos_emb, modeling_qwen3.Qwen3RMSNorm, modeling_qwen3.Qwen3MLP)
    with pytest.raises(AssertionError):
        apply_mojo_to_qwen3(cross_entropy=True, fused_linear_cross_entropy=True)
    old = os.environ.get("MOJO_BACKEND")
    try:
        if not torch.cuda.is_available():
            os.environ.pop("MOJO_BACKEND", None)  # no b200 registration on a CPU box: the core op is the interface
        apply_mojo_to_qwen3()
        assert isinstance(modeling_qwen3.apply_rotary_pos_emb, m.MojoApplyRoPE)
        assert modeling_qwen3.Qwen3RMSNorm is m.MojoRMSNorm
        assert modeling_qwen3.Qwen3MLP.__name__ == "MojoSwiGLUMLP"
        apply_mojo_to_qwen3(rope=False, rms_norm=False, swiglu=False)  # a no-op keeps the first originals
    finally:
        revert_mojo_from_qwen3()
        if old is not None:
            os.environ["MOJO_BACKEND"] = old
    assert (modeling_qwen3.apply_rotary_pos_emb, modeling_qwen3.Qwen3RMSNorm, modeling_qwen3.Qwen3MLP) == orig


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/mojo_b200.h against its ctypes signature in _lib.SIGNATURES: same number of
    parameters and the same kind at every position (pointer / int / int64_t / size_t / float) - a marshalling slip
    (a 64-bit stride passed as a 32-bit int, a swapped pointer) would otherwise only show up as garbage on the GPU."""
    from mojo_opset_b200 import _lib

    text = open(os.path.join(ROOT, "include", "mojo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    kinds = {ctypes.c_void_p: "ptr", ctypes.c_int: "int", ctypes.c_int64: "i64", ctypes.c_float: "f32",
             ctypes.c_size_t: "size", ctypes.c_char_p: "ptr"}

    def kind_of(param: str) -> str:
        p = " ".join(param.split())
        if "*" in p:
            return "ptr"
        base = p.rsplit(" ", 1)[0] if " " in p else p
        base = base.replace("const ", "").strip()
        return {"int": "int", "int64_t": "i64", "size_t": "size", "float": "f32", "cudaStream_t": "ptr"}[base]

    protos = re.findall(r"MOJO_B200_API\s+([\w\s\*]+?)\b(mojo_b200_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(protos) == len(_lib.SIGNATURES)
    for ret, name, params in protos:
        params = params.strip()
        declared = [] if params in ("", "void") else [kind_of(x) for x in params.split(",")]
        restype, argtypes = _lib.SIGNATURES[name]
        bound = [kinds[a] if a in kinds else "ptr" for a in argtypes]
        assert declared == bound, f"{name}: header {declared} != ctypes {bound}"
        want_ret = "ptr" if "*" in ret else kind_of(ret.strip() + " x")
        assert kinds.get(restype, "ptr") == want_ret, f"{name}: return type"


def test_every_pdl_launched_kernel_waits_before_touching_memory():
    """Source lint for programmatic dependent launch (csrc/common.cuh): a kernel launched through `launch_pdl` (or with
    the programmatic-serialization attribute) may start while its predecessor is still running, so it must execute
    `pdl_wait()` (griddepcontrol.wait).  Every `__global__` kernel of a file that uses PDL must therefore either contain
    `pdl_wait()` or be launched with plain `<<<...>>>` only."""
    import re

    csrc = os.path.join(ROOT, "mojo_opset_b200", "csrc")
    checked = 0
    for fname in sorted(os.listdir(csrc)):
        if not fname.endswith(".cu"):
            continue
        src = open(os.path.join(csrc, fname)).read()
        if "launch_pdl(" not in src and "ProgrammaticStreamSerialization" not in src:
            continue
        for m in re.finditer(r"__global__\s+void\s+(?:__launch_bounds__\s*\([^)]*\)\s*)?(\w+)\s*\(", src):
            name = m.group(1)
            start = src.index("{", m.end())
            depth, i = 0, start
            while True:  # the kernel's body by brace matching
                depth += {"{": 1, "}": -1}.get(src[i], 0)
                if depth == 0:
                    break
                i += 1
            body = src[start:i]
            plain = re.search(r"\b%s\s*(<[^;]*?>)?\s*<<<" % re.escape(name), src) is not None
            via_pdl = re.search(r"launch_pdl\(\s*%s\b" % re.escape(name), src) is not None or \
                re.search(r"=\s*(\w+\s*\?\s*)?%s\s*<" % re.escape(name), src) is not None  # `auto kern = name<...>`
            if via_pdl or not plain:
                assert "pdl_wait()" in body, f"{fname}: kernel {name} is launched with PDL but never calls pdl_wait()"
                checked += 1
    assert checked >= 9, checked


def test_decode_traffic_capture_is_bound_to_the_measured_kernel(monkeypatch, tmp_path):
    """bench.py takes `roofline.traffic` from an ncu capture only if it was taken on the decode kernel it runs: the same
    sources, or - after a change elsewhere in the file - the same machine code of the measured instantiation (the SASS
    hash `mojo_opset_b200/build.py` pins at build time).  Anything else is refused."""
    import json

    sys.path.insert(0, ROOT)
    import bench
    from mojo_opset_b200 import build

    capture = os.path.join(ROOT, "profiles", "decode_traffic.json")
    if not os.path.exists(capture) or not os.path.exists(build.KERNEL_SASS_PATH):
        pytest.skip("no capture / no built kernel hash in this tree")
    with open(capture) as f:
        t = json.load(f)
    traffic, source = bench.load_decode_traffic()
    assert traffic == t["dram_bytes_per_launch"], source          # the committed capture matches the committed tree
    monkeypatch.setattr(bench, "sources_sha256", lambda paths=bench.DECODE_SOURCES: "0" * 64)
    traffic, source = bench.load_decode_traffic()                 # other sources, same SASS: accepted, and it says so
    if t.get("kernel_sass_sha256"):
        assert traffic == t["dram_bytes_per_launch"] and "SASS is identical" in source
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))             # no built hash next to the sources: refused
    os.makedirs(tmp_path / "profiles")
    with open(tmp_path / "profiles" / "decode_traffic.json", "w") as f:
        json.dump(t, f)
    traffic, source = bench.load_decode_traffic()
    assert traffic is None and "refused" in source
