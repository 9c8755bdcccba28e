"""Build ``libmojo_b200.so`` (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

``python -m mojo_opset_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
Objects go to ``build/obj`` (git- and gpurun-ignored); the shared library is written next to this file so
it travels to the GPU box with the repo snapshot.
"""

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB_PATH = os.path.join(PKG_DIR, "libmojo_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-DMOJO_B200_BUILD",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the b200 backend cannot be built (there is no CPU fallback)")
    return nvcc


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "mojo_b200.h"))
    return hdrs


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = _headers()
    jobs = []
    objects = []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objects.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([nvcc, *NVCC_FLAGS, *os.environ.get("MOJO_B200_EXTRA_NVCC_FLAGS", "").split(), "-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
        return res

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(run, jobs))
    if jobs or force or _stale(LIB_PATH, objects):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objects,
             "-cudart", "static", "-Xlinker", "--exclude-libs,ALL"])
    _write_kernel_sass()
    return LIB_PATH


# The headline decode kernel (cfg2: bf16, head_dim 128, split-halves layout, single split).  bench.py accepts an ncu
# DRAM-traffic capture (profiles/decode_traffic.json) only for the kernel it runs: same sources, or - when a source file
# changed without changing this kernel - the same machine code, which is what this hash pins.
DECODE_KERNEL_SYMBOL = "_ZN4mojo23paged_decode_mma_kernelI13__nv_bfloat16Li128ELb1ELb0EEEv14CUtensorMap_stS2_NS_12DecodeParamsE"
KERNEL_SASS_PATH = os.path.join(PKG_DIR, "kernel_sass.json")


def kernel_sass_sha256(obj: str, symbol: str):
    """sha256 over the instruction lines (address, mnemonic, operands, encoding) of one kernel in an object file."""
    import hashlib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(obj):
        return None
    res = subprocess.run([cuobjdump, "-sass", "-fun", symbol, obj], capture_output=True, text=True)
    lines = [ln.strip() for ln in res.stdout.splitlines() if ln.lstrip().startswith("/*")]
    if res.returncode != 0 or len(lines) < 100:
        return None
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


def _write_kernel_sass():
    import json

    obj = os.path.join(OBJ_DIR, "paged_decode.o")
    if os.path.exists(KERNEL_SASS_PATH) and os.path.exists(obj) and os.path.getmtime(KERNEL_SASS_PATH) >= os.path.getmtime(obj):
        return
    digest = kernel_sass_sha256(obj, DECODE_KERNEL_SYMBOL)
    if digest:
        with open(KERNEL_SASS_PATH, "w") as f:
            json.dump({"paged_decode_mma_kernel<bf16, 128, split-halves, single split>":
                       {"symbol": DECODE_KERNEL_SYMBOL, "sass_sha256": digest}}, f, indent=1)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
