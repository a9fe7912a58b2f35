"""In-tree build of the sm_100a extension (``zero_transformer_b200/_ext/_ztb_cuda*.so``).

``build()`` cross-compiles without a GPU (nvcc 12.9, ``-gencode arch=compute_100a,code=sm_100a``);
the resulting ``.so`` is git-ignored but travels to the GPU box with the tree.  ``load()`` only
imports an already-built ``.so`` -- nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_ROOT, "csrc")
EXT_DIR = os.path.join(_PKG, "_ext")
NAME = "_ztb_cuda"
SOURCES = ["bindings.cpp", "host_util.cu", "elementwise.cu", "optim.cu", "gemm_tcgen05.cu", "gemm2_tcgen05.cu", "attention.cu", "comm.cu"]
CUDA_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]

_mod = None


def _so_path():
    hits = sorted(glob.glob(os.path.join(EXT_DIR, NAME + "*.so")))
    return hits[-1] if hits else None


def build(verbose: bool = False):
    from torch.utils import cpp_extension

    os.makedirs(EXT_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    mod = cpp_extension.load(
        name=NAME, sources=srcs, extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=CUDA_FLAGS,
        extra_include_paths=[CSRC], build_directory=EXT_DIR, verbose=verbose, with_cuda=True,
    )
    global _mod
    _mod = mod
    return mod


def load():
    """Import the prebuilt extension; raise loudly if it is missing (no eager fallback on GPUs)."""
    global _mod
    if _mod is not None:
        return _mod
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    so = _so_path()
    if so is None:
        raise RuntimeError(
            f"{NAME}.so not found under {EXT_DIR}: the sm_100a kernels are not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` from the repo root.")
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[NAME] = mod
    _mod = mod
    return mod


def is_built() -> bool:
    return _so_path() is not None
