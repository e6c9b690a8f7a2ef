"""Build libclip_event_b200.so in-tree with nvcc for sm_100a (no torch extension machinery)."""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libclip_event_b200.so")
SOURCES = ["api_common.cu", "ot_kernels.cu", "ot_fused.cu", "ot_stream.cu", "ot_wide.cu", "dense_ce.cu", "head_step.cu", "contrastive.cu"]
HEADERS = ["ce_common.cuh", "umma_gemm.cuh", "ot_fused.cuh", os.path.join("..", "..", "include", "clip_event_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; clip_event_b200 needs the CUDA toolkit to build")
    return exe


def _stale() -> bool:
    if not os.path.exists(LIB) or os.environ.get("CE_EXTRA_NVCC_FLAGS"):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    flags += os.environ.get("CE_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DCE_GEMM_TRACE (tools/gemm_trace.py)
    cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
    log = os.path.join(OBJ, src.replace(".cu", ".log"))
    if os.path.exists(obj) and os.path.exists(log) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
        with open(log) as f:
            if f.readline().rstrip("\n") == " ".join(cmd):   # same sources AND same command line
                return obj
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (src, (r.stdout + r.stderr)[-6000:]))
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library; returns its path."""
    if not force and not _stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
