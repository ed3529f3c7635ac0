"""In-tree build of the C-ABI library: nvcc, sm_100a only, -lineinfo.

    python -m isp_tts_b200.build          # or __graft_entry__.build()

The result, isp-tts_b200/libisp_tts_b200.so, is git-ignored but travels with
the working tree.  nvcc cross-compiles, so this runs on a box without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
SO_PATH = os.path.join(PKG_DIR, "libisp_tts_b200.so")
SOURCES = ["isp_capi.cu", "isp_mas.cu", "isp_mas2.cu", "isp_mas_wide.cu", "isp_mas_cluster.cu", "isp_loglik.cu", "isp_loglik_bwd.cu", "isp_loglik_wide.cu", "isp_consumers.cu", "isp_stage.cu", "isp_ctc.cu", "isp_gemm.cu", "isp_stacks.cu"]
HEADERS = ["common.cuh", "isp_internal.h", "isp_mas_ptx.cuh", "isp_tc05.cuh", os.path.join("..", "..", "include", "isp_tts_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-Xptxas", "-v",
    "--ftz=false", "--prec-div=true", "--prec-sqrt=true", "--fmad=true",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or install the CUDA toolkit); the library cannot be built")


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libisp_tts_b200.so.  Returns its path."""
    if not force and not _stale():
        return SO_PATH
    nvcc = _nvcc()
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stdout + r.stderr))
        objs.append(obj)
    # the CUDA runtime is linked dynamically (torch has already loaded libcudart.so.12 into the process; the
    # rpath covers a bare ctypes load): the static runtime would embed every runtime entry point in the library
    cmd = [nvcc, "-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
           "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
