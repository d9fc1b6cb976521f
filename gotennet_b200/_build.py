"""In-tree build of the C-ABI shared library (nvcc, sm_100a only).

    python -m gotennet_b200._build        # or __graft_entry__.build()

The .so lands in gotennet_b200/lib/ (git-ignored, but it travels to the GPU box
with the gpurun snapshot).  There is no JIT and no fallback: if the library is
missing the package raises at first use.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libgotennet_b200.so"
SOURCES = ["graph.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_tc16.cu", "init.cu", "gata.cu", "gata_staged.cu", "edge_fused.cu", "htr.cu", "eqff.cu", "readout.cu", "heads.cu", "optim.cu", "norm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def lib_path() -> str:
    # GOTEN_LIB_PATH: load another build of the same ABI (A/B timing of two kernel versions on one box)
    return os.environ.get("GOTEN_LIB_PATH") or os.path.join(LIBDIR, LIBNAME)


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libgotennet_b200.so")
    return cand


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    digest = _digest()
    if not force and os.path.exists(lib_path()) and os.path.exists(stamp):
        if open(stamp).read().strip() == digest:
            return lib_path()
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib_path(), *objs, "-lcudart", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"[gotennet_b200] built {lib_path()}", file=sys.stderr)
    return lib_path()


if __name__ == "__main__":
    build(force="--force" in sys.argv)
