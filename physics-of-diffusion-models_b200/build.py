"""Build libpdm_b200.so (hand-written sm_100a CUDA kernels behind the C ABI of include/pdm_b200.h).

In-tree build with nvcc so the shared library travels with the repo snapshot to the GPU box:
    python physics-of-diffusion-models_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libpdm_b200.so")

SOURCES = ["cabi.cu", "prep_kernels.cu", "noise_philox.cu", "merge.cu", "screen.cu", "stats_exact.cu", "stats_tcgen05.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
              "-DPDM_BUILD"] + ARCH


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA kernels cannot be built")
    return exe


def _deps() -> list[str]:
    inc = os.path.join(os.path.dirname(HERE), "include", "pdm_b200.h")
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [inc, os.path.abspath(__file__)]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, defines=(), out: str = "") -> str:
    """``defines`` / ``out``: dev builds of kernel variants (lib/<out>.so, loaded with PDM_B200_LIB=<path>)."""
    if not force and not is_stale() and not out:
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    exe = nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJDIR, (out + "_" if out else "") + src.replace(".cu", ".o"))
        cmd = [exe, "-c", os.path.join(CSRC, src), "-o", obj] + NVCC_FLAGS + extra + [f"-D{d}" for d in defines]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    lib = os.path.join(LIBDIR, out + ".so") if out else LIB
    tmp = lib + ".tmp"
    cmd = [exe, "-shared", "-o", tmp] + objs + ARCH + ["-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--define", action="append", default=[])
    ap.add_argument("--out", type=str, default="")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, defines=a.define, out=a.out))
