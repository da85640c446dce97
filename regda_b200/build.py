"""Build the C-ABI library (libregda_b200.so) from regda_b200/csrc/*.cu for sm_100a.

In-tree output (regda_b200/lib/) so the built library travels to the GPU box with the
repo snapshot.  nvcc cross-compiles without a GPU.  The CUDA runtime is linked statically,
so the library has no dependency beyond libstdc++/libdl/libpthread and can be loaded by any
host (ctypes from Python here; cgo/JNI/... elsewhere).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(LIBDIR, "libregda_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the regda_b200 CUDA library cannot be built")
    return exe


def _stamp(src: str, headers: list[str]) -> str:
    h = hashlib.sha1()
    for p in [src] + headers:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: str, headers: list[str], verbose: bool) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    base = os.path.splitext(os.path.basename(src))[0]
    obj = os.path.join(OBJDIR, base + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src, headers)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj
    cmd = [nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
    if force and os.path.isdir(OBJDIR):
        shutil.rmtree(OBJDIR)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, headers, verbose), srcs))
    os.makedirs(LIBDIR, exist_ok=True)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", LIB] + objs + ["-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
