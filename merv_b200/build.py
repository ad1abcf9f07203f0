"""Build libmerv_fusion.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python merv_b200/build.py [--force] [-v]      (run as a script: importing the package needs the library)

The shared library is a plain C-ABI object (include/merv_fusion.h); it is git-ignored but travels with the
repo snapshot to the GPU box.  No torch headers are involved.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(REPO, "include")
# A/B builds for the labs: MERV_BUILD_DEFINES="-DX=1 ..." adds compile definitions, MERV_BUILD_TAG=name writes libmerv_fusion_<name>.so
# (load it with MERV_FUSION_LIB=<path>, merv_b200/_lib.py); the production library is the untagged one.
_TAG = os.environ.get("MERV_BUILD_TAG", "")
LIB = os.path.join(PKG, f"libmerv_fusion_{_TAG}.so" if _TAG else "libmerv_fusion.so")
OBJ_DIR = os.path.join(PKG, "csrc", f"build_{_TAG}" if _TAG else "build")
SOURCES = ["api.cu", "pool3d.cu", "mix.cu", "gemm_simt.cu", "gemm_tcgen05.cu", "backward.cu", "norm.cu", "attention.cu", "attention_tcgen05.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE, "-I", CSRC,
    *os.environ.get("MERV_BUILD_DEFINES", "").split(),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libmerv_fusion.so cannot be built (set NVCC=/path/to/nvcc)")


def _digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in sorted(os.listdir(CSRC)):
        path = os.path.join(CSRC, name)
        if os.path.isfile(path) and name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(path, "rb").read())
    h.update(open(os.path.join(INCLUDE, "merv_fusion.h"), "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = LIB + ".sha256"
    digest = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
