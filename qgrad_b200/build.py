"""Builds libqgrad_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m qgrad_b200.build [--force]

One object per translation unit (compiled in parallel), linked into
``qgrad_b200/libqgrad_b200.so``.  The built library travels to the GPU box with the
repo snapshot; it is git-ignored.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libqgrad_b200.so")
UNITS = ["api", "expm_small", "expm_tiny", "expm_large", "state_ops", "builders", "learn", "rand", "spectral", "snap", "gemm_tc"]
HEADERS = ["common.cuh", "gemm.cuh", os.path.join("..", "..", "include", "qgrad_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(unit, force):
    src = os.path.join(CSRC, unit + ".cu")
    obj = os.path.join(OBJ, unit + ".o")
    deps = [src] + [os.path.join(CSRC, h) for h in HEADERS]
    if not force and not _stale(obj, deps):
        return unit, False, ""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(OBJ, unit + ".ptxas.log"), "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
    return unit, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda u: _compile(u, force), UNITS))
    rebuilt = [u for u, did, _ in results if did]
    objs = [os.path.join(OBJ, u + ".o") for u in UNITS]
    if rebuilt or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"rebuilt: {rebuilt or 'nothing'} -> {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
