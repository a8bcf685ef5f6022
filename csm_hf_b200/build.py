"""Build libcsm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcsm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-t", "0",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(
        os.path.join(CSRC, "*.inl")) + [
        os.path.join(os.path.dirname(HERE), "include", "csm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Build the library.  `defines` / `out`: experiment variants (tools/gpu_variants.sh), loaded with CSM_LIB=path."""
    out = out or LIB
    if out == LIB and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + sources()   # cudart only: no cuBLAS, no libcuda
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libcsm_b200.so")
    if out == LIB:
        with open(os.path.join(HERE, "build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
