"""In-tree build of libsgw_b200.so (hand-written CUDA for sm_100a + the C ABI of include/sgw_b200.h).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libsgw_b200.so"
SOURCES = ["api.cu", "fft.cu", "gemm.cu", "operator.cu", "bicgstab.cu", "subspace.cu", "coulomb.cu", "invert.cu", "sigma.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-O3", "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++"]


def _newer(src: Path, dst: Path) -> bool:
    if not dst.exists():
        return True
    deps = [src] + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "sgw_b200.h"]
    return any(d.stat().st_mtime > dst.stat().st_mtime for d in deps)


def build(verbose: bool = False, force: bool = False) -> Path:
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    jobs = []
    for s in srcs:
        o = objdir / (s.stem + ".o")
        if force or _newer(s, o):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC, *FLAGS, "-Xptxas", "-v", "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (objdir / (s.stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stdout}\n{r.stderr}")
        return s.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print("compiled", name)
    objs = [objdir / (s.stem + ".o") for s in srcs]
    if force or jobs or not LIB.exists():
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
               "-o", str(LIB), *map(str, objs), "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
