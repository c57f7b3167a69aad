"""Build libpdephi_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m pde_fluid_phi_b200.build [--force]

The library is a plain C-ABI shared object (include/pdephi_b200.h); it links the CUDA runtime
statically and needs nothing from torch.  The built .so is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libpdephi_b200.so")
SOURCES = ["plan.cu", "api.cu", "transforms_simt.cu", "transforms_tc.cu", "transforms_tc3.cu", "mix.cu", "pointwise.cu", "pointwise_wide.cu", "block_tc.cu", "xstage_tc.cu", "transforms_gen.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "pdephi_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """variant / defines: an experiment build, lib/libpdephi_b200.<variant>.so compiled with the extra -D flags (loaded with
    PDEPHI_LIB=<path>); the product is the plain build."""
    lib = LIB if not variant else os.path.join(LIBDIR, f"libpdephi_b200.{variant}.so")
    if not variant and not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(PKG, "build" if not variant else f"build_{variant}")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant,
                        defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
