"""In-tree build of libpimdk.so (hand-written CUDA for sm_100a + the C ABI of include/pimdk.h).

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
Every translation unit is built with -fmad=false (arithmetic in the reference's operation order; FMAs
appear only where the source says fma()), except the second copy of the CCpol kernels, which is the
opt-in "fast" mode, and the analytic-gradient kernels (opt-in mode 2: no operation-order promise).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpimdk.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"]

UNITS = [
    # (source, object, extra flags)
    ("ccpol_kernels.cu", "ccpol_strict.o", ["-fmad=false", "-DPIMDK_CCPOL_STRICT=1"]),
    ("ccpol_kernels.cu", "ccpol_fast.o", ["-fmad=true", "-DPIMDK_CCPOL_STRICT=0"]),
    ("ccpol_grad_kernels.cu", "ccpol_grad.o", ["-fmad=true"]),
    ("pes_simple.cu", "pes_simple.o", ["-fmad=false"]),
    ("watmeth_kernels.cu", "watmeth.o", ["-fmad=false"]),
    ("malon_kernels.cu", "malon.o", ["-fmad=false"]),
    ("malon_tables.cpp", "malon_tables.o", []),
    ("nm_kernels.cu", "nm_kernels.o", ["-fmad=false"]),
    ("fused_small.cu", "fused_small.o", ["-fmad=false"]),
    ("um_kernels.cu", "um_kernels.o", ["-fmad=false"]),
    ("hess_kernels.cu", "hess_kernels.o", ["-fmad=false"]),
    ("fp64_peak.cu", "fp64_peak.o", ["-fmad=false"]),
    ("pimdk_api.cu", "pimdk_api.o", ["-fmad=false"]),
    ("ccpol_tables.cpp", "ccpol_tables.o", []),
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libpimdk.so cannot be built (there is no CPU fallback)")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in ("pimdk.h", "pimdk_detmath.h")]
    headers.append(os.path.abspath(__file__))
    objs = []
    for src, obj, extra in UNITS:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, obj)
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _newer(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart=shared", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
