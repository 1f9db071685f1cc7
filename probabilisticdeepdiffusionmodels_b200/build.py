"""Build libpddm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

``python -m probabilisticdeepdiffusionmodels_b200.build`` or ``__graft_entry__.build()``.
nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG, "libpddm_b200.so")
SOURCES = ["host_common.cu", "conv_fwd.cu", "conv_wgrad.cu", "diffusion.cu", "layout.cu", "groupnorm.cu", "groupnorm_pipe.cu",
           "attention.cu", "highprec.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(ROOT, "include", "pddm.h"))
    return hs


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdrs = _headers()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


def build_cuda_tests(force=False):
    """Standalone (torch-free) GPU test executables; named *.so so that they travel with the gpurun snapshot."""
    build_library(force=force)
    out = os.path.join(ROOT, "tests", "cuda", "test_conv_exe.so")
    src = os.path.join(ROOT, "tests", "cuda", "test_conv.cu")
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in ("host_common.cu", "conv_fwd.cu", "conv_wgrad.cu")]
    if force or _stale(out, [src] + objs):
        r = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", out, src]
                           + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building test_conv: " + r.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_cuda_tests())
