"""Builds libtfhe_ntt_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

Usage: python build.py [--force] [--nvtx]
The shared object is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtfhe_ntt_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

SOURCES = ["ntt_engine.cu", "ntt_fast_solinas.cu", "ntt_fast_shoup64.cu", "ntt_fast_shoup32.cu",
           "ntt_fast_exact.cu", "ntt_pbs_solinas.cu", "capi_prime.cu", "capi_native.cu", "capi_product.cu", "capi_pbs.cu",
           "capi_custum_radix.cu"]
NVCC_FLAGS = [*(os.environ.get("NTT_B200_EXTRA_NVCC", "").split()),
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "tfhe_ntt_b200.h"), os.path.abspath(__file__)]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, nvtx=False):
    """nvtx=True (python build.py --nvtx): NVTX ranges around the host entry points (-DNTT_B200_NVTX)."""
    deps = _deps()
    flags = NVCC_FLAGS + (["-DNTT_B200_NVTX"] if nvtx else [])
    force = force or nvtx
    if not force and not _stale(OUT, deps):
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(obj, deps):
            cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-lcudart"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, nvtx="--nvtx" in sys.argv))
