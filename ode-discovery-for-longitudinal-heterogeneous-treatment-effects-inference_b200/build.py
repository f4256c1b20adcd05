"""In-tree build of libb200insite.so (hand-written sm_100a CUDA behind a C ABI, include/b200i.h).

    python -m b200_insite.build          # or: python <pkg>/build.py

nvcc cross-compiles for sm_100a without a GPU.  The library is linked against the static CUDA
runtime, so it has no dependency on PyTorch; the Python host layer passes raw device pointers.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libb200insite.so")
SOURCES = ["runtime.cu", "sim_factual.cu", "theta_gram.cu", "fit_rollout.cu", "sim_cf.cu", "insite_fit.cu", "cf_eval.cu", "poly_library.cu"]  # missing files are skipped
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def _newest_dep_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for fn in os.listdir(root):
            if fn.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, fn)))
    return m


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= _newest_dep_mtime():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    dep_m = _newest_dep_mtime()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) >= dep_m:
            return obj
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
