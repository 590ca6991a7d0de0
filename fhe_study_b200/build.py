"""Builds libfhe_b200.so (hand-written CUDA for sm_100a + the C ABI of include/fhe_b200.h) in-tree.

nvcc cross-compiles without a GPU, so this runs in the CPU container; the resulting .so is git-ignored
but travels to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# FHE_BUILD_DIR / FHE_LIB_OUT: build a tuning variant (with FHE_EXTRA_NVCC_FLAGS) beside the product library;
# FHE_B200_LIB (see _capi.py) then selects it for an A/B measurement.
OBJ = os.environ.get("FHE_BUILD_DIR") or os.path.join(HERE, "_build")
LIB = os.environ.get("FHE_LIB_OUT") or os.path.join(HERE, "libfhe_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--fmad=false",
] + os.environ.get("FHE_EXTRA_NVCC_FLAGS", "").split()  # tuning experiments only (e.g. -DFHE_A_SMEM_MINB=6)


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_stamp() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".hpp", ".h")):
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "headers.stamp")
    stamp = _deps_stamp()
    headers_changed = force or not os.path.exists(stamp_file) or open(stamp_file).read() != stamp
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if headers_changed or not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(s):
            jobs.append([NVCC, *FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else []))

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for log in ex.map(run, jobs):
                if verbose:
                    sys.stderr.write(log)
    if jobs or not os.path.exists(LIB):
        run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"])
        open(stamp_file, "w").write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
