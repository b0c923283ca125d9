"""Compile csrc/*.cu into libgrandplus_b200.so IN-TREE with nvcc for sm_100a only.

    python -m grandplus_b200.build            (or: python __graft_entry__.py build)

The .so is git-ignored but travels with the repo snapshot to the GPU box.  nvcc cross-compiles
without a GPU.  There is deliberately no multi-arch list and no fallback target.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "_obj")
LIB_PATH = os.path.join(_HERE, "libgrandplus_b200.so")
SOURCES = ["capi.cu", "gfpush.cu", "gfpush_cluster.cu", "gfpush_bucket.cu", "aggregate.cu", "emb.cu"]
HEADERS = [os.path.join(CSRC, "gp_common.cuh"), os.path.join(CSRC, "gfpush_shared.cuh"), os.path.join(CSRC, "gfpush_cluster.h"), os.path.join(CSRC, "gfpush_bucket.h"), os.path.join(os.path.dirname(_HERE), "include", "grandplus_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgrandplus_b200.so cannot be built and there is no fallback")


def _stamp(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: str) -> str:
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp([os.path.join(CSRC, src)] + HEADERS)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj
    cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout[-4000:]}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s), SOURCES))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs + \
              ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout[-4000:]}")
    if verbose:
        print(f"built {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
