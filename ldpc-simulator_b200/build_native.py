"""Build recipe of the CUDA library behind include/ldpc_b200.h.

    python ldpc-simulator_b200/build_native.py [--force] [--verbose]

Compiles csrc/*.cu for sm_100a only (B200; no other architecture, no JIT/PTX
fallback) and links ldpc-simulator_b200/lib/libldpc_b200.so in-tree, so the
built library travels with the repository snapshot.  nvcc cross-compiles
without a GPU.  ptxas resource usage (-Xptxas -v) is kept in lib/ptxas.log.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libldpc_b200.so")
SOURCES = ["graph.cu", "spa_generic.cu", "spa_qc_resident.cu", "mc.cu", "api.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
          "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build libldpc_b200.so")


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(HERE, "..", "include", "ldpc_b200.h"))
    files.append(os.path.abspath(__file__))
    return files


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for path in sorted(_deps()):
        with open(path, "rb") as fh:
            h.update(os.path.basename(path).encode() + b"\0" + fh.read())
    return h.hexdigest()


def is_stale() -> bool:
    """Content based (snapshots do not preserve mtimes): lib/build_hash.txt records the sources."""
    stamp = os.path.join(LIBDIR, "build_hash.txt")
    if not os.path.exists(LIB) or not os.path.exists(stamp):
        return True
    with open(stamp) as fh:
        return fh.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *CFLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    logs = []
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = []
    for src, obj, r in results:
        logs.append(f"==== {src} ====\n{r.stderr}")
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(LIBDIR, "build_hash.txt"), "w") as f:
        f.write(_source_hash() + "\n")
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
