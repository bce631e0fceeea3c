"""Build libgccvae.so (sm_100a only) in-tree with nvcc.  No torch extension machinery: the
library is a plain C-ABI shared object (include/gccvae.h) loaded with ctypes."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
# GCCVAE_TIMELINE=1 builds the instrumented variant (pipeline timeline hooks of scripts/timeline_probe.py compiled in)
TIMELINE = os.environ.get("GCCVAE_TIMELINE", "0") not in ("", "0")
LIB = os.path.join(CSRC, "libgccvae_tl.so" if TIMELINE else "libgccvae.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", INCLUDE, "-I", CSRC,
] + (["-DGCCVAE_TIMELINE"] if TIMELINE else [])


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgccvae.so cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return sorted(deps + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)])


def source_hash() -> str:
    """sha256 over the names and contents of every source / header and the compiler flags: the library is rebuilt when
    this changes, whatever the file times say (the .so travels to the GPU box with the snapshot, file times do not
    survive that reliably)."""
    h = hashlib.sha256(" ".join(f for f in NVCC_FLAGS if not f.startswith("/")).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stamp_path() -> str:
    return LIB + ".sha256"


def _stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(_stamp_path()):
        return True
    with open(_stamp_path()) as fh:
        return fh.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build_tl" if TIMELINE else "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for {}:\n{}\n{}".format(src, r.stdout, r.stderr))
        if verbose and r.stderr:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    # -cudart shared: the runtime is the process's libcudart (torch has loaded one already), not a second static copy
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", "-Xcompiler", "-fPIC",
           "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n{}\n{}".format(r.stdout, r.stderr))
    with open(_stamp_path(), "w") as fh:
        fh.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
