"""Build the in-tree CUDA library (libpyfem_b200.so) for sm_100a with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot, so it is built here
(nvcc cross-compiles without a GPU) and only rebuilt when the sources' content hash differs from the one stamped
beside the library (libpyfem_b200.so.srchash).
"""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libpyfem_b200.so")
SOURCES = ["pfg_setup.cu", "pfg_assemble.cu", "pfg_solve.cu", "pfg_probe.cu"]
HEADERS = ["pfg_internal.cuh", "pfg_elem.cuh"]
NVCC_FLAGS = [
    "-std=c++20", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Wno-deprecated-declarations",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libpyfem_b200.so")


STAMP_PATH = LIB_PATH + ".srchash"


def _deps():
    return [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "pyfem_b200.h")]


def _source_hash():
    """Content hash of every source the library is built from (+ the flags): file times do not survive the copy to
    the GPU box, contents do."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale():
    if not os.path.isfile(LIB_PATH) or not os.path.isfile(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link libpyfem_b200.so next to this file (no-op when the library matches the
    sources).  Serialised across processes: several ranks may import the package at once."""
    if not force and not _stale():
        return LIB_PATH
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():  # another process built it while this one waited
            return LIB_PATH
        return _build_locked(objdir, verbose)


def _build_locked(objdir, verbose):
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    tmp = LIB_PATH + ".tmp"
    link = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("link failed: " + " ".join(link))
    os.replace(tmp, LIB_PATH)
    with open(STAMP_PATH, "w") as f:
        f.write(_source_hash() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
