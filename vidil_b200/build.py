"""Build vidil_b200/_C/libvidil_b200.so from vidil_b200/csrc/*.cu with nvcc for sm_100a.

The library is a plain C-ABI shared object (include/vidil_b200.h): no torch headers, no libcuda link
(the driver's tensor-map encoder is resolved at run time through cudart), so it loads on a machine
without a GPU and is built in-tree so that it travels with the repository snapshot.

    python -m vidil_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import argparse
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(PKG_DIR, "csrc")
OUT_DIR = os.path.join(PKG_DIR, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libvidil_b200.so")
SOURCES = ["api.cu", "gemm.cu", "layernorm.cu", "attention.cu", "attention_tc.cu", "attention_tc257.cu", "attention_tcl.cu", "elementwise.cu", "topk.cu", "preprocess.cu", "med.cu", "med_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; vidil_b200 needs the CUDA 12.9 toolkit to build its kernels")
    return exe


def _source_digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(SRC_DIR)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(SRC_DIR, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    with open(os.path.join(PKG_DIR, "..", "include", "vidil_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = os.path.join(OUT_DIR, "build.stamp")
    if not (os.path.exists(LIB_PATH) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of the shared library.

    Safe under torchrun: every rank may call this at import.  The whole build runs under an exclusive file lock, objects
    go to a private temporary directory and the finished library is moved into place with os.replace, so no process can
    ever dlopen a half-written .so or link half-written objects; ranks that lose the race find the stamp current."""
    if not force and is_current():
        return LIB_PATH
    import fcntl
    import tempfile
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():  # another process built it while this one waited for the lock
                return LIB_PATH
            nvcc = _nvcc()
            extra = ["-Xptxas", "-v"] if verbose else []
            with tempfile.TemporaryDirectory(dir=OUT_DIR, prefix=".build-") as tmp:

                def compile_one(src: str) -> str:
                    obj = os.path.join(tmp, src.replace(".cu", ".o"))
                    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(SRC_DIR, src), "-o", obj]
                    r = subprocess.run(cmd, capture_output=True, text=True)
                    if r.returncode != 0:
                        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
                    if verbose:
                        sys.stderr.write(f"== {src}\n{r.stderr}\n")
                    return obj

                with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
                    objs = list(ex.map(compile_one, SOURCES))
                tmp_lib = os.path.join(tmp, "libvidil_b200.so")
                cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp_lib, *objs]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
                os.replace(tmp_lib, LIB_PATH)
            tmp_stamp = os.path.join(OUT_DIR, f".build.stamp.{os.getpid()}")
            with open(tmp_stamp, "w") as f:
                f.write(_source_digest())
            os.replace(tmp_stamp, os.path.join(OUT_DIR, "build.stamp"))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
