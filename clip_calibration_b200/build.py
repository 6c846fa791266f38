"""Build libccal.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m clip_calibration_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libccal.so")
STAMP = os.path.join(HERE, ".libccal.stamp")

SOURCES = ["ccal_api.cu", "bin_stats.cu", "logits_ops.cu", "knn_dac.cu", "knn_tc.cu", "score_fused.cu", "density_ratio.cu", "isotonic.cu"]
HEADERS = ["ccal_common.cuh", "sm100_ptx.cuh", "sort_scan.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--cudart", "static",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    files.append(os.path.join(os.path.dirname(HERE), "include", "ccal.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _up_to_date(digest: str) -> bool:
    if os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            return fh.read().strip() == digest
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libccal.so if it is missing or older than the sources.  Safe to call from several
    processes at once (torchrun ranks): an exclusive file lock serialises the compile."""
    import fcntl
    digest = _digest()
    if not force and _up_to_date(digest):
        return LIB
    with open(os.path.join(HERE, ".libccal.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and _up_to_date(digest):          # another process built it while we waited
            return LIB
        return _compile(digest, verbose)


def _compile(digest: str, verbose: bool) -> str:
    cmd = [_nvcc()] + NVCC_FLAGS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


C_DEMO_SRC = os.path.join(os.path.dirname(HERE), "examples", "c_abi_smoke.c")
C_DEMO_BIN = os.path.join(HERE, "c_abi_smoke")


def build_c_demo() -> str:
    """Plain-C client of the ABI (no Python / torch): examples/c_abi_smoke.c -> clip_calibration_b200/c_abi_smoke."""
    build()
    if os.path.exists(C_DEMO_BIN) and os.path.getmtime(C_DEMO_BIN) >= max(os.path.getmtime(C_DEMO_SRC), os.path.getmtime(LIB)):
        return C_DEMO_BIN
    cmd = [_nvcc(), "-O2", "-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-o", C_DEMO_BIN, C_DEMO_SRC,
           "-I", os.path.join(os.path.dirname(HERE), "include"), "-L", HERE, "-lccal",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return C_DEMO_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_c_demo())
