"""Build libudt_b200.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU; the resulting .so sits next to this file so it travels to the
GPU box with the repo snapshot.  `python -m udifftext_b200.build [--force]`.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libudt_b200.so")
STAMP_PATH = os.path.join(HERE, "build", "libudt_b200.stamp")
SOURCES = ["udt_host.cu", "udt_igemm.cu", "udt_fmha.cu", "udt_norm.cu", "udt_elem.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-shared",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def _flags() -> list:
    """UDT_TRACE=1 gives a TUNING build: the role-timeline trace of udt_igemm (scripts/igemm_trace.py) and every experiment
    switch (UDT_* environment variables, udt_host.h tune_int) are compiled in.  A production build contains neither: it
    reads no environment variable and keeps debug code out of the hot loops."""
    tuning = os.environ.get("UDT_TRACE", "0") not in ("", "0")
    stamps = os.environ.get("UDT_STAMPS", "0") not in ("", "0")   # per-tile clock stamps in the FMHA kernel (scripts/fmha_timeline.py)
    return NVCC_FLAGS + (["-DUDT_IGEMM_TRACE", "-DUDT_TUNING"] if tuning else []) + (["-DUDT_FMHA_STAMPS"] if tuning and stamps else [])


def _source_digest() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/udt_api.h"]
    for name in names:
        path = os.path.normpath(os.path.join(CSRC, name))
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if sources changed; returns the path of the .so."""
    digest = _source_digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH):
        with open(STAMP_PATH) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    os.makedirs(os.path.dirname(STAMP_PATH), exist_ok=True)
    cmd = [_nvcc(), *_flags(), "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(os.path.dirname(STAMP_PATH), "nvcc.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    with open(STAMP_PATH, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print("built", path)
