"""In-tree build of the sm_100a kernel library.

`nvcc` cross-compiles for Blackwell without a GPU, so this runs on the CPU build box; the resulting
``_olsb200_kernels.so`` sits next to this file (git-ignored, shipped to the GPU box by ``gpurun``) and is
loaded with ``ctypes`` (no torch C++ headers => a full rebuild takes seconds).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "_olsb200_kernels.so"
STAMP = HERE / "_olsb200_kernels.stamp"
SOURCES = ["api.cu", "api_nn.cu"]
HEADERS = ["tc05.cuh", "conv_tc.cuh", "fl_kernels.cuh", "linear_tc.cuh", "bgemm_tc.cuh", "nn_kernels.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "--threads", "2",
    "-diag-suppress", "177",
]


def _nvcc() -> str | None:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else None


def source_hash() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library if sources changed.  Raises if nvcc is missing or compilation fails."""
    if not force and is_fresh():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        if LIB.exists():
            return LIB  # stale-but-present beats nothing on a box without a toolkit
        raise RuntimeError("nvcc not found and no prebuilt _olsb200_kernels.so present")
    tmp = LIB.with_suffix(".so.tmp%d" % os.getpid())          # link into a scratch name, publish atomically
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(tmp)] + [str(CSRC / s) for s in SOURCES] + ["-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB)
    STAMP.write_text(source_hash())
    if verbose:
        print(res.stderr)
    return LIB


def sass_summary() -> dict:
    """Count the SASS mnemonics that prove the tcgen05 / TMA path (see profiles/)."""
    import re
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    pats = {"UTCHMMA": r"UTCHMMA", "UTMALDG": r"UTMALDG", "LDTM": r"LDTM", "UTCBAR": r"UTCBAR",
            "HMMA_legacy": r"\bHMMA"}
    return {k: len(re.findall(v, out)) for k, v in pats.items()}


if __name__ == "__main__":
    import sys
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p, sass_summary())
