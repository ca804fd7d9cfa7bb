"""Build libmermaid_b200.so in-tree with nvcc for sm_100a (B200).

    python -m mermaid_classifier_b200.build [--force] [-v]

nvcc cross-compiles without a GPU, so this runs in the CPU build container; the resulting
``mermaid_classifier_b200/libmermaid_b200.so`` travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libmermaid_b200.so"
STAMP = PKG / ".libmermaid_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # No --use_fast_math: the head / Platt / Adam kernels need IEEE division, sqrt and accurate expf / logf (the
    # reference's 1e-6 export gate).  Backbone kernels name their fast intrinsics (__expf, __fdividef, tanh.approx)
    # explicitly.  Denormals are flushed (irrelevant at the tolerances in play, and it keeps MUFU sequences short).
    "-ftz=true",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _extra_defs() -> list[str]:
    """Extra -D flags for diagnosis builds, e.g. MC_NVCC_DEFS="-DMC_TC_TIMING=1"."""
    return os.environ.get("MC_NVCC_DEFS", "").split()


def _sources_digest() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*")) + list(INCLUDE.glob("*.h")))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS + _extra_defs()).encode())
    return h.hexdigest()


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libmermaid_b200 cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    return not (LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == _sources_digest())


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _sources_digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, *_extra_defs(), "-I", str(INCLUDE), "-o", str(LIB), str(CSRC / "api.cu"), "-lcuda", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout + proc.stderr)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
