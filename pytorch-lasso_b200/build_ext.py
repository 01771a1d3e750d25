"""Build ``csrc/liblasso_b200.so`` in-tree with nvcc for sm_100a.

    python pytorch-lasso_b200/build_ext.py [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("LASSO_B200_LIB") or os.path.join(CSRC, "liblasso_b200.so")
SOURCES = ["cabi.cu", "fista_ffma.cu", "fista_tc.cu", "fista_res.cu", "fista_blk.cu", "aux_kernels.cu", "conv_lip.cu", "ridge.cu", "gram_tc.cu", "fista_gram.cu", "sweep_blk.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "lasso_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "128",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build liblasso_b200.so")
    return exe


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    built = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return all(os.path.getmtime(p) <= built for p in deps if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    flags = list(NVCC_FLAGS) + os.environ.get("LASSO_B200_EXTRA_FLAGS", "").split()
    if os.environ.get("LASSO_B200_BUILD_TRACE"):   # per-warp timeline instrumentation (tools/tc_trace.py)
        flags.append("-DLASSO_RES_TRACE")
    cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building liblasso_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
