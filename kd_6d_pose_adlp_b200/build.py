"""In-tree build of ``libkdot.so`` (CUDA kernels + C ABI) for sm_100a.

``python -m kd_6d_pose_adlp_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU; the ``.so`` lands in ``kd_6d_pose_adlp_b200/lib/`` (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libkdot.so")
# test-only A/B build of the same ABI: the streaming kernel evaluates every column tile (no exact tile skipping)
NOSKIP_LIB_PATH = os.path.join(LIB_DIR, "libkdot_noskip.so")
STAMP = os.path.join(LIB_DIR, "libkdot.stamp")
SOURCES = ["kdot_api.cu", "kdot_small.cu", "kdot_tiled.cu", "kdot_stream.cu", "kdot_mmd.cu", "kdot_select.cu", "kdot_decode.cu", "kdot_losses.cu", "kdot_targets.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libkdot.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(PKG_DIR, "..", "include", "kdot.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources changed since the last build; returns the library path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(NOSKIP_LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB_PATH
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    jobs = [subprocess.Popen([_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *srcs], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True),
            subprocess.Popen([_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], "-DKDOT_NO_TILE_SKIP", "-o", NOSKIP_LIB_PATH, *srcs],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)]
    outs = [j.communicate()[0] for j in jobs]
    if any(j.returncode != 0 for j in jobs):
        raise RuntimeError("nvcc failed:\n" + "\n".join(outs))
    res = type("R", (), {"stdout": outs[0], "stderr": ""})
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as fh:
        fh.write(outs[0])
    with open(STAMP, "w") as fh:
        fh.write(dig)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
