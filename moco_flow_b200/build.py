"""In-tree build of the C-ABI CUDA library (sm_100a only) and of nothing else.

    python -m moco_flow_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting ``csrc/libmoco_flow_b200.so`` is git-ignored but
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libmoco_flow_b200.so")
STAMP = os.path.join(CSRC, ".build_stamp")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# (source, extra flags).  render_ops.cu keeps fp32 mul/add unfused to mirror the reference's op sequence.
UNITS = [
    ("render_ops.cu", ["-fmad=false"]),
    ("chain.cu", []),
    ("nof_chain.cu", []),
    ("gemm_dw.cu", []),
    ("optim.cu", []),
    ("camera.cu", ["-fmad=false"]),
    ("correspondence.cu", ["-fmad=false"]),
    ("plan_host.cu", []),
]
HEADERS = [os.path.join(CSRC, "ptx.cuh"), os.path.join(CSRC, "nof_math.cuh"),
           os.path.join(ROOT, "include", "moco_flow_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for f in [os.path.join(CSRC, u) for u, _ in UNITS] + HEADERS + [os.path.abspath(__file__)]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    for src, extra in UNITS:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}")
        objs.append(obj)
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
