#!/usr/bin/env python
"""Builds the reference's own KNN kernel into oracle/_ref/libknn_cuda_ref.so (test infrastructure only).

The one native CUDA dependency of the reference is knn_cuda (datasets/moco_flow_dataset.py:8,76,121:
``KNN(k=1, transpose_mode=True)``).  Its source ships INSIDE the reference tree, in the wheel
``/root/reference/docker/KNN_CUDA-0.2-py3-none-any.whl`` (``knn_cuda/csrc/cuda/knn.cu``, a modified kNN-CUDA of
V. Garcia: cuComputeDistanceGlobal -> cuInsertionSort -> cuParallelSqrt).  ``knn.cu`` has no torch dependency, so this
script extracts it to a temporary directory (nothing of it is copied into this repository), wraps its
``knn_device(ref, ref_nb, query, query_nb, dim, k, dist, ind, stream)`` in one ``extern "C"`` entry and cross-compiles
it with nvcc for sm_100a, with nvcc's defaults (-fmad=true), exactly as ``knn_cuda/__init__.py`` JIT-builds it.
The .so is git-ignored and travels to the GPU box with the repo snapshot (/root/reference does not exist there).

    python oracle/build_knn_ref.py        # no-op when /root/reference is absent
"""
import os
import subprocess
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
WHEEL = "/root/reference/docker/KNN_CUDA-0.2-py3-none-any.whl"
OUT = os.path.join(HERE, "_ref", "libknn_cuda_ref.so")

WRAPPER = r'''
#include "knn.cu"
// dist: [ref_nb][query_nb] scratch (row 0 = distance to the nearest neighbour after the call), ind: [k][query_nb], 1-based
extern "C" int knn_ref(float* ref_dev, int ref_nb, float* query_dev, int query_nb, int dim, int k, float* dist_dev,
                       long* ind_dev, cudaStream_t stream) {
  knn_device(ref_dev, ref_nb, query_dev, query_nb, dim, k, dist_dev, ind_dev, stream);
  return (int)cudaGetLastError();
}
'''


def build(verbose: bool = False) -> str:
    if not os.path.exists(WHEEL):
        return OUT if os.path.exists(OUT) else ""
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(WHEEL) and \
            os.path.getmtime(OUT) >= os.path.getmtime(__file__):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        with zipfile.ZipFile(WHEEL) as z:
            z.extract("knn_cuda/csrc/cuda/knn.cu", tmp)
        src_dir = os.path.join(tmp, "knn_cuda", "csrc", "cuda")
        with open(os.path.join(src_dir, "wrap.cu"), "w") as fh:
            fh.write(WRAPPER)
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
               os.path.join(src_dir, "wrap.cu"), "-o", OUT]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for the reference knn.cu")
    return OUT


if __name__ == "__main__":
    print(build(verbose=True) or "reference wheel not present; nothing built")
