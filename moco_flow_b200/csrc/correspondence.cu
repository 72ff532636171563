// Nearest SMPL vertex + per-vertex rigid transform of query points: the KNN(k=1) correspondence
// sampling that feeds the flow networks' supervision, datasets/moco_flow_dataset.py:121-130 (knn_cuda's
// KNN(k=1, transpose_mode=True) over ~6.9 K vertices, then trans[ind] @ [x, 1]).  Brute force with the vertices staged
// through shared memory; one thread per query.  The squared distance is formed exactly as knn_cuda's
// cuComputeDistanceGlobal does (knn_cuda/csrc/cuda/knn.cu inside docker/KNN_CUDA-0.2-py3-none-any.whl: tmp = ref - query,
// ssd += tmp*tmp over x, y, z, contracted to FMAs by nvcc's default -fmad=true): d2 = fma(dz,dz, fma(dy,dy, dx*dx));
// cuInsertionSort with k = 1 keeps the FIRST strict minimum; cuParallelSqrt takes sqrtf.  Indices and distances are
// bit-identical to the reference kernel (tests/test_correspondence.py against oracle/_ref/libknn_cuda_ref.so).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/moco_flow_b200.h"

namespace mcf {

constexpr int kKnnThreads = 256;
constexpr int kKnnTile = 1024;

__global__ void __launch_bounds__(kKnnThreads) k_nearest_vertex(const float* __restrict__ verts, int n_verts,
                                                                const float* __restrict__ trans,
                                                                const float* __restrict__ query, long long n_query,
                                                                float thickness, float* __restrict__ dist,
                                                                long long* __restrict__ ind, float* __restrict__ cano,
                                                                unsigned char* __restrict__ inside) {
  __shared__ float sx[kKnnTile], sy[kKnnTile], sz[kKnnTile];
  const long long q = blockIdx.x * (long long)kKnnThreads + threadIdx.x;
  const bool live = q < n_query;
  float x = 0.f, y = 0.f, z = 0.f;
  if (live) { x = query[q * 3 + 0]; y = query[q * 3 + 1]; z = query[q * 3 + 2]; }
  float best = 3.402823466e38f;
  int best_i = 0;
  for (int v0 = 0; v0 < n_verts; v0 += kKnnTile) {
    const int n = min(kKnnTile, n_verts - v0);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kKnnThreads) {
      sx[i] = verts[(long long)(v0 + i) * 3 + 0];
      sy[i] = verts[(long long)(v0 + i) * 3 + 1];
      sz[i] = verts[(long long)(v0 + i) * 3 + 2];
    }
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int i = 0; i < n; ++i) {
        const float dx = sx[i] - x, dy = sy[i] - y, dz = sz[i] - z;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d2 < best) { best = d2; best_i = v0 + i; }
      }
    }
  }
  if (!live) return;
  const float d = sqrtf(best);
  if (dist) dist[q] = d;
  if (ind) ind[q] = best_i;
  if (inside) inside[q] = d < thickness ? 1 : 0;
  if (cano && trans) {
    const float* T = trans + (long long)best_i * 16;   // row-major 4x4; rows 0..2 applied to [x y z 1]
#pragma unroll
    for (int r = 0; r < 3; ++r) cano[q * 3 + r] = ((T[r * 4 + 0] * x + T[r * 4 + 1] * y) + T[r * 4 + 2] * z) + T[r * 4 + 3];
  }
}

}  // namespace mcf

extern "C" int mcf_nearest_vertex(const float* verts, int n_verts, const float* trans, const float* query,
                                  long long n_query, float thickness, float* dist, long long* ind, float* cano,
                                  unsigned char* inside, cudaStream_t stream) {
  if (n_verts <= 0 || n_query < 0 || !verts || (n_query > 0 && !query)) return MCF_ERR_BAD_ARG;
  if (n_query == 0) return 0;
  mcf::k_nearest_vertex<<<(unsigned)((n_query + mcf::kKnnThreads - 1) / mcf::kKnnThreads), mcf::kKnnThreads, 0, stream>>>(
      verts, n_verts, trans, query, n_query, thickness, dist, ind, cano, inside);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
