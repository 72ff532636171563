// Ray generation and canvas scatter on the device: the step on the input side of render_rays and the
// image assembly on its output side.  Reference semantics: utils/camera.py:29-82 (gen_ray_directions, gen_rays),
// :134-148 (Camera.make_rays) and trainer/trainer_moco_flow.py:226-268 (render: masked gather before, canvas
// scatter after).  fp32, one thread per ray / pixel; compiled with -fmad=false so the arithmetic is the
// reference's op sequence (divide, three products summed left to right, normalise by a sqrt).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/moco_flow_b200.h"

namespace mcf {

struct Cam {
  float c2w[12];  // row-major [3][4]
};

__global__ void k_make_rays(int H, int W, float fx, float cx, float cy, Cam cam, float near, float far, float idx,
                            const long long* __restrict__ pix, long long n, int has_c2w, float* __restrict__ rays,
                            int ray_stride) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long long p = pix ? pix[k] : k;
  // pixel p = row j, column i (utils/camera.py:41-43: meshgrid 'ij' then transposed)
  const float i = (float)(p % W), j = (float)(p / W);
  const float dx = (i - cx) / fx, dy = -(j - cy) / fx, dz = -1.0f;   // both axes use focal[0] (utils/camera.py:49)
  float rx = dx, ry = dy, rz = dz, ox = 0.f, oy = 0.f, oz = 0.f;
  if (has_c2w) {  // directions @ c2w[:, :3].T
    rx = dx * cam.c2w[0] + dy * cam.c2w[1] + dz * cam.c2w[2];
    ry = dx * cam.c2w[4] + dy * cam.c2w[5] + dz * cam.c2w[6];
    rz = dx * cam.c2w[8] + dy * cam.c2w[9] + dz * cam.c2w[10];
    ox = cam.c2w[3]; oy = cam.c2w[7]; oz = cam.c2w[11];
  }
  const float nrm = sqrtf(rx * rx + ry * ry + rz * rz);
  float* r = rays + k * ray_stride;
  r[0] = ox; r[1] = oy; r[2] = oz;
  r[3] = rx / nrm; r[4] = ry / nrm; r[5] = rz / nrm;
  r[6] = near; r[7] = far; r[8] = idx;
}

// trainer_moco_flow.py:247-262: canvas = background, depth = 10; masked pixels get depth 8; masked pixels whose
// opacity is > 0 get the rendered colour and depth
__global__ void k_canvas_init(const float* __restrict__ background, long long n_pix, float* __restrict__ img,
                              float* __restrict__ depth) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  img[p * 3 + 0] = background[p * 3 + 0];
  img[p * 3 + 1] = background[p * 3 + 1];
  img[p * 3 + 2] = background[p * 3 + 2];
  depth[p] = 10.0f;
}

__global__ void k_canvas_scatter(const long long* __restrict__ pix, long long n, const float* __restrict__ rgb,
                                 const float* __restrict__ dep, const float* __restrict__ opacity,
                                 float* __restrict__ img, float* __restrict__ depth) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long long p = pix ? pix[k] : k;
  if (opacity[k] > 0.f) {
    img[p * 3 + 0] = rgb[k * 3 + 0];
    img[p * 3 + 1] = rgb[k * 3 + 1];
    img[p * 3 + 2] = rgb[k * 3 + 2];
    depth[p] = dep[k];
  } else {
    depth[p] = 8.0f;
  }
}

}  // namespace mcf

extern "C" int mcf_make_rays(int H, int W, float focal, float cx, float cy, const float* c2w_host, float near, float far,
                             float img_ind, const long long* pixel_index, long long n_rays, float* rays,
                             int ray_stride, cudaStream_t stream) {
  if (H <= 0 || W <= 0 || n_rays < 0 || ray_stride < 9 || !(focal != 0.f)) return MCF_ERR_BAD_ARG;
  if (n_rays == 0) return 0;
  if (!pixel_index && n_rays > (long long)H * W) return MCF_ERR_BAD_ARG;
  mcf::Cam cam;
  for (int i = 0; i < 12; ++i) cam.c2w[i] = c2w_host ? c2w_host[i] : 0.f;
  mcf::k_make_rays<<<(unsigned)((n_rays + 255) / 256), 256, 0, stream>>>(H, W, focal, cx, cy, cam, near, far, img_ind,
                                                                        pixel_index, n_rays, c2w_host ? 1 : 0, rays,
                                                                        ray_stride);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int mcf_canvas_scatter(const float* background, long long n_pixels, const long long* pixel_index,
                                  long long n_rays, const float* rgb, const float* depth, const float* opacity,
                                  float* img_out, float* depth_out, cudaStream_t stream) {
  if (n_pixels < 0 || n_rays < 0 || (!pixel_index && n_rays > n_pixels)) return MCF_ERR_BAD_ARG;
  if (n_pixels == 0) return 0;
  mcf::k_canvas_init<<<(unsigned)((n_pixels + 255) / 256), 256, 0, stream>>>(background, n_pixels, img_out, depth_out);
  if (n_rays > 0)
    mcf::k_canvas_scatter<<<(unsigned)((n_rays + 255) / 256), 256, 0, stream>>>(pixel_index, n_rays, rgb, depth, opacity,
                                                                               img_out, depth_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
