// fp32 head arithmetic of the NoF flow network, shared by the chain kernels.
#pragma once
#include <cuda_runtime.h>

namespace mcf {

// quaternion head of NoF (models/nof.py:75-80 with kornia 0.6.5 semantics)
__device__ __forceinline__ void nof_quat_apply(const float* h9, const float* x, float* out) {
  float v0 = h9[0], v1 = h9[1], v2 = h9[2];
  float n = fmaxf(sqrtf(v0 * v0 + v1 * v1 + v2 * v2), 1e-8f);
  float sn, cs;
  sincosf(n, &sn, &cs);
  float a = sn / n;
  float q0 = v0 * a, q1 = v1 * a, q2 = v2 * a, q3 = cs;
  float inv = 1.0f / fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);
  float qx = q0 * inv, qy = q1 * inv, qz = q2 * inv, qw = q3 * inv;
  float tx = 2.f * qx, ty = 2.f * qy, tz = 2.f * qz;
  float twx = tx * qw, twy = ty * qw, twz = tz * qw;
  float txx = tx * qx, txy = ty * qx, txz = tz * qx;
  float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  float r00 = 1.f - (tyy + tzz), r01 = txy - twz, r02 = txz + twy;
  float r10 = txy + twz, r11 = 1.f - (txx + tzz), r12 = tyz - twx;
  float r20 = txz - twy, r21 = tyz + twx, r22 = 1.f - (txx + tyy);
  float y0 = x[0] - h9[3], y1 = x[1] - h9[4], y2 = x[2] - h9[5];
  out[0] = y0 * r00 + y1 * r10 + y2 * r20 + h9[3] + h9[6];
  out[1] = y0 * r01 + y1 * r11 + y2 * r21 + h9[4] + h9[7];
  out[2] = y0 * r02 + y1 * r12 + y2 * r22 + h9[5] + h9[8];
}

// backward of the quaternion head: g = dL/d out (3).  d9 = dL/d{v,s,t}; dxin = dL/dx.
__device__ __forceinline__ void nof_quat_backward(const float* h9, const float* x, const float* g, float* d9,
                                                  float* dxin) {
  float v0 = h9[0], v1 = h9[1], v2 = h9[2];
  float nraw = sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
  bool clamped = nraw < 1e-8f;
  float n = fmaxf(nraw, 1e-8f);
  float sn, cs;
  sincosf(n, &sn, &cs);
  float a = sn / n;
  float q[4] = {v0 * a, v1 * a, v2 * a, cs};
  float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
  float inv = 1.0f / qn;
  float qx = q[0] * inv, qy = q[1] * inv, qz = q[2] * inv, qw = q[3] * inv;
  float r00 = 1.f - 2.f * (qy * qy + qz * qz), r01 = 2.f * (qx * qy - qw * qz), r02 = 2.f * (qx * qz + qw * qy);
  float r10 = 2.f * (qx * qy + qw * qz), r11 = 1.f - 2.f * (qx * qx + qz * qz), r12 = 2.f * (qy * qz - qw * qx);
  float r20 = 2.f * (qx * qz - qw * qy), r21 = 2.f * (qy * qz + qw * qx), r22 = 1.f - 2.f * (qx * qx + qy * qy);
  float y0 = x[0] - h9[3], y1 = x[1] - h9[4], y2 = x[2] - h9[5];
  // out_j = sum_i y_i R_ij + s_j + t_j
  float dy0 = r00 * g[0] + r01 * g[1] + r02 * g[2];
  float dy1 = r10 * g[0] + r11 * g[1] + r12 * g[2];
  float dy2 = r20 * g[0] + r21 * g[1] + r22 * g[2];
  dxin[0] = dy0; dxin[1] = dy1; dxin[2] = dy2;
  d9[3] = g[0] - dy0; d9[4] = g[1] - dy1; d9[5] = g[2] - dy2;  // d s
  d9[6] = g[0]; d9[7] = g[1]; d9[8] = g[2];                    // d t
  // dR_ij = y_i g_j
  float d00 = y0 * g[0], d01 = y0 * g[1], d02 = y0 * g[2];
  float d10 = y1 * g[0], d11 = y1 * g[1], d12 = y1 * g[2];
  float d20 = y2 * g[0], d21 = y2 * g[1], d22 = y2 * g[2];
  // gradient w.r.t. the unit quaternion (x,y,z,w)
  float gx = 2.f * (qy * (d01 + d10) + qz * (d02 + d20) - 2.f * qx * (d11 + d22) + qw * (d21 - d12));
  float gy = 2.f * (qx * (d01 + d10) + qz * (d12 + d21) - 2.f * qy * (d00 + d22) + qw * (d02 - d20));
  float gz = 2.f * (qx * (d02 + d20) + qy * (d12 + d21) - 2.f * qz * (d00 + d11) + qw * (d10 - d01));
  float gw = 2.f * (qz * (d10 - d01) + qy * (d02 - d20) + qx * (d21 - d12));
  // through the normalisation q_hat = q/|q|
  float dot = qx * gx + qy * gy + qz * gz + qw * gw;
  float e0 = (gx - qx * dot) * inv, e1 = (gy - qy * dot) * inv, e2 = (gz - qz * dot) * inv, e3 = (gw - qw * dot) * inv;
  // q = (v a(n), cos n), a = sin(n)/n
  d9[0] = a * e0; d9[1] = a * e1; d9[2] = a * e2;
  if (!clamped) {
    float dadn = (n < 1e-2f) ? (-n / 3.0f + n * n * n / 30.0f) : (n * cs - sn) / (n * n);
    float vd = v0 * e0 + v1 * e1 + v2 * e2;
    float coef = (vd * dadn - e3 * sn) / n;
    d9[0] += coef * v0; d9[1] += coef * v1; d9[2] += coef * v2;
  }
}

}  // namespace mcf
