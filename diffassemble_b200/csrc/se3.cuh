// Closed-form per-node SO(3)/SE(3) device code for the 3D (Breaking-Bad) head and sampler.
// Replaces the torch.matrix_exp / pytorch3d / torch.linalg.eigh chains of
//   efficient_gat_3d.py:30-45,217-218, utils_3d.py:1018-1061 and
//   spatial_diffusion_3d_test_double_diffusion.py:595-685
// with Rodrigues' formula and the atan2 log map; quaternions are real-first (w,x,y,z) and are
// parity-checked up to sign (pytorch3d releases differ on sign standardisation).
#pragma once
#include "common.cuh"

namespace da {
namespace se3 {

struct Mat3 { float m[3][3]; };

__device__ __forceinline__ Mat3 mat_identity() {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = (i == j) ? 1.f : 0.f;
  return r;
}
__device__ __forceinline__ Mat3 mat_mul(const Mat3& a, const Mat3& b) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
__device__ __forceinline__ Mat3 mat_T(const Mat3& a) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}

// exp(hat(w)), hat(w) = [[0,-w2,w1],[w2,0,-w0],[-w1,w0,0]]  (vec2skew, efficient_gat_3d.py:30-35)
__device__ __forceinline__ Mat3 exp_so3(float w0, float w1, float w2) {
  const float th2 = w0 * w0 + w1 * w1 + w2 * w2;
  const float th = sqrtf(th2);
  float A, B;  // sin(th)/th, (1-cos(th))/th^2
  if (th < 1e-3f) {
    A = 1.f - th2 * (1.f / 6.f) + th2 * th2 * (1.f / 120.f);
    B = 0.5f - th2 * (1.f / 24.f) + th2 * th2 * (1.f / 720.f);
  } else {
    float s, c;
    sincosf(th, &s, &c);
    A = s / th;
    const float sh = sinf(0.5f * th);
    B = 2.f * sh * sh / th2;  // (1-cos) without cancellation
  }
  Mat3 R;
  R.m[0][0] = 1.f - B * (w1 * w1 + w2 * w2);
  R.m[1][1] = 1.f - B * (w0 * w0 + w2 * w2);
  R.m[2][2] = 1.f - B * (w0 * w0 + w1 * w1);
  R.m[0][1] = B * w0 * w1 - A * w2;
  R.m[1][0] = B * w0 * w1 + A * w2;
  R.m[0][2] = B * w0 * w2 + A * w1;
  R.m[2][0] = B * w0 * w2 - A * w1;
  R.m[1][2] = B * w1 * w2 - A * w0;
  R.m[2][1] = B * w1 * w2 + A * w0;
  return R;
}

// log map (utils_3d.py:1018-1046): returns the rotation vector theta * axis.
__device__ __forceinline__ void log_so3(const Mat3& R, float w[3]) {
  const float k0 = R.m[2][1] - R.m[1][2];   // skew[2,1]
  const float k1 = -(R.m[2][0] - R.m[0][2]);  // -skew[2,0]
  const float k2 = R.m[1][0] - R.m[0][1];   // skew[1,0]
  const float s_angle = 0.5f * sqrtf(k0 * k0 + k1 * k1 + k2 * k2);
  const float c_angle = 0.5f * (R.m[0][0] + R.m[1][1] + R.m[2][2] - 1.f);
  const float angle = atan2f(s_angle, c_angle);
  if (angle == 0.f) { w[0] = w[1] = w[2] = 0.f; return; }
  if (s_angle == 0.f) {
    // rotation by exactly pi: R = 2 a a^T - I.  The reference falls back to an eigendecomposition
    // whose axis sign is implementation-defined; take the axis from the largest diagonal entry.
    float d0 = R.m[0][0], d1 = R.m[1][1], d2 = R.m[2][2];
    float a0, a1, a2;
    if (d0 >= d1 && d0 >= d2) { a0 = sqrtf(fmaxf(0.5f * (d0 + 1.f), 0.f)); a1 = R.m[0][1] / (2.f * a0); a2 = R.m[0][2] / (2.f * a0); }
    else if (d1 >= d2)        { a1 = sqrtf(fmaxf(0.5f * (d1 + 1.f), 0.f)); a0 = R.m[0][1] / (2.f * a1); a2 = R.m[1][2] / (2.f * a1); }
    else                      { a2 = sqrtf(fmaxf(0.5f * (d2 + 1.f), 0.f)); a0 = R.m[0][2] / (2.f * a2); a1 = R.m[1][2] / (2.f * a2); }
    w[0] = angle * a0; w[1] = angle * a1; w[2] = angle * a2;
    return;
  }
  const float scale = angle / (2.f * s_angle);
  w[0] = scale * k0; w[1] = scale * k1; w[2] = scale * k2;
}

__device__ __forceinline__ Mat3 so3_scale(const Mat3& R, float c) {  // utils_3d.py:1049-1061
  float w[3];
  log_so3(R, w);
  return exp_so3(c * w[0], c * w[1], c * w[2]);
}

// pytorch3d.transforms.matrix_to_quaternion (4-candidate sqrt_positive_part method, floor 0.1)
__device__ __forceinline__ void mat_to_quat(const Mat3& M, float q[4]) {
  const float m00 = M.m[0][0], m01 = M.m[0][1], m02 = M.m[0][2];
  const float m10 = M.m[1][0], m11 = M.m[1][1], m12 = M.m[1][2];
  const float m20 = M.m[2][0], m21 = M.m[2][1], m22 = M.m[2][2];
  float qa[4] = {1.f + m00 + m11 + m22, 1.f + m00 - m11 - m22, 1.f - m00 + m11 - m22, 1.f - m00 - m11 + m22};
#pragma unroll
  for (int i = 0; i < 4; ++i) qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f;
  int best = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i) if (qa[i] > qa[best]) best = i;
  float c[4];
  if (best == 0)      { c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
  else if (best == 1) { c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20; }
  else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21; }
  else                { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3]; }
  const float d = 2.f * fmaxf(qa[best], 0.1f);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = c[i] / d;
}

// pytorch3d.transforms.quaternion_to_matrix (works for non-unit quaternions)
__device__ __forceinline__ Mat3 quat_to_mat(const float q[4]) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  const float two_s = 2.f / (r * r + i * i + j * j + k * k);
  Mat3 M;
  M.m[0][0] = 1.f - two_s * (j * j + k * k);
  M.m[0][1] = two_s * (i * j - k * r);
  M.m[0][2] = two_s * (i * k + j * r);
  M.m[1][0] = two_s * (i * j + k * r);
  M.m[1][1] = 1.f - two_s * (i * i + k * k);
  M.m[1][2] = two_s * (j * k - i * r);
  M.m[2][0] = two_s * (i * k - j * r);
  M.m[2][1] = two_s * (j * k + i * r);
  M.m[2][2] = 1.f - two_s * (i * i + j * j);
  return M;
}

// efficient_gat_3d.py:217-218: q = normalize(matrix_to_quaternion(exp(hat(r))), eps=1e-12)
__device__ __forceinline__ void axis_angle_to_unit_quat(float r0, float r1, float r2, float q[4]) {
  Mat3 R = exp_so3(r0, r1, r2);
  mat_to_quat(R, q);
  const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] /= n;
}

// One node of spatial_diffusion_3d_test_double_diffusion.py:595-685 (eta = 0).
// x, out, y are [quat(4), t(3)].
__device__ __forceinline__ void ddim_update_se3(const float x[7], const float out[7], const da_step_coef& c, float y[7]) {
  float x0[7];
  if (c.pred == DA_PRED_START_X) {
#pragma unroll
    for (int i = 0; i < 7; ++i) x0[i] = out[i];
  } else {
    const float sb = sqrtf(1.f - c.acp), sa = sqrtf(c.acp);
#pragma unroll
    for (int i = 0; i < 7; ++i) x0[i] = (x[i] - sb * out[i]) / sa;
  }
  const float sq_prev = sqrtf(c.acp_prev);
  const float sq_1m_prev = sqrtf(1.f - c.acp_prev);
  // translation: Euclidean DDIM
#pragma unroll
  for (int i = 4; i < 7; ++i) {
    const float eps = (c.sqrt_recip_acp * x[i] - x0[i]) / c.sqrt_recipm1_acp;
    y[i] = sq_prev * x0[i] + sq_1m_prev * eps;
  }
  // rotation: SO(3) DDIM through log / exp maps
  const Mat3 Rx = quat_to_mat(x);
  const Mat3 R0 = quat_to_mat(x0);
  const Mat3 term_x = so3_scale(Rx, c.sqrt_recip_acp / c.sqrt_recipm1_acp);
  const Mat3 term_0 = so3_scale(R0, 1.f / c.sqrt_recipm1_acp);
  const Mat3 eps_rot_m = mat_mul(term_x, mat_T(term_0));
  float eps_q[4];
  mat_to_quat(eps_rot_m, eps_q);
  const Mat3 dir_rot = so3_scale(quat_to_mat(eps_q), sq_1m_prev);
  const Mat3 prev_r = mat_mul(so3_scale(R0, sq_prev), dir_rot);
  mat_to_quat(prev_r, y);
}

}  // namespace se3
}  // namespace da
