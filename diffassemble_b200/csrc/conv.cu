// Convolution pieces of the EfficientNet-B0 patch encoder (scope row N4: Eff_GAT.visual_features,
// puzzle_diff/model/backbones/efficient_gat.py:40-42,149-189).  Everything is NHWC fp32 so that the point-wise (1x1)
// convolutions -- 90 % of the encoder's FLOPs -- are plain GEMMs over [N*H*W, C] rows on the library's linear operator
// (da_op_linear, bias + SiLU fused); what is left are the 3x3 stem, the depth-wise 3x3 / 5x5 convolutions and the
// squeeze-excite reductions below, all HBM-bound streaming kernels with eval-mode BatchNorm folded into weights / bias.
// The encoder runs once per sample (spatial_diffusion.py:653), not once per denoising step.
#include "common.cuh"

namespace da {
namespace {

// direct convolution for a small input-channel count (the stem: 3 -> 32, 3x3 / 2): one thread per output value,
// output channels fastest (coalesced stores; the few weights stay in L1)
__global__ void conv2d_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                   float* __restrict__ y, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                                   int Ho, int Wo, int act) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * Ho * Wo * Cout) return;
  const int co = (int)(idx % Cout);
  size_t p = idx / Cout;
  const int ow = (int)(p % Wo); p /= Wo;
  const int oh = (int)(p % Ho);
  const int n = (int)(p / Ho);
  float acc = b ? __ldg(b + co) : 0.f;
  for (int kh = 0; kh < k; ++kh) {
    const int ih = oh * stride - pad + kh;
    if (ih < 0 || ih >= H) continue;
    for (int kw = 0; kw < k; ++kw) {
      const int iw = ow * stride - pad + kw;
      if (iw < 0 || iw >= W) continue;
      const float* xp = x + (((size_t)n * H + ih) * W + iw) * Cin;
      const float* wp = w + (((size_t)co * k + kh) * k + kw) * Cin;
      for (int ci = 0; ci < Cin; ++ci) acc = fmaf(__ldg(xp + ci), __ldg(wp + ci), acc);
    }
  }
  y[idx] = apply_act_rt(acc, act);
}

// depth-wise k x k convolution: one thread per (pixel, 4 channels); channels fastest (float4, fully coalesced)
__global__ void dwconv2d_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                     float* __restrict__ y, int N, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                                     int act) {
  const int C4 = C >> 2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * Ho * Wo * C4) return;
  const int c = (int)(idx % C4) * 4;
  size_t p = idx / C4;
  const int ow = (int)(p % Wo); p /= Wo;
  const int oh = (int)(p % Ho);
  const int n = (int)(p / Ho);
  float4 acc = b ? __ldg(reinterpret_cast<const float4*>(b + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int kh = 0; kh < k; ++kh) {
    const int ih = oh * stride - pad + kh;
    if (ih < 0 || ih >= H) continue;
    for (int kw = 0; kw < k; ++kw) {
      const int iw = ow * stride - pad + kw;
      if (iw < 0 || iw >= W) continue;
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + ih) * W + iw) * C + c));
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + ((size_t)kh * k + kw) * C + c));
      acc.x = fmaf(xv.x, wv.x, acc.x); acc.y = fmaf(xv.y, wv.y, acc.y);
      acc.z = fmaf(xv.z, wv.z, acc.z); acc.w = fmaf(xv.w, wv.w, acc.w);
    }
  }
  acc.x = apply_act_rt(acc.x, act); acc.y = apply_act_rt(acc.y, act);
  acc.z = apply_act_rt(acc.z, act); acc.w = apply_act_rt(acc.w, act);
  *reinterpret_cast<float4*>(y + (((size_t)n * Ho + oh) * Wo + ow) * C + c) = acc;
}

// squeeze: y[n, c] = mean over the HW pixels of x[n, :, c]; one thread per (n, c), channels fastest
__global__ void spatial_mean_kernel(const float* __restrict__ x, float* __restrict__ y, int ldy, int N, int HW, int C) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * C) return;
  const int c = (int)(idx % C), n = (int)(idx / C);
  const float* p = x + (size_t)n * HW * C + c;
  float s = 0.f;
  for (int i = 0; i < HW; ++i) s += __ldg(p + (size_t)i * C);
  y[(size_t)n * ldy + c] = s / (float)HW;
}

// excite: x[n, p, c] *= gate[n, c]
__global__ void channel_scale_kernel(float* __restrict__ x, const float* __restrict__ gate, int ldg, int N, int HW, int C) {
  const int C4 = C >> 2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * HW * C4) return;
  const int c = (int)(idx % C4) * 4;
  const int n = (int)(idx / ((size_t)HW * C4));
  float4 v = *reinterpret_cast<float4*>(x + idx * 4);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)n * ldg + c));
  v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
  *reinterpret_cast<float4*>(x + idx * 4) = v;
}

// input layout + normalisation in one pass: y[n, h, w, c] = (x[n, c, h, w] - mean[c]) / std[c]  (efficient_gat.py:150)
__global__ void normalize_to_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ stdv,
                                         float* __restrict__ y, int N, int C, int HW) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * HW * C) return;
  const int c = (int)(idx % C);
  const size_t p = idx / C;
  const int hw = (int)(p % HW), n = (int)(p / HW);
  y[idx] = (x[((size_t)n * C + c) * HW + hw] - __ldg(mean + c)) / __ldg(stdv + c);
}

__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, size_t n4) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  float4 a = reinterpret_cast<float4*>(y)[idx];
  const float4 b = __ldg(reinterpret_cast<const float4*>(x) + idx);
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  reinterpret_cast<float4*>(y)[idx] = a;
}

}  // namespace

cudaError_t launch_normalize_to_nhwc(const float* x, const float* mean, const float* stdv, float* y, int N, int C, int HW, cudaStream_t s) {
  const size_t total = (size_t)N * C * HW;
  if (total == 0) return cudaSuccess;
  normalize_to_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, mean, stdv, y, N, C, HW);
  return cudaGetLastError();
}

cudaError_t launch_add_inplace(float* y, const float* x, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (n % 4) return cudaErrorInvalidValue;
  add_inplace_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(y, x, n / 4);
  return cudaGetLastError();
}

cudaError_t launch_conv2d_nhwc(const float* x, const float* w, const float* b, float* y, int N, int H, int W, int Cin, int Cout,
                               int k, int stride, int pad, int act, cudaStream_t s) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const size_t total = (size_t)N * Ho * Wo * Cout;
  if (total == 0) return cudaSuccess;
  conv2d_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, w, b, y, N, H, W, Cin, Cout, k, stride, pad, Ho, Wo, act);
  return cudaGetLastError();
}

cudaError_t launch_dwconv2d_nhwc(const float* x, const float* w, const float* b, float* y, int N, int H, int W, int C, int k,
                                 int stride, int pad, int act, cudaStream_t s) {
  if (C % 4) return cudaErrorInvalidValue;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const size_t total = (size_t)N * Ho * Wo * (C / 4);
  if (total == 0) return cudaSuccess;
  dwconv2d_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, w, b, y, N, H, W, C, k, stride, pad, Ho, Wo, act);
  return cudaGetLastError();
}

cudaError_t launch_spatial_mean(const float* x, float* y, int ldy, int N, int HW, int C, cudaStream_t s) {
  const size_t total = (size_t)N * C;
  if (total == 0) return cudaSuccess;
  spatial_mean_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, ldy, N, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_channel_scale(float* x, const float* gate, int ldg, int N, int HW, int C, cudaStream_t s) {
  if (C % 4 || ldg % 4) return cudaErrorInvalidValue;
  const size_t total = (size_t)N * HW * (C / 4);
  if (total == 0) return cudaSuccess;
  channel_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, gate, ldg, N, HW, C);
  return cudaGetLastError();
}

}  // namespace da
