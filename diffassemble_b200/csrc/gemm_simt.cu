// Exact-fp32 linear layer on CUDA cores: y = act(a @ w^T + bias).
// This is the anchor / debug path (DA_GEMM_FP32_SIMT); the production projections run on
// tcgen05 tensor cores (gemm_umma.cu).  Replaces every nn.Linear / PyG Linear on the path
// (efficient_gat.py:88-102, TransformerConv lin_{query,key,value,skip}).
#include "common.cuh"

namespace da {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PAD = 4;

template <int ACT>
__global__ void __launch_bounds__(NT, 2)
linear_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                   const float* __restrict__ bias, float* __restrict__ Cf, int ldc,
                   __nv_bfloat16* __restrict__ Chi, __nv_bfloat16* __restrict__ Clo, int ldsp, int M, int N,
                   int K) {
  __shared__ float As[2][BK][BM + PAD];
  __shared__ float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = tid / 16, tx = tid % 16;
  // global->smem mapping: each thread moves two float4 of A and two of W per k-tile
  const int lrow = tid / 4;        // 0..63
  const int lk = (tid % 4) * 4;    // 0,4,8,12

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int gm = m0 + lrow + r * 64, gn = n0 + lrow + r * 64, gk = k0 + lk;
      ra[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gm < M) {
        const float* p = A + (size_t)gm * lda + gk;
        if (gk + 3 < K) ra[r] = *reinterpret_cast<const float4*>(p);
        else {
          if (gk + 0 < K) ra[r].x = p[0];
          if (gk + 1 < K) ra[r].y = p[1];
          if (gk + 2 < K) ra[r].z = p[2];
        }
      }
      if (gn < N) {
        const float* p = W + (size_t)gn * ldw + gk;
        if (gk + 3 < K) rb[r] = *reinterpret_cast<const float4*>(p);
        else {
          if (gk + 0 < K) rb[r].x = p[0];
          if (gk + 1 < K) rb[r].y = p[1];
          if (gk + 2 < K) rb[r].z = p[2];
        }
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int row = lrow + r * 64;
      As[buf][lk + 0][row] = ra[r].x; As[buf][lk + 1][row] = ra[r].y;
      As[buf][lk + 2][row] = ra[r].z; As[buf][lk + 3][row] = ra[r].w;
      Bs[buf][lk + 0][row] = rb[r].x; Bs[buf][lk + 1][row] = rb[r].y;
      Bs[buf][lk + 2][row] = rb[r].z; Bs[buf][lk + 3][row] = rb[r].w;
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int gn = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float b = (bias != nullptr && gn + j < N) ? bias[gn + j] : 0.f;
        v[j] = apply_act<ACT>(acc[i][jh * 4 + j] + b);
      }
      if (gn + 3 < N) {
        if (Cf) *reinterpret_cast<float4*>(Cf + (size_t)gm * ldc + gn) = make_float4(v[0], v[1], v[2], v[3]);
        if (Chi) {
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            h[j] = __float2bfloat16_rn(v[j]);
            l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
          }
          *reinterpret_cast<uint2*>(Chi + (size_t)gm * ldsp + gn) = *reinterpret_cast<uint2*>(h);
          *reinterpret_cast<uint2*>(Clo + (size_t)gm * ldsp + gn) = *reinterpret_cast<uint2*>(l);
        }
      } else {
        for (int j = 0; j < 4; ++j)
          if (gn + j < N) {
            if (Cf) Cf[(size_t)gm * ldc + gn + j] = v[j];
            if (Chi) {
              __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
              Chi[(size_t)gm * ldsp + gn + j] = h;
              Clo[(size_t)gm * ldsp + gn + j] = __float2bfloat16_rn(v[j] - __bfloat162float(h));
            }
          }
      }
    }
  }
}

// row_map (optional): output row r is taken from input row row_map[r] (the planner's internal node order)
__global__ void split_bf16_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int ld, int rows, int cols,
                                  const int32_t* __restrict__ row_map) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)rows * cols;
  if (idx >= total) return;
  int r = (int)(idx / cols), c = (int)(idx % cols);
  const int rs = row_map ? row_map[r] : r;
  float v = x[(size_t)rs * ldx + c];
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[(size_t)r * ld + c] = h;
  lo[(size_t)r * ld + c] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void gather_rows_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int rows, int cols,
                                   const int32_t* __restrict__ row_map) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * cols) return;
  int r = (int)(idx / cols), c = (int)(idx % cols);
  y[(size_t)r * ldy + c] = x[(size_t)row_map[r] * ldx + c];
}

}  // namespace

cudaError_t launch_linear_simt(const float* a, int lda, const float* w, int ldw, const float* bias,
                               const LinearOut& out, int M, int N, int K, int act, cudaStream_t s) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  // float4 paths need 16-byte aligned rows
  if ((lda % 4) || (ldw % 4) || (out.f32 && (out.ldc % 4)) || (out.hi && (out.ld_split % 4))) return cudaErrorInvalidValue;
#define DA_LAUNCH(ACT)                                                                                   \
  linear_simt_kernel<ACT><<<grid, NT, 0, s>>>(a, lda, w, ldw, bias, out.f32, out.ldc, out.hi, out.lo,    \
                                              out.ld_split, M, N, K)
  switch (act) {
    case ACT_NONE: DA_LAUNCH(ACT_NONE); break;
    case ACT_GELU: DA_LAUNCH(ACT_GELU); break;
    case ACT_LRELU: DA_LAUNCH(ACT_LRELU); break;
    case ACT_RELU: DA_LAUNCH(ACT_RELU); break;
    case ACT_SILU: DA_LAUNCH(ACT_SILU); break;
    case ACT_SIGMOID: DA_LAUNCH(ACT_SIGMOID); break;
    default: return cudaErrorInvalidValue;
  }
#undef DA_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_split_bf16(const float* x, int ldx, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld,
                              int rows, int cols, cudaStream_t s, const int32_t* row_map) {
  size_t total = (size_t)rows * cols;
  if (total == 0) return cudaSuccess;
  split_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, ldx, hi, lo, ld, rows, cols, row_map);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* x, int ldx, float* y, int ldy, int rows, int cols, const int32_t* row_map,
                               cudaStream_t s) {
  size_t total = (size_t)rows * cols;
  if (total == 0) return cudaSuccess;
  gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, ldx, y, ldy, rows, cols, row_map);
  return cudaGetLastError();
}

}  // namespace da
