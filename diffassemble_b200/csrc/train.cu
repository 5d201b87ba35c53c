// Training-side kernels (scope row N1): backward of the TransformerConv attention stage over CSR / CSC, and
// the weight-gradient ("TN") GEMM.  Forward kernels are shared with inference (attn_csr.cu, gemm_*.cu).
//
// Forward (per target i, head h, in-edges e = (j -> i)):  a_e = scale <q_i, k_j>,  alpha_e = exp(a_e - m_i) / (l_i + 1e-16),
//                                                        o_i = sum_e alpha_e v_j          (+ skip, handled by the caller)
// Backward given dO:   dp_e = <dO_i, v_j>,  delta_i = sum_e alpha_e dp_e,  ds_e = alpha_e (dp_e - delta_i)
//                      dq_i = scale sum_e ds_e k_j          (by target: CSR)
//                      dk_j = scale sum_e ds_e q_i,  dv_j = sum_e alpha_e dO_i      (by source: CSC, no atomics)
// (PyG's softmax denominator l + 1e-16 is treated as l for the derivative, exactly as autograd does through
//  out / (out_sum + 1e-16) when out_sum >> 1e-16.)
#include "common.cuh"

namespace da {
namespace {

constexpr int TW = 8;  // warps per CTA

// ---- by target: delta_i and dq_i ------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(TW * 32)
attn_bwd_target_kernel(const float* __restrict__ qkvs, int ld, const float* __restrict__ dO, int ldo,
                       const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                       const float* __restrict__ stats, int n, int H, int C, float scale,
                       float* __restrict__ dqkvs, int ldg, float* __restrict__ delta) {
  extern __shared__ __align__(16) float sm[];   // per warp: q[C] | dO[C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * TW + warp;
  if (gw >= (long long)n * H) return;
  const int node = (int)(gw / H), head = (int)(gw % H);
  const int HC = H * C;
  float* q = sm + warp * 2 * C;
  float* go = q + C;
  for (int c = lane; c < C; c += 32) {
    q[c] = qkvs[(size_t)node * ld + head * C + c] * scale;
    go[c] = dO[(size_t)node * ldo + head * C + c];
  }
  __syncwarp();
  const float m = stats[((size_t)node * H + head) * 2 + 0];
  const float inv_l = 1.f / (stats[((size_t)node * H + head) * 2 + 1] + 1e-16f);
  const float* Kb = qkvs + HC + head * C;
  const float* Vb = qkvs + 2 * HC + head * C;
  const int beg = rowptr[node], end = rowptr[node + 1];
  float a1[R], a2[R];   // sum alpha dp k , sum alpha k   (channel = lane + 32 r)
#pragma unroll
  for (int r = 0; r < R; ++r) { a1[r] = 0.f; a2[r] = 0.f; }
  float dsum = 0.f;
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    const bool valid = e < end;
    const int j = valid ? col[e] : 0;
    float alpha = 0.f, dp = 0.f;
    if (valid) {
      const float* kr = Kb + (size_t)j * ld;
      const float* vr = Vb + (size_t)j * ld;
      float s = 0.f;
      if ((C & 3) == 0) {
        for (int c = 0; c < C; c += 4) {
          const float4 kk = __ldg(reinterpret_cast<const float4*>(kr + c)), vv = __ldg(reinterpret_cast<const float4*>(vr + c));
          const float4 qq = *reinterpret_cast<const float4*>(q + c), gg = *reinterpret_cast<const float4*>(go + c);
          s = fmaf(kk.x, qq.x, s); s = fmaf(kk.y, qq.y, s); s = fmaf(kk.z, qq.z, s); s = fmaf(kk.w, qq.w, s);
          dp = fmaf(vv.x, gg.x, dp); dp = fmaf(vv.y, gg.y, dp); dp = fmaf(vv.z, gg.z, dp); dp = fmaf(vv.w, gg.w, dp);
        }
      } else {
        for (int c = 0; c < C; ++c) { s = fmaf(__ldg(kr + c), q[c], s); dp = fmaf(__ldg(vr + c), go[c], dp); }
      }
      alpha = expf(s - m) * inv_l;
    }
    dsum += warp_sum(alpha * dp);
    const int cnt = min(32, end - base);
    for (int t = 0; t < cnt; ++t) {
      const float at = __shfl_sync(0xffffffffu, alpha, t);
      const float adp = at * __shfl_sync(0xffffffffu, dp, t);
      const float* kr = Kb + (size_t)__shfl_sync(0xffffffffu, j, t) * ld;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = lane + 32 * r;
        if (c < C) { const float kv = __ldg(kr + c); a1[r] = fmaf(adp, kv, a1[r]); a2[r] = fmaf(at, kv, a2[r]); }
      }
    }
  }
  if (lane == 0) delta[(size_t)node * H + head] = dsum;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = lane + 32 * r;
    if (c < C) dqkvs[(size_t)node * ldg + head * C + c] = scale * (a1[r] - dsum * a2[r]);
  }
}

// ---- by source: dk_j and dv_j ---------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(TW * 32)
attn_bwd_source_kernel(const float* __restrict__ qkvs, int ld, const float* __restrict__ dO, int ldo,
                       const int32_t* __restrict__ colptr, const int32_t* __restrict__ row,
                       const float* __restrict__ stats, const float* __restrict__ delta, int n, int H, int C,
                       float scale, float* __restrict__ dqkvs, int ldg) {
  extern __shared__ __align__(16) float sm[];   // per warp: k[C] | v[C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * TW + warp;
  if (gw >= (long long)n * H) return;
  const int node = (int)(gw / H), head = (int)(gw % H);   // node = source j
  const int HC = H * C;
  float* k = sm + warp * 2 * C;
  float* v = k + C;
  for (int c = lane; c < C; c += 32) {
    k[c] = qkvs[(size_t)node * ld + HC + head * C + c];
    v[c] = qkvs[(size_t)node * ld + 2 * HC + head * C + c];
  }
  __syncwarp();
  const int beg = colptr[node], end = colptr[node + 1];
  float dk[R], dv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { dk[r] = 0.f; dv[r] = 0.f; }
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    const bool valid = e < end;
    const int i = valid ? row[e] : 0;   // target
    float alpha = 0.f, ds = 0.f;
    if (valid) {
      const float* qr = qkvs + (size_t)i * ld + head * C;
      const float* gr = dO + (size_t)i * ldo + head * C;
      float s = 0.f, dp = 0.f;
      if ((C & 3) == 0) {
        for (int c = 0; c < C; c += 4) {
          const float4 qq = __ldg(reinterpret_cast<const float4*>(qr + c)), gg = __ldg(reinterpret_cast<const float4*>(gr + c));
          const float4 kk = *reinterpret_cast<const float4*>(k + c), vv = *reinterpret_cast<const float4*>(v + c);
          s = fmaf(qq.x, kk.x, s); s = fmaf(qq.y, kk.y, s); s = fmaf(qq.z, kk.z, s); s = fmaf(qq.w, kk.w, s);
          dp = fmaf(gg.x, vv.x, dp); dp = fmaf(gg.y, vv.y, dp); dp = fmaf(gg.z, vv.z, dp); dp = fmaf(gg.w, vv.w, dp);
        }
      } else {
        for (int c = 0; c < C; ++c) { s = fmaf(__ldg(qr + c), k[c], s); dp = fmaf(__ldg(gr + c), v[c], dp); }
      }
      const float m = stats[((size_t)i * H + head) * 2 + 0];
      const float inv_l = 1.f / (stats[((size_t)i * H + head) * 2 + 1] + 1e-16f);
      alpha = expf(s * scale - m) * inv_l;
      ds = alpha * (dp - delta[(size_t)i * H + head]);
    }
    const int cnt = min(32, end - base);
    for (int t = 0; t < cnt; ++t) {
      const float at = __shfl_sync(0xffffffffu, alpha, t);
      const float dst = __shfl_sync(0xffffffffu, ds, t);
      const int it = __shfl_sync(0xffffffffu, i, t);
      const float* qr = qkvs + (size_t)it * ld + head * C;
      const float* gr = dO + (size_t)it * ldo + head * C;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = lane + 32 * r;
        if (c < C) { dk[r] = fmaf(dst, __ldg(qr + c), dk[r]); dv[r] = fmaf(at, __ldg(gr + c), dv[r]); }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = lane + 32 * r;
    if (c < C) {
      dqkvs[(size_t)node * ldg + HC + head * C + c] = scale * dk[r];
      dqkvs[(size_t)node * ldg + 2 * HC + head * C + c] = dv[r];
    }
  }
}

// ---- dense puzzle graphs: one CTA per (graph, head), operands of the whole graph staged in shared memory -------
// The edge-list kernels above gather K / V (or Q / dO) rows from L2 once per EDGE: 144 times per row on a
// dense 12x12 puzzle, which makes them L2-gather bound (12.5 of the 16.8 ms of a c5 training step).  When a
// graph's whole edge multiset sits in its adjacency bitmap (DensePlan, no residual edges) and its [n, C] operands
// fit in shared memory, the same two passes run out of shared memory: rows are padded to C + 4 floats so that
// lane-per-neighbour dot products (float4 reads at stride C + 4: 8 lanes per phase hit 8 distinct bank groups for
// C = 32 and 144) and lane-per-channel updates (stride 1) are both conflict free.  (Requires C % 4 == 0.)
struct GraphDesc { int32_t node0, n, bm_words, pad; int64_t bm_off; };

__device__ __forceinline__ bool bm_bit(const uint32_t* bm, int bm_words, int i, int j) {
  return (bm[i * bm_words + (j >> 5)] >> (j & 31)) & 1u;
}

// MODE 0: by target (A = K, B = V staged; per row i: delta_i, dq_i).  MODE 1: by source (A = Q, B = dO staged; per
// row j: dk_j, dv_j).
template <int MODE>
__global__ void __launch_bounds__(TW * 32)
attn_bwd_graph_kernel(const float* __restrict__ qkvs, int ld, const float* __restrict__ dO, int ldo,
                      const GraphDesc* __restrict__ graphs, const uint32_t* __restrict__ bitmap,
                      const float* __restrict__ stats, float* __restrict__ delta, int H, int C, float scale,
                      float* __restrict__ dqkvs, int ldg) {
  extern __shared__ __align__(16) float sm[];
  const GraphDesc gd = graphs[blockIdx.x];
  const int head = blockIdx.y, n = gd.n, HC = H * C, P = C + 4;
  float* A = sm;                          // [n][P]
  float* B = A + (size_t)n * P;           // [n][P]
  const int n4 = (n + 3) & ~3;            // (16-byte aligned per-warp vectors)
  float* wbuf = B + (size_t)n * P;        // per warp: a[C] | b[C] | w1[n4] | w2[n4]
  uint32_t* bm = reinterpret_cast<uint32_t*>(wbuf + (size_t)TW * (2 * C + 2 * n4));   // [n][bm_words]
  float* mrow = reinterpret_cast<float*>(bm + (size_t)n * gd.bm_words);               // [n] max
  float* lrow = mrow + n;                                                             // [n] 1 / (l + 1e-16)
  float* drow = lrow + n;                                                             // [n] delta (MODE 1)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int offA = (MODE == 0 ? HC : 0) + head * C;        // K (by target) or Q (by source)
  for (int idx = threadIdx.x; idx < n * C; idx += TW * 32) {
    const int r = idx / C, c = idx - r * C;
    const size_t node = (size_t)(gd.node0 + r);
    A[r * P + c] = qkvs[node * ld + offA + c];
    B[r * P + c] = MODE == 0 ? qkvs[node * ld + 2 * HC + head * C + c] : dO[node * ldo + head * C + c];
  }
  for (int idx = threadIdx.x; idx < n * gd.bm_words; idx += TW * 32) bm[idx] = bitmap[gd.bm_off + idx];
  for (int r = threadIdx.x; r < n; r += TW * 32) {
    const size_t sidx = ((size_t)(gd.node0 + r) * H + head) * 2;
    mrow[r] = stats[sidx];
    lrow[r] = 1.f / (stats[sidx + 1] + 1e-16f);
    if (MODE == 1) drow[r] = delta[(size_t)(gd.node0 + r) * H + head];
  }
  __syncthreads();
  float* a = wbuf + (size_t)warp * (2 * C + 2 * n4);
  float* b = a + C;
  float* w1 = b + C;     // MODE 0: alpha * dp   | MODE 1: ds
  float* w2 = w1 + n4;   // MODE 0: alpha        | MODE 1: alpha
  for (int r = warp; r < n; r += TW) {
    const size_t node = (size_t)(gd.node0 + r);
    for (int c = lane; c < C; c += 32) {
      if (MODE == 0) { a[c] = qkvs[node * ld + head * C + c] * scale; b[c] = dO[node * ldo + head * C + c]; }
      else { a[c] = qkvs[node * ld + HC + head * C + c]; b[c] = qkvs[node * ld + 2 * HC + head * C + c]; }
    }
    __syncwarp();
    float dsum = 0.f;
    for (int base = 0; base < n; base += 32) {
      const int o = base + lane;   // the other end of the edge: source j (MODE 0) / target i (MODE 1)
      const bool valid = o < n && (MODE == 0 ? bm_bit(bm, gd.bm_words, r, o) : bm_bit(bm, gd.bm_words, o, r));
      float x1 = 0.f, x2 = 0.f;
      if (valid) {
        const float* ar = A + o * P;
        const float* br = B + o * P;
        // float4 reads: the per-lane row reads cost the same shared-memory bandwidth, the broadcast reads of the
        // row's own vector a quarter; four independent FMA chains per dot product
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 2
        for (int c = 0; c < C; c += 4) {
          const float4 av = *reinterpret_cast<const float4*>(ar + c), bv = *reinterpret_cast<const float4*>(br + c);
          const float4 aa = *reinterpret_cast<const float4*>(a + c), bb = *reinterpret_cast<const float4*>(b + c);
          s0 = fmaf(av.x, aa.x, s0); s1 = fmaf(av.y, aa.y, s1); s2 = fmaf(av.z, aa.z, s2); s3 = fmaf(av.w, aa.w, s3);
          d0 = fmaf(bv.x, bb.x, d0); d1 = fmaf(bv.y, bb.y, d1); d2 = fmaf(bv.z, bb.z, d2); d3 = fmaf(bv.w, bb.w, d3);
        }
        const float s = (s0 + s1) + (s2 + s3), dp = (d0 + d1) + (d2 + d3);
        if (MODE == 0) {
          const float alpha = expf(s - mrow[r]) * lrow[r];
          x1 = alpha * dp; x2 = alpha;
        } else {
          const float alpha = expf(s * scale - mrow[o]) * lrow[o];
          x1 = alpha * (dp - drow[o]); x2 = alpha;
        }
      }
      if (o < n) { w1[o] = x1; w2[o] = x2; }
      if (MODE == 0) dsum += warp_sum(x1);
    }
    __syncwarp();
    // lane-per-channel accumulation over the row's neighbours (zeros for absent edges)
    for (int c = lane; c < C; c += 32) {
      float acc1 = 0.f, acc2 = 0.f, acc1b = 0.f, acc2b = 0.f;
      int o = 0;
      for (; o + 4 <= n; o += 4) {   // the per-neighbour weights are broadcast four at a time
        const float4 x1 = *reinterpret_cast<const float4*>(w1 + o), x2 = *reinterpret_cast<const float4*>(w2 + o);
        const float a0 = A[o * P + c], a1 = A[(o + 1) * P + c], a2 = A[(o + 2) * P + c], a3 = A[(o + 3) * P + c];
        acc1 = fmaf(x1.x, a0, acc1); acc1b = fmaf(x1.y, a1, acc1b); acc1 = fmaf(x1.z, a2, acc1); acc1b = fmaf(x1.w, a3, acc1b);
        if (MODE == 0) {
          acc2 = fmaf(x2.x, a0, acc2); acc2b = fmaf(x2.y, a1, acc2b); acc2 = fmaf(x2.z, a2, acc2); acc2b = fmaf(x2.w, a3, acc2b);
        } else {
          acc2 = fmaf(x2.x, B[o * P + c], acc2); acc2b = fmaf(x2.y, B[(o + 1) * P + c], acc2b);
          acc2 = fmaf(x2.z, B[(o + 2) * P + c], acc2); acc2b = fmaf(x2.w, B[(o + 3) * P + c], acc2b);
        }
      }
      for (; o < n; ++o) { acc1 = fmaf(w1[o], A[o * P + c], acc1); acc2 = fmaf(w2[o], MODE == 0 ? A[o * P + c] : B[o * P + c], acc2); }
      acc1 += acc1b; acc2 += acc2b;
      if (MODE == 0) dqkvs[node * ldg + head * C + c] = scale * (acc1 - dsum * acc2);
      else {
        dqkvs[node * ldg + HC + head * C + c] = scale * acc1;       // dk_j = scale sum_i ds_ij q_i
        dqkvs[node * ldg + 2 * HC + head * C + c] = acc2;           // dv_j = sum_i alpha_ij dO_i
      }
    }
    if (MODE == 0 && lane == 0) delta[node * H + head] = dsum;
    __syncwarp();
  }
}

size_t attn_bwd_graph_smem(int n, int C, int bm_words) {
  const size_t n4 = ((size_t)n + 3) & ~(size_t)3;
  return sizeof(float) * ((size_t)2 * n * (C + 4) + (size_t)TW * (2 * C + 2 * n4) + 3 * (size_t)n) + sizeof(uint32_t) * (size_t)n * bm_words;
}

__global__ void copy_skip_grad_kernel(const float* __restrict__ dO, int ldo, float* __restrict__ dqkvs, int ldg, int n, int HC) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * HC) return;
  const int r = (int)(idx / HC), c = (int)(idx % HC);
  dqkvs[(size_t)r * ldg + 3 * HC + c] = dO[(size_t)r * ldo + c];
}

// ---- weight gradient: dW[N, K] = sum_m dY[m, N] * X[m, K]   (split over m, fp32 atomics) ------------------
constexpr int WG_BN = 64, WG_BK = 64, WG_BM = 32, WG_NT = 256;
__global__ void __launch_bounds__(WG_NT)
linear_wgrad_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx, float* __restrict__ dW,
                    int ldw, int M, int N, int K, int m_chunk) {
  __shared__ float ys[WG_BM][WG_BN + 4];
  __shared__ float xs[WG_BM][WG_BK + 4];
  const int n0 = blockIdx.x * WG_BN, k0 = blockIdx.y * WG_BK;
  const int m_beg = blockIdx.z * m_chunk, m_end = min(M, m_beg + m_chunk);
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;   // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int m0 = m_beg; m0 < m_end; m0 += WG_BM) {
    for (int idx = tid; idx < WG_BM * WG_BN; idx += WG_NT) {
      const int mm = idx / WG_BN, nn = idx % WG_BN;
      ys[mm][nn] = (m0 + mm < m_end && n0 + nn < N) ? dY[(size_t)(m0 + mm) * ldy + n0 + nn] : 0.f;
    }
    for (int idx = tid; idx < WG_BM * WG_BK; idx += WG_NT) {
      const int mm = idx / WG_BK, kk = idx % WG_BK;
      xs[mm][kk] = (m0 + mm < m_end && k0 + kk < K) ? X[(size_t)(m0 + mm) * ldx + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int mm = 0; mm < WG_BM; ++mm) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = ys[mm][ty * 4 + i]; b[i] = xs[mm][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + ty * 4 + i, kk = k0 + tx * 4 + j;
      if (nn < N && kk < K) atomicAdd(dW + (size_t)nn * ldw + kk, acc[i][j]);
    }
}

__global__ void colsum_kernel(const float* __restrict__ dY, int ldy, float* __restrict__ db, int M, int N, int m_chunk) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m_beg = blockIdx.y * m_chunk, m_end = min(M, m_beg + m_chunk);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // independent chains: the loads of a chunk are all in flight
  int m = m_beg;
  for (; m + 4 <= m_end; m += 4) {
    s0 += dY[(size_t)m * ldy + n]; s1 += dY[(size_t)(m + 1) * ldy + n];
    s2 += dY[(size_t)(m + 2) * ldy + n]; s3 += dY[(size_t)(m + 3) * ldy + n];
  }
  for (; m < m_end; ++m) s0 += dY[(size_t)m * ldy + n];
  atomicAdd(db + n, (s0 + s1) + (s2 + s3));
}

// ---------------------------------------------------------------------------------------------
// Fused Adafactor (transformers.optimization.Adafactor.step with its default arguments: factored second
// moments for matrices, relative step, parameter scaling, update clipping, no first moment).  One CTA per
// parameter tensor walks the whole update: the stock implementation issues ~30 small kernels and one host
// synchronisation (`max(eps, RMS)` on a device scalar) per tensor.
// ---------------------------------------------------------------------------------------------
constexpr int AF_NT = 1024;

__device__ __forceinline__ float af_block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < AF_NT / 32) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__global__ void __launch_bounds__(AF_NT)
adafactor_kernel(const da_adafactor_param* __restrict__ params, float eps1, float eps2, float clip, float weight_decay) {
  __shared__ float red[33];
  const da_adafactor_param q = params[blockIdx.x];
  const int R = q.rows, C = q.cols;
  const size_t n = (size_t)R * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float beta = q.beta2t, omb = 1.f - q.beta2t;
  float pp = 0.f;   // ||p||^2
  float uu = 0.f;   // ||update||^2
  float row_mean_all = 1.f;
  const bool factored = q.sq_row != nullptr;
  if (factored) {
    // pass A: column means of g^2 + eps1 (thread per column, coalesced rows)
    for (int c = threadIdx.x; c < C; c += AF_NT) {
      float s0 = 0.f, s1 = 0.f;
      int r = 0;
      for (; r + 2 <= R; r += 2) {
        const float g0 = q.g[(size_t)r * C + c], g1 = q.g[(size_t)(r + 1) * C + c];
        s0 += g0 * g0 + eps1; s1 += g1 * g1 + eps1;
      }
      if (r < R) { const float g0 = q.g[(size_t)r * C + c]; s0 += g0 * g0 + eps1; }
      q.sq_col[c] = beta * q.sq_col[c] + omb * ((s0 + s1) / (float)R);
    }
    // pass B: row means (warp per row) and ||p||^2
    float rsum = 0.f;
    for (int r = warp; r < R; r += AF_NT / 32) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float g = q.g[(size_t)r * C + c], p = q.p[(size_t)r * C + c];
        s += g * g + eps1;
        pp += p * p;
      }
      s = warp_sum(s);
      const float nr = beta * q.sq_row[r] + omb * (s / (float)C);
      if (lane == 0) { q.sq_row[r] = nr; rsum += nr; }
    }
    row_mean_all = af_block_sum(rsum, red) / (float)R;
    pp = af_block_sum(pp, red);
    // pass C: ||update||^2 with update = g * rsqrt(row / mean(row)) * rsqrt(col)
    for (int r = warp; r < R; r += AF_NT / 32) {
      const float rf = rsqrtf(q.sq_row[r] / row_mean_all);
      for (int c = lane; c < C; c += 32) {
        const float u = q.g[(size_t)r * C + c] * rf * rsqrtf(q.sq_col[c]);
        uu += u * u;
      }
    }
  } else {
    for (size_t i = threadIdx.x; i < n; i += AF_NT) {
      const float g = q.g[i], p = q.p[i];
      const float v = beta * q.sq[i] + omb * (g * g + eps1);
      q.sq[i] = v;
      const float u = g * rsqrtf(v);
      uu += u * u;
      pp += p * p;
    }
    pp = af_block_sum(pp, red);
  }
  uu = af_block_sum(uu, red);
  const float rms_p = sqrtf(pp) / sqrtf((float)n);
  const float rms_u = sqrtf(uu) / sqrtf((float)n);
  const float lr = fmaxf(eps2, rms_p) * q.rel_step;
  const float scale = lr / fmaxf(1.f, rms_u / clip);
  if (threadIdx.x == 0 && q.rms_out) *q.rms_out = rms_p;
  if (factored) {
    for (int r = warp; r < R; r += AF_NT / 32) {
      const float rf = rsqrtf(q.sq_row[r] / row_mean_all) * scale;
      for (int c = lane; c < C; c += 32) {
        const size_t i = (size_t)r * C + c;
        float p = q.p[i];
        if (weight_decay != 0.f) p += p * (-weight_decay * lr);
        q.p[i] = p - q.g[i] * rf * rsqrtf(q.sq_col[c]);
      }
    }
  } else {
    for (size_t i = threadIdx.x; i < n; i += AF_NT) {
      float p = q.p[i];
      if (weight_decay != 0.f) p += p * (-weight_decay * lr);
      q.p[i] = p - q.g[i] * rsqrtf(q.sq[i]) * scale;
    }
  }
}

}  // namespace

cudaError_t launch_adafactor(const da_adafactor_param* params_dev, int n, float eps1, float eps2, float clip, float weight_decay,
                             cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  adafactor_kernel<<<n, AF_NT, 0, s>>>(params_dev, eps1, eps2, clip, weight_decay);
  return cudaGetLastError();
}

cudaError_t launch_attn_backward(const float* qkvs, const float* dO, const CsrGraph& by_target, const CsrGraph& by_source,
                                 const float* stats, int n, int H, int C, float* dqkvs, float* delta, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int HC = H * C, ld = 4 * HC;
  const long long warps = (long long)n * H;
  const unsigned grid = (unsigned)((warps + TW - 1) / TW);
  const size_t smem = (size_t)TW * 2 * C * sizeof(float);
  const float scale = 1.0f / sqrtf((float)C);
  const int R = (C + 31) / 32;
#define DA_LAUNCH(RR)                                                                                                      \
  do {                                                                                                                     \
    attn_bwd_target_kernel<RR><<<grid, TW * 32, smem, s>>>(qkvs, ld, dO, HC, by_target.rowptr, by_target.col, stats, n, H, \
                                                           C, scale, dqkvs, ld, delta);                                    \
    attn_bwd_source_kernel<RR><<<grid, TW * 32, smem, s>>>(qkvs, ld, dO, HC, by_source.rowptr, by_source.col, stats,       \
                                                           delta, n, H, C, scale, dqkvs, ld);                              \
  } while (0)
  if (R <= 1) DA_LAUNCH(1);
  else if (R <= 2) DA_LAUNCH(2);
  else if (R <= 5) DA_LAUNCH(5);
  else if (R <= 13) DA_LAUNCH(13);
  else return cudaErrorInvalidValue;
#undef DA_LAUNCH
  const size_t total = (size_t)n * HC;
  copy_skip_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dO, HC, dqkvs, ld, n, HC);
  return cudaGetLastError();
}

bool attn_backward_dense_fits(int n_max, int C, int bm_words_max) {
  if (C % 4) return false;
  return attn_bwd_graph_smem(n_max, C, bm_words_max) <= (size_t)227 * 1024;
}

cudaError_t launch_attn_backward_dense(const float* qkvs, const float* dO, const void* graphs_dev, int n_graphs, int n_max,
                                       int bm_words_max, const uint32_t* bitmap, const float* stats, int n, int H, int C,
                                       float* dqkvs, float* delta, cudaStream_t s) {
  if (n_graphs <= 0) return cudaSuccess;
  const size_t smem = attn_bwd_graph_smem(n_max, C, bm_words_max);
  if (smem > (size_t)227 * 1024) return cudaErrorInvalidValue;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_graph_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_graph_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  const int HC = H * C, ld = 4 * HC;
  const float scale = 1.0f / sqrtf((float)C);
  const GraphDesc* gd = reinterpret_cast<const GraphDesc*>(graphs_dev);
  dim3 grid((unsigned)n_graphs, (unsigned)H);
  attn_bwd_graph_kernel<0><<<grid, TW * 32, smem, s>>>(qkvs, ld, dO, HC, gd, bitmap, stats, delta, H, C, scale, dqkvs, ld);
  attn_bwd_graph_kernel<1><<<grid, TW * 32, smem, s>>>(qkvs, ld, dO, HC, gd, bitmap, stats, delta, H, C, scale, dqkvs, ld);
  const size_t total = (size_t)n * HC;
  copy_skip_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dO, HC, dqkvs, ld, n, HC);
  return cudaGetLastError();
}

cudaError_t launch_linear_wgrad(const float* dY, const float* X, float* dW, float* db, int M, int N, int K, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, s);
  if (e != cudaSuccess) return e;
  int splits = (M + 2047) / 2048;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  int m_chunk = ((M + splits - 1) / splits + WG_BM - 1) / WG_BM * WG_BM;
  dim3 grid((N + WG_BN - 1) / WG_BN, (K + WG_BK - 1) / WG_BK, (M + m_chunk - 1) / m_chunk);
  linear_wgrad_kernel<<<grid, WG_NT, 0, s>>>(dY, N, X, K, dW, K, M, N, K, m_chunk);
  if (db) {
    e = cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, s);
    if (e != cudaSuccess) return e;
    const int c_chunk = 64;   // short row chunks: thousands of CTAs instead of a few dozen latency-bound ones
    dim3 g2((N + 255) / 256, (M + c_chunk - 1) / c_chunk);
    colsum_kernel<<<g2, 256, 0, s>>>(dY, N, db, M, N, c_chunk);
  }
  return cudaGetLastError();
}

}  // namespace da
