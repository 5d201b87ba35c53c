// CSR-by-target multigraph attention: the message / softmax / aggregate stage of
// torch_geometric.nn.TransformerConv without ever materialising the
// [E, H, C] per-edge tensors the reference formulation gathers.
//
//   a[e,h]   = <q[i,h,:], k[j,h,:]> / sqrt(C)                      for every in-edge e = (j -> i)
//   alpha    = exp(a - max_i) / (sum_i exp(a - max_i) + 1e-16)      softmax over the in-edges of i
//   y[i,h,:] = sum_e alpha[e,h] v[j,h,:] + skip[i,h,:] (+ resid)    then optional activation
//
// One warp owns one (target node, head).  Lanes are spread over the node's in-edges for the
// score phase (each lane does a full C-long dot product against its own K row), the segment
// max / sum are warp-shuffle reductions with an online (running max) rescale every 32 edges,
// and for the aggregate phase lanes are spread over channels while the 32 probabilities are
// broadcast by shuffle.  Edges are a multiset: a duplicate edge simply appears twice.
#include "common.cuh"

namespace da {
namespace {

constexpr int WARPS_PER_CTA = 8;

template <int R>  // R = ceil(C / 32) channel accumulators per lane
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
attn_csr_kernel(const float* __restrict__ qkvs, int ld, const int32_t* __restrict__ rowptr,
                const int32_t* __restrict__ col, int n_targets, int H, int C, float scale,
                const float* __restrict__ resid, int ld_resid, int act, float* __restrict__ yf, int ldc,
                __nv_bfloat16* __restrict__ yhi, __nv_bfloat16* __restrict__ ylo, int ldsp,
                float* __restrict__ scores, float* __restrict__ stats, const float* __restrict__ init_acc,
                const float* __restrict__ init_stats, const int32_t* __restrict__ init_slot) {
  extern __shared__ __align__(16) float q_sm[];  // [WARPS_PER_CTA][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * WARPS_PER_CTA + warp;
  if (gw >= (long long)n_targets * H) return;  // whole warp exits together
  const int node = (int)(gw / H), head = (int)(gw % H);
  const int HC = H * C;
  float* q = q_sm + warp * C;
  const float* qrow = qkvs + (size_t)node * ld + head * C;
  for (int c = lane; c < C; c += 32) q[c] = qrow[c] * scale;
  __syncwarp();

  const float* Kbase = qkvs + HC + head * C;
  const float* Vbase = qkvs + 2 * HC + head * C;
  const int beg = rowptr[node], end = rowptr[node + 1];

  float m = -INFINITY, l = 0.f;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  if (init_slot != nullptr && init_slot[node] >= 0) {
    // continue the online softmax the dense-tile kernel started on this node's bitmap edges
    m = init_stats[((size_t)node * H + head) * 2 + 0];
    l = init_stats[((size_t)node * H + head) * 2 + 1];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = lane + 32 * r;
      if (c < C) acc[r] = init_acc[(size_t)node * HC + head * C + c];
    }
  }

  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    const bool valid = e < end;
    int j = valid ? col[e] : 0;
    float s = -INFINITY;
    if (valid) {
      const float* kr = Kbase + (size_t)j * ld;
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      if ((C & 3) == 0) {
        for (int c = 0; c < C; c += 4) {
          float4 kk = __ldg(reinterpret_cast<const float4*>(kr + c));
          float4 qq = *reinterpret_cast<const float4*>(q + c);
          d0 = fmaf(kk.x, qq.x, d0); d1 = fmaf(kk.y, qq.y, d1);
          d2 = fmaf(kk.z, qq.z, d2); d3 = fmaf(kk.w, qq.w, d3);
        }
      } else {
        for (int c = 0; c < C; ++c) d0 = fmaf(__ldg(kr + c), q[c], d0);
      }
      s = (d0 + d1) + (d2 + d3);
      if (scores) scores[(size_t)e * H + head] = s;
    }
    const float m_new = fmaxf(m, warp_max(s));
    const float p = valid ? expf(s - m_new) : 0.f;
    const float rescale = (m == -INFINITY) ? 0.f : expf(m - m_new);
    l = l * rescale + warp_sum(p);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] *= rescale;
    m = m_new;
    const int cnt = min(32, end - base);
    for (int t = 0; t < cnt; ++t) {
      const float pt = __shfl_sync(0xffffffffu, p, t);
      const int jt = __shfl_sync(0xffffffffu, j, t);
      const float* vr = Vbase + (size_t)jt * ld;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = lane + 32 * r;
        if (c < C) acc[r] = fmaf(pt, __ldg(vr + c), acc[r]);
      }
    }
  }
  if (stats && lane == 0) {
    stats[((size_t)node * H + head) * 2 + 0] = m;
    stats[((size_t)node * H + head) * 2 + 1] = l;
  }
  const float inv = 1.f / (l + 1e-16f);
  const float* srow = qkvs + (size_t)node * ld + 3 * HC + head * C;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = lane + 32 * r;
    if (c < C) {
      float v = acc[r] * inv + srow[c];
      if (resid) v += resid[(size_t)node * ld_resid + head * C + c];
      v = apply_act_rt(v, act);
      const size_t o = (size_t)head * C + c;
      if (yf) yf[(size_t)node * ldc + o] = v;
      if (yhi) {
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        yhi[(size_t)node * ldsp + o] = h;
        ylo[(size_t)node * ldsp + o] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
  }
}

__global__ void alpha_normalize_kernel(const float* __restrict__ scores, const float* __restrict__ stats,
                                       const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
                                       int n_targets, int H, float* __restrict__ alpha) {
  // one warp per target node; lanes over (edge, head) pairs
  const int lane = threadIdx.x & 31;
  const int node = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (node >= n_targets) return;
  const int beg = rowptr[node], end = rowptr[node + 1];
  const long long total = (long long)(end - beg) * H;
  for (long long idx = lane; idx < total; idx += 32) {
    const int p = beg + (int)(idx / H), h = (int)(idx % H);
    const float m = stats[((size_t)node * H + h) * 2 + 0];
    const float l = stats[((size_t)node * H + h) * 2 + 1];
    alpha[(size_t)eid[p] * H + h] = expf(scores[(size_t)p * H + h] - m) / (l + 1e-16f);
  }
}

}  // namespace

cudaError_t launch_attn_csr(const AttnCsrArgs& a, cudaStream_t s) {
  if (a.n_targets <= 0) return cudaSuccess;
  const long long warps = (long long)a.n_targets * a.H;
  const unsigned grid = (unsigned)((warps + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
  const size_t smem = (size_t)WARPS_PER_CTA * a.C * sizeof(float);
  const float scale = 1.0f / sqrtf((float)a.C);
  const int R = (a.C + 31) / 32;
#define DA_LAUNCH(RR)                                                                                       \
  attn_csr_kernel<RR><<<grid, WARPS_PER_CTA * 32, smem, s>>>(                                               \
      a.qkvs, a.ld, a.rowptr, a.col, a.n_targets, a.H, a.C, scale, a.resid, a.ld_resid, a.act, a.out.f32,   \
      a.out.ldc, a.out.hi, a.out.lo, a.out.ld_split, a.scores, a.stats, a.init_acc, a.init_stats, a.init_slot)
  if (R <= 1) DA_LAUNCH(1);
  else if (R <= 2) DA_LAUNCH(2);
  else if (R <= 5) DA_LAUNCH(5);
  else if (R <= 13) DA_LAUNCH(13);
  else return cudaErrorInvalidValue;  // head dims above 416 (resnet50 trunk) are not built
#undef DA_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_alpha_normalize(const float* scores, const float* stats, const int32_t* rowptr,
                                   const int32_t* eid, int n_targets, int H, float* alpha, cudaStream_t s) {
  if (n_targets <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)(((long long)n_targets * 32 + 255) / 256);
  alpha_normalize_kernel<<<grid, 256, 0, s>>>(scores, stats, rowptr, eid, n_targets, H, alpha);
  return cudaGetLastError();
}

}  // namespace da
