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
                const int32_t* __restrict__ col, const float* __restrict__ weight, int n_targets, int H, int C, float scale,
                const float* __restrict__ resid, int ld_resid, int act, float* __restrict__ yf, int ldc,
                __nv_bfloat16* __restrict__ yhi, __nv_bfloat16* __restrict__ ylo, int ldsp,
                float* __restrict__ scores, float* __restrict__ stats, const float* __restrict__ init_acc,
                const float* __restrict__ init_stats, const int32_t* __restrict__ init_slot,
                const int32_t* __restrict__ node_list) {
  extern __shared__ __align__(16) float q_sm[];  // [WARPS_PER_CTA][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * WARPS_PER_CTA + warp;
  if (gw >= (long long)n_targets * H) return;  // whole warp exits together
  const int node = node_list ? node_list[gw / H] : (int)(gw / H);
  const int head = (int)(gw % H);
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
    float p = valid ? expf(s - m_new) : 0.f;
    if (weight != nullptr && valid) p *= weight[e];
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

// Heavy rows (hundreds of in-edges: the virtual nodes of the Exphander wiring): one CTA owns one
// (target node, head); its 8 warps take interleaved 32-edge chunks with the same online softmax as
// attn_csr_kernel, then the 8 partial states (m, l, acc) are merged through shared memory.
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

constexpr int HEAVY_WARPS = 32;   // one 32-edge chunk per warp for rows of ~1000 in-edges

template <int R>
__global__ void __launch_bounds__(HEAVY_WARPS * 32)
attn_csr_heavy_kernel(const float* __restrict__ qkvs, int ld, const int32_t* __restrict__ rowptr,
                      const int32_t* __restrict__ col, const float* __restrict__ weight,
                      const int32_t* __restrict__ node_list, int H, int C, float scale,
                      const float* __restrict__ resid, int ld_resid, int act, float* __restrict__ yf, int ldc,
                      __nv_bfloat16* __restrict__ yhi, __nv_bfloat16* __restrict__ ylo, int ldsp,
                      const float* __restrict__ init_acc, const float* __restrict__ init_stats,
                      const int32_t* __restrict__ init_slot, const int32_t* __restrict__ img_slot,
                      const __nv_bfloat16* __restrict__ kimg, const __nv_bfloat16* __restrict__ vimg, int Cpad) {
  extern __shared__ __align__(16) float sm[];  // q[C] | acc[WARPS][C] | m[WARPS] | l[WARPS]
  float* q = sm;
  float* acc_s = sm + C;
  float* m_s = acc_s + HEAVY_WARPS * C;
  float* l_s = m_s + HEAVY_WARPS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int node = node_list[blockIdx.x / H], head = blockIdx.x % H;
  const int HC = H * C;
  const float* qrow = qkvs + (size_t)node * ld + head * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) q[c] = qrow[c] * scale;
  __syncthreads();
  const float* Kbase = qkvs + HC + head * C;
  const float* Vbase = qkvs + 2 * HC + head * C;
  const int beg = rowptr[node], end = rowptr[node + 1];
  float m = -INFINITY, l = 0.f;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  if (warp == 0 && init_slot != nullptr && init_slot[node] >= 0) {
    m = init_stats[((size_t)node * H + head) * 2 + 0];
    l = init_stats[((size_t)node * H + head) * 2 + 1];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = lane + 32 * r;
      if (c < C) acc[r] = init_acc[(size_t)node * HC + head * C + c];
    }
  }
  for (int base = beg + warp * 32; base < end; base += 32 * HEAVY_WARPS) {
    const int e = base + lane;
    const bool valid = e < end;
    const int j = valid ? col[e] : 0;
    const int sj = (img_slot != nullptr && valid) ? __ldg(img_slot + j) : -1;
    float s = -INFINITY;
    if (valid) {
      const float* kr = Kbase + (size_t)j * ld;
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      if (sj >= 0) {   // K row from the operand image: 8-channel chunks, hi plane then lo plane
        const __nv_bfloat16* kb = kimg + ((size_t)(sj >> 6) * H + head) * ((size_t)2 * 64 * Cpad) + (sj & 63) * 8;
        for (int c = 0; c < C; c += 8) {
          const uint4 hi = __ldg(reinterpret_cast<const uint4*>(kb + (size_t)(c >> 3) * 512));
          const uint4 lo = __ldg(reinterpret_cast<const uint4*>(kb + (size_t)64 * Cpad + (size_t)(c >> 3) * 512));
          const float4 q0 = *reinterpret_cast<const float4*>(q + c), q1 = *reinterpret_cast<const float4*>(q + c + 4);
          d0 = fmaf(bf_lo(hi.x) + bf_lo(lo.x), q0.x, d0); d1 = fmaf(bf_hi(hi.x) + bf_hi(lo.x), q0.y, d1);
          d2 = fmaf(bf_lo(hi.y) + bf_lo(lo.y), q0.z, d2); d3 = fmaf(bf_hi(hi.y) + bf_hi(lo.y), q0.w, d3);
          d0 = fmaf(bf_lo(hi.z) + bf_lo(lo.z), q1.x, d0); d1 = fmaf(bf_hi(hi.z) + bf_hi(lo.z), q1.y, d1);
          d2 = fmaf(bf_lo(hi.w) + bf_lo(lo.w), q1.z, d2); d3 = fmaf(bf_hi(hi.w) + bf_hi(lo.w), q1.w, d3);
        }
      } else if ((C & 3) == 0) {
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
    }
    const float m_new = fmaxf(m, warp_max(s));
    float p = valid ? expf(s - m_new) : 0.f;
    if (weight != nullptr && valid) p *= weight[e];
    const float rescale = (m == -INFINITY) ? 0.f : expf(m - m_new);
    l = l * rescale + warp_sum(p);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] *= rescale;
    m = m_new;
    // aggregate: 4 V rows in flight per step (invalid lanes carry p = 0 and row 0)
#pragma unroll 1
    for (int t = 0; t < 32; t += 4) {
      float pt[4]; const float* vr[4]; const __nv_bfloat16* vi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pt[u] = __shfl_sync(0xffffffffu, p, t + u);
        vr[u] = Vbase + (size_t)__shfl_sync(0xffffffffu, j, t + u) * ld;
        const int su = __shfl_sync(0xffffffffu, sj, t + u);   // warp-uniform: image row or fp32 row
        vi[u] = su >= 0 ? vimg + ((size_t)(su >> 6) * H + head) * ((size_t)2 * 64 * Cpad) + (su & 63) * 8 : nullptr;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = lane + 32 * r;
        if (c < C) {
          const size_t io = (size_t)(c >> 3) * 512 + (c & 7);
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            v[u] = vi[u] ? __bfloat162float(vi[u][io]) + __bfloat162float(vi[u][(size_t)64 * Cpad + io]) : __ldg(vr[u] + c);
          acc[r] = fmaf(pt[0], v[0], acc[r]); acc[r] = fmaf(pt[1], v[1], acc[r]);
          acc[r] = fmaf(pt[2], v[2], acc[r]); acc[r] = fmaf(pt[3], v[3], acc[r]);
        }
      }
    }
  }
  if (lane == 0) { m_s[warp] = m; l_s[warp] = l; }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = lane + 32 * r;
    if (c < C) acc_s[warp * C + c] = acc[r];
  }
  __syncthreads();
  if (warp != 0) return;
  // lane w holds partial w's (m, l); M = max over partials, factor f = exp(m_w - M)
  const float mw = m_s[lane], lw = l_s[lane];
  const float M = warp_max(mw);
  const float fw = (mw == -INFINITY) ? 0.f : expf(mw - M);
  const float L = warp_sum(lw * fw);
  const float inv = 1.f / (L + 1e-16f);
  const float* srow = qkvs + (size_t)node * ld + 3 * HC + head * C;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = lane + 32 * r;
    float a = 0.f;
#pragma unroll 8
    for (int w = 0; w < HEAVY_WARPS; ++w) {   // shuffles are executed by every lane (c may be out of range)
      const float f = __shfl_sync(0xffffffffu, fw, w);
      if (c < C) a = fmaf(acc_s[w * C + c], f, a);
    }
    if (c < C) {
      float v = a * inv + srow[c];
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

// Lane-per-edge variant for 32-channel heads (the hidden layers): one CTA owns one (target, head) of a row that is
// NOT part of a dense tile (virtual nodes: a few hub rows with ~1000 in-edges, many rows with a handful).  Each
// lane fetches the whole K row and V row of ITS edge up front (16 independent 16-byte loads: one memory latency per
// 32-edge chunk instead of one per 4 edges), computes its score alone, and the 32 weighted V rows of a chunk are
// summed by a transposing butterfly (31 shuffles) that leaves channel c in lane c.
constexpr int VROW_WARPS = 8;

__device__ __forceinline__ void load_row32(const float* __restrict__ fp32_row, const __nv_bfloat16* __restrict__ img_row,
                                           size_t lo_off, float* out) {
  if (img_row != nullptr) {   // split-bf16 operand image: 4 chunks of 8 channels, 512 elements apart; lo plane at lo_off
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(img_row + (size_t)ch * 512));
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(img_row + lo_off + (size_t)ch * 512));
      out[8 * ch + 0] = bf_lo(hi.x) + bf_lo(lo.x); out[8 * ch + 1] = bf_hi(hi.x) + bf_hi(lo.x);
      out[8 * ch + 2] = bf_lo(hi.y) + bf_lo(lo.y); out[8 * ch + 3] = bf_hi(hi.y) + bf_hi(lo.y);
      out[8 * ch + 4] = bf_lo(hi.z) + bf_lo(lo.z); out[8 * ch + 5] = bf_hi(hi.z) + bf_hi(lo.z);
      out[8 * ch + 6] = bf_lo(hi.w) + bf_lo(lo.w); out[8 * ch + 7] = bf_hi(hi.w) + bf_hi(lo.w);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(fp32_row + c));
      out[c] = t.x; out[c + 1] = t.y; out[c + 2] = t.z; out[c + 3] = t.w;
    }
  }
}

// Launch shape: the first n_coop rows of node_list (hubs with hundreds of in-edges) get one CTA per (row, head), its
// 8 warps sharing the edge list; the remaining rows (a handful of in-edges) get one WARP per (row, head), 8 pairs
// per CTA -- 224 instead of 1792 CTAs for the light virtual rows of the c3 batch.
__global__ void __launch_bounds__(VROW_WARPS * 32)
attn_csr_vrow32_kernel(const float* __restrict__ qkvs, int ld, const int32_t* __restrict__ rowptr,
                       const int32_t* __restrict__ col, const float* __restrict__ weight,
                       const int32_t* __restrict__ node_list, int n_rows, int n_coop, int H, float scale,
                       const float* __restrict__ resid, int ld_resid, int act, float* __restrict__ yf, int ldc,
                       __nv_bfloat16* __restrict__ yhi, __nv_bfloat16* __restrict__ ylo, int ldsp,
                       const int32_t* __restrict__ img_slot, const __nv_bfloat16* __restrict__ kimg,
                       const __nv_bfloat16* __restrict__ vimg, int Cpad,
                       int n_vrow_ctas, int gx_n, const int32_t* __restrict__ gx_src, const int32_t* __restrict__ gx_slot,
                       __nv_bfloat16* __restrict__ gx_kimg, __nv_bfloat16* __restrict__ gx_vimg) {
  constexpr int C = 32;
  __shared__ float acc_s[VROW_WARPS][C];
  __shared__ float m_s[VROW_WARPS], l_s[VROW_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_trigger();
  pdl_wait();   // everything below reads the projection GEMM's output
  if ((int)blockIdx.x >= n_vrow_ctas) {
    // Riders of the same launch (persistent dense path, api.cu): the per-layer gather of the planner's promoted extra
    // sources -- fp32 K / V rows -> split-bf16 rows of the padding area of the operand images (gather_extra_kernel of
    // attn_dense.cu, 32-channel heads).  One warp per (entry, K | V); it touches image rows no CSR row reads.
    const int wg = ((int)blockIdx.x - n_vrow_ctas) * VROW_WARPS + warp;
    if (wg >= gx_n * 2) return;
    const int ent = wg >> 1, part = 1 + (wg & 1);
    const int node = __ldg(gx_src + ent), slot = __ldg(gx_slot + ent);
    const int blk = slot >> 6, rb = slot & 63;
    const float* row = qkvs + (size_t)node * ld + part * H * C;
    __nv_bfloat16* img = part == 1 ? gx_kimg : gx_vimg;
    for (int it = lane; it < H * 4; it += 32) {
      const int h = it >> 2, ch = it & 3;
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(row + h * C + ch * 8));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(row + h * C + ch * 8 + 4));
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      __nv_bfloat16 hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        hi[e] = __float2bfloat16_rn(v[e]);
        lo[e] = __float2bfloat16_rn(v[e] - __bfloat162float(hi[e]));
      }
      __nv_bfloat16* base = img + ((size_t)blk * H + h) * ((size_t)2 * 64 * C);
      const size_t off = (size_t)ch * (64 * 8) + (size_t)rb * 8;
      *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(base + (size_t)64 * C + off) = *reinterpret_cast<uint4*>(lo);
    }
    return;
  }
  const bool coop = (int)blockIdx.x < n_coop * H;
  int pair = coop ? (int)blockIdx.x : n_coop * H + ((int)blockIdx.x - n_coop * H) * VROW_WARPS + warp;
  if (pair >= n_rows * H) return;   // (warp mode only: whole warps leave, no barrier follows for them)
  const int node = node_list[pair / H], head = pair % H;
  const int HC = H * C;
  const int wstart = coop ? warp : 0, wstep = coop ? VROW_WARPS : 1;
  float q[C];
  {
    const float* qrow = qkvs + (size_t)node * ld + head * C;
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(qrow + c));
      q[c] = t.x * scale; q[c + 1] = t.y * scale; q[c + 2] = t.z * scale; q[c + 3] = t.w * scale;
    }
  }
  const int beg = rowptr[node], end = rowptr[node + 1];
  const size_t lo_off = (size_t)64 * Cpad;
  float m = -INFINITY, l = 0.f, acc = 0.f;   // acc: channel `lane` of this warp's partial sum
  // edge indices run one chunk ahead of the K / V rows (col -> img_slot -> row address is a chain of dependent loads)
  auto fetch_edge = [&](int base, int& j_, float& w_, int& sj_, bool& valid_) {
    const int e = base + lane;
    valid_ = e < end;
    j_ = valid_ ? __ldg(col + e) : node;
    w_ = (valid_ && weight != nullptr) ? __ldg(weight + e) : 1.f;
    sj_ = (img_slot != nullptr) ? __ldg(img_slot + j_) : -1;
  };
  int j_n, sj_n; float w_n; bool valid_n;
  fetch_edge(beg + wstart * 32, j_n, w_n, sj_n, valid_n);
  for (int base = beg + wstart * 32; base < end; base += 32 * wstep) {
    const int j = j_n, sj = sj_n; const float w = w_n; const bool valid = valid_n;
    if (base + 32 * wstep < end) fetch_edge(base + 32 * wstep, j_n, w_n, sj_n, valid_n);
    const __nv_bfloat16* kb = sj >= 0 ? kimg + ((size_t)(sj >> 6) * H + head) * ((size_t)2 * 64 * Cpad) + (sj & 63) * 8 : nullptr;
    const __nv_bfloat16* vb = sj >= 0 ? vimg + ((size_t)(sj >> 6) * H + head) * ((size_t)2 * 64 * Cpad) + (sj & 63) * 8 : nullptr;
    float kf[C], vf[C];
    load_row32(qkvs + (size_t)j * ld + HC + head * C, kb, lo_off, kf);
    load_row32(qkvs + (size_t)j * ld + 2 * HC + head * C, vb, lo_off, vf);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      d0 = fmaf(kf[c], q[c], d0); d1 = fmaf(kf[c + 1], q[c + 1], d1);
      d2 = fmaf(kf[c + 2], q[c + 2], d2); d3 = fmaf(kf[c + 3], q[c + 3], d3);
    }
    const float s = valid ? (d0 + d1) + (d2 + d3) : -INFINITY;
    const float m_new = fmaxf(m, warp_max(s));
    const float p = valid ? w * expf(s - m_new) : 0.f;
    const float rescale = (m == -INFINITY) ? 0.f : expf(m - m_new);
    l = l * rescale + warp_sum(p);
    m = m_new;
    // sum over the 32 lanes of p * v[c], channel c ending up in lane c: halve the vector at every step
#pragma unroll
    for (int c = 0; c < C; ++c) vf[c] *= p;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float keep = upper ? vf[i + off] : vf[i];
        const float give = upper ? vf[i] : vf[i + off];
        vf[i] = keep + __shfl_xor_sync(0xffffffffu, give, off);
      }
    }
    acc = fmaf(acc, rescale, vf[0]);
  }
  float a, inv;
  if (coop) {   // merge the 8 partial states of this (row, head)
    if (lane == 0) { m_s[warp] = m; l_s[warp] = l; }
    acc_s[warp][lane] = acc;
    __syncthreads();
    if (warp != 0) return;
    const float mw = lane < VROW_WARPS ? m_s[lane] : -INFINITY, lw = lane < VROW_WARPS ? l_s[lane] : 0.f;
    const float M = warp_max(mw);
    const float fw = (mw == -INFINITY) ? 0.f : expf(mw - M);
    const float L = warp_sum(lw * fw);
    inv = 1.f / (L + 1e-16f);
    a = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < VROW_WARPS; ++w2) a = fmaf(acc_s[w2][lane], __shfl_sync(0xffffffffu, fw, w2), a);
  } else {
    a = acc;
    inv = 1.f / (l + 1e-16f);
  }
  float v = a * inv + __ldg(qkvs + (size_t)node * ld + 3 * HC + head * C + lane);
  if (resid) v += resid[(size_t)node * ld_resid + head * C + lane];
  v = apply_act_rt(v, act);
  const size_t o = (size_t)head * C + lane;
  if (yf) yf[(size_t)node * ldc + o] = v;
  if (yhi) {
    const __nv_bfloat16 hb = __float2bfloat16_rn(v);
    yhi[(size_t)node * ldsp + o] = hb;
    ylo[(size_t)node * ldsp + o] = __float2bfloat16_rn(v - __bfloat162float(hb));
  }
}

// Row-parallel variant for low-degree targets (the residual edges of a dense-tile plan: typically one
// virtual->real edge per node).  One warp owns one target node and ALL heads: lane l holds the VPL =
// H*C/32 contiguous channels [l*VPL, (l+1)*VPL) of the row, which belong to head l / (32/H); every
// row access (Q, K_j, V_j, skip, residual, dense accumulator, output) is a fully coalesced H*C-wide
// read or write, and the per-head score is a shuffle reduction over the 32/H lanes of that head.
template <int VPL>
__global__ void __launch_bounds__(256)
attn_csr_rows_kernel(const float* __restrict__ qkvs, int ld, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ col, const float* __restrict__ weight,
                     const int32_t* __restrict__ node_list, int n_list, int H, int C, float scale,
                     const float* __restrict__ resid, int ld_resid, int act, float* __restrict__ yf, int ldc,
                     __nv_bfloat16* __restrict__ yhi, __nv_bfloat16* __restrict__ ylo, int ldsp,
                     const float* __restrict__ init_acc, const float* __restrict__ init_stats,
                     const int32_t* __restrict__ init_slot, int wpn) {
  // wpn warps share one node: each owns 32 * VPL contiguous channels (a whole number of heads)
  constexpr int W = (VPL % 4 == 0) ? 4 : ((VPL % 2 == 0) ? 2 : 1);
  const int lane = threadIdx.x & 31;
  const int wg = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (wg >= n_list * wpn) return;
  const int node = node_list ? node_list[wg / wpn] : wg / wpn;
  const int HC = H * C, lph = C / VPL;
  const int c0 = (wg % wpn) * (32 * VPL) + lane * VPL;  // first channel (within the H*C row) owned by this lane
  const int head = c0 / C;

  auto load_row = [&](const float* p, float* dst) {
#pragma unroll
    for (int i = 0; i < VPL; i += W) {
      if (W == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(p + i)); dst[i] = t.x; dst[i + 1] = t.y; dst[i + 2] = t.z; dst[i + 3] = t.w; }
      else if (W == 2) { float2 t = __ldg(reinterpret_cast<const float2*>(p + i)); dst[i] = t.x; dst[i + 1] = t.y; }
      else dst[i] = __ldg(p + i);
    }
  };
  float q[VPL], acc[VPL];
  const int beg = rowptr[node], end = rowptr[node + 1];
  if (end > beg) {   // rows without residual edges never read their fp32 Q (the GEMM may not even have stored it)
    load_row(qkvs + (size_t)node * ld + c0, q);
#pragma unroll
    for (int i = 0; i < VPL; ++i) q[i] *= scale;
  }
  float m = -INFINITY, l = 0.f;
  if (init_slot != nullptr && init_slot[node] >= 0) {
    m = init_stats[((size_t)node * H + head) * 2 + 0];
    l = init_stats[((size_t)node * H + head) * 2 + 1];
    load_row(init_acc + (size_t)node * HC + c0, acc);
  } else {
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i] = 0.f;
  }
  for (int e = beg; e < end; ++e) {
    const int j = col[e];
    const float w = weight ? weight[e] : 1.f;
    float kv[VPL];
    load_row(qkvs + (size_t)j * ld + HC + c0, kv);
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) d = fmaf(q[i], kv[i], d);
    for (int o = lph >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    load_row(qkvs + (size_t)j * ld + 2 * HC + c0, kv);
    const float m_new = fmaxf(m, d);
    const float sc = (m == -INFINITY) ? 0.f : expf(m - m_new);
    const float p = w * expf(d - m_new);
    l = l * sc + p;
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i] = fmaf(p, kv[i], acc[i] * sc);
    m = m_new;
  }
  const float inv = 1.f / (l + 1e-16f);
  float sk[VPL];
  load_row(qkvs + (size_t)node * ld + 3 * HC + c0, sk);
#pragma unroll
  for (int i = 0; i < VPL; ++i) sk[i] = fmaf(acc[i], inv, sk[i]);
  if (resid) {
    float rr[VPL];
    load_row(resid + (size_t)node * ld_resid + c0, rr);
#pragma unroll
    for (int i = 0; i < VPL; ++i) sk[i] += rr[i];
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) sk[i] = apply_act_rt(sk[i], act);
  if (yf) {
    float* dst = yf + (size_t)node * ldc + c0;
#pragma unroll
    for (int i = 0; i < VPL; i += W) {
      if (W == 4) *reinterpret_cast<float4*>(dst + i) = make_float4(sk[i], sk[i + 1], sk[i + 2], sk[i + 3]);
      else if (W == 2) *reinterpret_cast<float2*>(dst + i) = make_float2(sk[i], sk[i + 1]);
      else dst[i] = sk[i];
    }
  }
  if (yhi) {
    __nv_bfloat16* dh = yhi + (size_t)node * ldsp + c0;
    __nv_bfloat16* dl = ylo + (size_t)node * ldsp + c0;
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
      __nv_bfloat162 h2, l2;
      h2.x = __float2bfloat16_rn(sk[i]); h2.y = __float2bfloat16_rn(sk[i + 1]);
      l2.x = __float2bfloat16_rn(sk[i] - __bfloat162float(h2.x)); l2.y = __float2bfloat16_rn(sk[i + 1] - __bfloat162float(h2.y));
      *reinterpret_cast<__nv_bfloat162*>(dh + i) = h2;
      *reinterpret_cast<__nv_bfloat162*>(dl + i) = l2;
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
      a.qkvs, a.ld, a.rowptr, a.col, a.weight, a.n_targets, a.H, a.C, scale, a.resid, a.ld_resid, a.act, a.out.f32,   \
      a.out.ldc, a.out.hi, a.out.lo, a.out.ld_split, a.scores, a.stats, a.init_acc, a.init_stats, a.init_slot,   \
      a.node_list)
  if (R <= 1) DA_LAUNCH(1);
  else if (R <= 2) DA_LAUNCH(2);
  else if (R <= 5) DA_LAUNCH(5);
  else if (R <= 13) DA_LAUNCH(13);
  else return cudaErrorInvalidValue;  // head dims above 416 (resnet50 trunk) are not built
#undef DA_LAUNCH
  return cudaGetLastError();
}

bool attn_csr_vrows_supported(int H, int C) { return C == 32 && H > 0; }

cudaError_t launch_attn_csr_vrows(const AttnCsrArgs& a, cudaStream_t s) {
  if (a.n_targets <= 0) return cudaSuccess;
  if (!a.node_list || a.scores || a.stats || a.init_acc || a.C != 32 || (a.ld & 3)) return cudaErrorInvalidValue;
  if (a.img_slot != nullptr && (a.kimg == nullptr || a.vimg == nullptr || a.img_Cpad != 32)) return cudaErrorInvalidValue;
  if (a.gx_n > 0 && (!a.gx_src || !a.gx_slot || !a.gx_kimg || !a.gx_vimg)) return cudaErrorInvalidValue;
  const float scale = 1.0f / sqrtf((float)a.C);
  const int n_coop = a.n_coop < a.n_targets ? a.n_coop : a.n_targets;
  const int light_pairs = (a.n_targets - n_coop) * a.H;
  const int n_vrow_ctas = n_coop * a.H + (light_pairs + VROW_WARPS - 1) / VROW_WARPS;
  const unsigned grid = (unsigned)(n_vrow_ctas + (a.gx_n * 2 + VROW_WARPS - 1) / VROW_WARPS);
  return launch_pdl(attn_csr_vrow32_kernel, dim3(grid), dim3(VROW_WARPS * 32), 0, s,
                    a.qkvs, a.ld, a.rowptr, a.col, a.weight, a.node_list, a.n_targets, n_coop, a.H, scale, a.resid, a.ld_resid, a.act,
                    a.out.f32, a.out.ldc, a.out.hi, a.out.lo, a.out.ld_split, a.img_slot, a.kimg, a.vimg, a.img_Cpad, n_vrow_ctas,
                    a.gx_n, a.gx_src, a.gx_slot, a.gx_kimg, a.gx_vimg);
}

cudaError_t launch_attn_csr_heavy(const AttnCsrArgs& a, cudaStream_t s) {
  if (a.n_targets <= 0) return cudaSuccess;
  if (!a.node_list || a.scores || a.stats) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(a.C + HEAVY_WARPS * a.C + 2 * HEAVY_WARPS) * sizeof(float);
  const float scale = 1.0f / sqrtf((float)a.C);
  const int R = (a.C + 31) / 32;
  const unsigned grid = (unsigned)a.n_targets * a.H;
#define DA_LAUNCH(RR)                                                                                          \
  attn_csr_heavy_kernel<RR><<<grid, HEAVY_WARPS * 32, smem, s>>>(                                            \
      a.qkvs, a.ld, a.rowptr, a.col, a.weight, a.node_list, a.H, a.C, scale, a.resid, a.ld_resid, a.act,       \
      a.out.f32, a.out.ldc, a.out.hi, a.out.lo, a.out.ld_split, a.init_acc, a.init_stats, a.init_slot,        \
      a.img_slot, a.kimg, a.vimg, a.img_Cpad)
  if (a.img_slot != nullptr && (a.C % 8 != 0 || a.kimg == nullptr || a.vimg == nullptr)) return cudaErrorInvalidValue;
  if (R <= 1) DA_LAUNCH(1);
  else if (R <= 2) DA_LAUNCH(2);
  else if (R <= 5) DA_LAUNCH(5);
  else if (R <= 13) DA_LAUNCH(13);
  else return cudaErrorInvalidValue;
#undef DA_LAUNCH
  return cudaGetLastError();
}

namespace {
// channels per lane of the row kernel: the whole H*C row over one warp when that is <= 8 values per lane,
// otherwise half a row per warp (two warps per node) -- more warps in flight for the 1152-wide last layer
int rows_vpl(int H, int C) {
  const int HC = H * C;
  if (HC % 32) return 0;
  int vpl = HC / 32;
  if (vpl > 8) { if (HC % 64) return 0; vpl = HC / 64; }
  if (!(vpl == 8 || vpl == 18 || vpl == 6)) return 0;
  if (C % vpl) return 0;
  const int lph = C / vpl;
  if (lph < 1 || lph > 32 || (lph & (lph - 1))) return 0;
  return vpl;
}
}  // namespace

bool attn_csr_rows_supported(int H, int C) { return rows_vpl(H, C) != 0; }

cudaError_t launch_attn_csr_rows(const AttnCsrArgs& a, cudaStream_t s) {
  if (a.n_targets <= 0) return cudaSuccess;
  const int vpl = rows_vpl(a.H, a.C);
  if (!vpl || a.scores || a.stats) return cudaErrorInvalidValue;
  const int wpn = a.H * a.C / (32 * vpl);
  const unsigned grid = (unsigned)(((long long)a.n_targets * wpn * 32 + 255) / 256);
  const float scale = 1.0f / sqrtf((float)a.C);
#define DA_LAUNCH(V)                                                                                              \
  attn_csr_rows_kernel<V><<<grid, 256, 0, s>>>(a.qkvs, a.ld, a.rowptr, a.col, a.weight, a.node_list, a.n_targets, a.H, \
                                               a.C, scale, a.resid, a.ld_resid, a.act, a.out.f32, a.out.ldc, a.out.hi, \
                                               a.out.lo, a.out.ld_split, a.init_acc, a.init_stats, a.init_slot, wpn)
  if (vpl == 8) DA_LAUNCH(8);
  else if (vpl == 18) DA_LAUNCH(18);
  else DA_LAUNCH(6);
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
