// Inference-time weight folding of the 2-D denoiser (efficient_gat.py:121-146) and the node-wise head that goes with it.
//
// Two places of Eff_GAT.forward_with_feats are a linear layer followed DIRECTLY by another linear map:
//   * combined = mlp[2](h) (efficient_gat.py:135, Linear(128, D), no activation) feeds the four projections of the first
//     TransformerConv (Transformer_GNN.py:33 / exophormer_gnn.py:203) and the trunk residual:
//         [Q|K|V|skip]_0 = combined W_0^T + b_0 = h (W_0 W_2)^T + (W_0 b_2 + b_0)                 K = 128 instead of 1152
//   * the last TransformerConv's output only enters final_mlp[0] = Linear(D, 32) (efficient_gat.py:143-145):
//         final_mlp[0](attn + skip + combined) = sum_heads alpha_h (x3 (W_a^h W_v^h)^T + W_a^h b_v^h)     V' has 32 channels per head
//                                               + x3 (W_a W_skip)^T + h (W_a W_2)^T + (b_a + W_a (b_skip + b_2))
//     so the aggregation runs on 32-channel values instead of 144-channel ones and neither the skip projection,
//     nor `combined`, nor the [M, D] residual sum ever exist.
// The products are formed once per da_load_weights in fp64 on the device and rounded to fp32 (then split to bf16 hi / lo
// like every other weight).  Virtual-node rows (exophormer_gnn.py:169-178) enter the first projection through one-hot
// columns appended to h: column 128 + v of the folded weight holds W_0 (virt_emb[v] - b_2).
#include <type_traits>

#include "common.cuh"

namespace da {
namespace {

// out[i * ldo_r + j * ldo_c] = sum_t A[i * lda + t] * B[t * ldb_r + j * ldb_c] + add_scale * add[i * add_si + j * add_sj]
__global__ void matmul_f64_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb_r, int ldb_c,
                                  const float* add, int add_si, int add_sj, float add_scale, float* out, int ldo_r, int ldo_c,
                                  int m, int n, int k) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n || i >= m) return;
  const float* a = A + (size_t)i * lda;
  const float* b = B + (size_t)j * ldb_c;
  double s = 0.0;
  for (int t = 0; t < k; ++t) s = fma((double)__ldg(a + t), (double)__ldg(b + (size_t)t * ldb_r), s);
  if (add) s += (double)add_scale * (double)add[(size_t)i * add_si + (size_t)j * add_sj];
  out[(size_t)i * ldo_r + (size_t)j * ldo_c] = (float)s;
}

__global__ void onehot_rows_kernel(__nv_bfloat16* __restrict__ hi, int ld, int col0, const int32_t* __restrict__ ids, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) hi[(size_t)r * ld + col0 + ids[r]] = __float2bfloat16_rn(1.0f);
}

constexpr int HF_NT = 256, HF_ROWS = 16, HF_NB = (HF_NT / 32) * HF_ROWS, HF_PITCH = 33;

// D (16 x 8, fp32) += A (16 x 16 bf16, row-major) * B (16 x 8 bf16, "col-major" = [n][k]); warp-level legacy tensor-core
// path -- this 0.7 GFLOP node-wise product is far too small for a TMA / tcgen05 pipeline of its own
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One warp = 16 nodes x all 32 channels of final_mlp[0]; K = Hm + hid input channels read straight from the split-bf16
// planes of h and x3 (the row-major A fragments of mma.sync), weights staged once per CTA in shared memory as
// [32][K + 8] bf16 hi / lo (the pad makes the B-fragment reads bank-conflict free).  Three MMAs per (k-step, n-tile):
// hi*hi + hi*lo + lo*hi, as everywhere else.  Then the per-head attention aggregates and the bias join, GELU,
// final_mlp[2] and the sampler update (spatial_diffusion.py:485-627) as in head_final_2d_kernel.
__global__ void __launch_bounds__(HF_NT)
head_fold_kernel(HeadFoldArgs a) {
  extern __shared__ __align__(16) uint8_t hf_sm[];
  pdl_trigger();
  const int Kt = a.Hm + a.hid, KP = Kt + 8;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(hf_sm);            // [32][KP]
  __nv_bfloat16* wl = wh + (size_t)32 * KP;
  float* u_all = reinterpret_cast<float*>(wl + (size_t)32 * KP);          // [warps][16][HF_PITCH]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cpr = Kt >> 3;   // 16-byte chunks per weight row
  for (int i = tid; i < 32 * cpr; i += HF_NT) {
    const int row = i / cpr, c8 = (i - row * cpr) << 3;
    *reinterpret_cast<uint4*>(wh + (size_t)row * KP + c8) = __ldg(reinterpret_cast<const uint4*>(a.w_hi + (size_t)row * Kt + c8));
    *reinterpret_cast<uint4*>(wl + (size_t)row * KP + c8) = __ldg(reinterpret_cast<const uint4*>(a.w_lo + (size_t)row * Kt + c8));
  }
  __syncthreads();
  pdl_wait();   // the weight stage above is constant data; h, x_3 and the per-head aggregates are the predecessors' outputs
  const HeadFinalArgs& f = a.fin;
  const int g = lane >> 2, t = lane & 3;
  const int node0 = blockIdx.x * HF_NB + warp * HF_ROWS;
  const int r0 = node0 + g, r1 = r0 + 8;
  const bool v0 = r0 < f.M, v1 = r1 < f.M;
  float acc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
  // A fragments of KB k-steps are requested together (KB x 8 independent 4-byte loads per thread) before their MMAs:
  // the loop is pure load latency otherwise.  KB = 8 when both inputs are multiples of 128 columns wide, else 1.
  auto k_block = [&](const int k_base, auto kb_tag) {
    constexpr int KB = decltype(kb_tag)::value;
    const bool from_h = k_base < a.Hm;
    const __nv_bfloat16* ph = from_h ? a.h_hi : a.x_hi;
    const __nv_bfloat16* pl = from_h ? a.h_lo : a.x_lo;
    const int ld = from_h ? a.ld_h : a.ld_x, kk0 = (from_h ? k_base : k_base - a.Hm) + 2 * t;
    uint32_t ah[KB][4], al[KB][4];
#pragma unroll
    for (int s_ = 0; s_ < KB; ++s_) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { ah[s_][e] = 0u; al[s_][e] = 0u; }
      if (v0) {
        const size_t o = (size_t)r0 * ld + kk0 + 16 * s_;
        ah[s_][0] = __ldg(reinterpret_cast<const uint32_t*>(ph + o)); ah[s_][2] = __ldg(reinterpret_cast<const uint32_t*>(ph + o + 8));
        al[s_][0] = __ldg(reinterpret_cast<const uint32_t*>(pl + o)); al[s_][2] = __ldg(reinterpret_cast<const uint32_t*>(pl + o + 8));
      }
      if (v1) {
        const size_t o = (size_t)r1 * ld + kk0 + 16 * s_;
        ah[s_][1] = __ldg(reinterpret_cast<const uint32_t*>(ph + o)); ah[s_][3] = __ldg(reinterpret_cast<const uint32_t*>(ph + o + 8));
        al[s_][1] = __ldg(reinterpret_cast<const uint32_t*>(pl + o)); al[s_][3] = __ldg(reinterpret_cast<const uint32_t*>(pl + o + 8));
      }
    }
#pragma unroll
    for (int s_ = 0; s_ < KB; ++s_) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const size_t wo = (size_t)(nt * 8 + g) * KP + k_base + 16 * s_ + 2 * t;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(wh + wo), bh1 = *reinterpret_cast<const uint32_t*>(wh + wo + 8);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(wl + wo), bl1 = *reinterpret_cast<const uint32_t*>(wl + wo + 8);
        mma_bf16_16816(acc[nt], ah[s_], bh0, bh1);
        mma_bf16_16816(acc[nt], ah[s_], bl0, bl1);
        mma_bf16_16816(acc[nt], al[s_], bh0, bh1);
      }
    }
  };
  if (((a.Hm | a.hid) & 127) == 0) {
    for (int k0 = 0; k0 < Kt; k0 += 128) k_block(k0, std::integral_constant<int, 8>{});
  } else {
    for (int k0 = 0; k0 < Kt; k0 += 16) k_block(k0, std::integral_constant<int, 1>{});
  }
  // + per-head aggregates + bias, GELU -> u (this warp's 16 x 32 tile in shared memory)
  float* u_s = u_all + (size_t)warp * HF_ROWS * HF_PITCH;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int node = half ? r1 : r0;
    const bool ok = half ? v1 : v0;
    // the per-head aggregates of this row: all H x 4 float2 words requested together (heads in groups of 8), summed in
    // head order afterwards -- one memory latency instead of H
    float2 ps[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    if (ok) {
      const float* pp = a.partial + (size_t)node * a.H * 32 + 2 * t;
      for (int h0 = 0; h0 < a.H; h0 += 8) {
        float2 p2[8][4];
#pragma unroll
        for (int hh = 0; hh < 8; ++hh)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            p2[hh][nt] = (h0 + hh < a.H) ? __ldg(reinterpret_cast<const float2*>(pp + (h0 + hh) * 32 + nt * 8)) : make_float2(0.f, 0.f);
#pragma unroll
        for (int hh = 0; hh < 8; ++hh)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) { ps[nt].x += p2[hh][nt].x; ps[nt].y += p2[hh][nt].y; }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = nt * 8 + 2 * t;
      const float s0 = acc[nt][2 * half] + ps[nt].x, s1 = acc[nt][2 * half + 1] + ps[nt].y;
      u_s[(g + 8 * half) * HF_PITCH + col] = gelu_erf(s0 + __ldg(a.bias + col));
      u_s[(g + 8 * half) * HF_PITCH + col + 1] = gelu_erf(s1 + __ldg(a.bias + col + 1));
    }
  }
  __syncwarp();
  // final_mlp[2] + sampler update: lane = (node of the tile, output parity)
  const int nl = lane & 15, node = node0 + nl;
  if (node < f.M) {
    const int ext = f.row_ext ? __ldg(f.row_ext + node) : node;
    da_step_coef cf = f.coef;
    if (f.tabs.t != nullptr && f.step_mode != STEP_NONE) cf = node_coef(f.coef, f.tabs, ext);
    for (int c = lane >> 4; c < f.C_out; c += 2) {
      float sacc = __ldg(f.b_b + c);
#pragma unroll
      for (int k = 0; k < 32; ++k) sacc = fmaf(__ldg(f.w_b + c * 32 + k), u_s[nl * HF_PITCH + k], sacc);
      const int eidx = ext * f.C_out + c;
      float x = 0.f, nz = 0.f;
      if (f.step_mode != STEP_NONE) {
        x = f.x_in[eidx];
        if (f.noise) nz = f.noise[eidx];
      }
      f.out[eidx] = step_update(f.step_mode, x, sacc, nz, cf);
    }
  }
}

}  // namespace

cudaError_t launch_matmul_f64(const float* A, int lda, const float* B, int ldb_r, int ldb_c, const float* add, int add_si,
                              int add_sj, float add_scale, float* out, int ldo_r, int ldo_c, int m, int n, int k, cudaStream_t s) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  dim3 grid((n + 127) / 128, m);
  matmul_f64_kernel<<<grid, 128, 0, s>>>(A, lda, B, ldb_r, ldb_c, add, add_si, add_sj, add_scale, out, ldo_r, ldo_c, m, n, k);
  return cudaGetLastError();
}

cudaError_t launch_onehot_rows(__nv_bfloat16* hi, int ld, int col0, const int32_t* ids, int rows, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  onehot_rows_kernel<<<(rows + 127) / 128, 128, 0, s>>>(hi, ld, col0, ids, rows);
  return cudaGetLastError();
}

cudaError_t launch_head_fold(const HeadFoldArgs& a, cudaStream_t s) {
  if (a.fin.M <= 0) return cudaSuccess;
  const int Kt = a.Hm + a.hid;
  if ((a.Hm & 15) || (a.hid & 15) || (a.ld_h & 1) || (a.ld_x & 1) || a.fin.Nh != 32 || a.fin.head_kind != DA_HEAD_2D) return cudaErrorInvalidValue;
  const size_t smem = (size_t)2 * 32 * (Kt + 8) * sizeof(__nv_bfloat16) + (size_t)(HF_NT / 32) * HF_ROWS * HF_PITCH * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(head_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  return launch_pdl(head_fold_kernel, dim3((a.fin.M + HF_NB - 1) / HF_NB), dim3(HF_NT), smem, s, a);
}

}  // namespace da
