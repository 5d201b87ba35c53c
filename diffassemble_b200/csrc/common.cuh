// Shared declarations for the diffassemble_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/diffassemble_b200.h"

namespace da {

// ---- programmatic dependent launch (PDL) along the kernel chain of one denoising step ------------------------------------
// Every kernel of the chain calls pdl_trigger() first (its dependents may be scheduled as soon as all of ITS CTAs are
// running) and pdl_wait() before it touches anything a predecessor wrote or a predecessor may still read -- after its own
// set-up (barrier init, TMEM allocation, tensor-map prefetch, staging of constant weights), which thereby runs under the
// previous kernel's tail.  Both are no-ops for launches without the attribute.  Transitivity: kernel N+2 waits for N+1,
// which cannot complete before its own wait for N has returned -- so every kernel waits before any early return.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef DA_PDL_EARLY_TRIGGER
#define DA_PDL_EARLY_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_trigger() {
#if DA_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif
bool pdl_enabled();   // api.cu: false with DA_NO_PDL=1
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}


enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_LRELU = 2, ACT_RELU = 3, ACT_SILU = 4, ACT_SIGMOID = 5 };

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() / F.gelu default = exact erf form (efficient_gat.py:89,95,100; Transformer_GNN.py:36)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float lrelu02(float x) { return x > 0.f ? x : 0.2f * x; }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }   // x * sigmoid(x) (EfficientNet, scope row N4)

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == ACT_GELU) return gelu_erf(x);
  if (ACT == ACT_LRELU) return lrelu02(x);
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == ACT_SILU) return silu_f(x);
  if (ACT == ACT_SIGMOID) return sigmoid_f(x);
  return x;
}
__device__ __forceinline__ float apply_act_rt(float x, int act) {
  if (act == ACT_GELU) return gelu_erf(x);
  if (act == ACT_LRELU) return lrelu02(x);
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  if (act == ACT_SILU) return silu_f(x);
  if (act == ACT_SIGMOID) return sigmoid_f(x);
  return x;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Host-side launch wrappers (each returns cudaError_t from the launch; all async on `stream`).
// ---------------------------------------------------------------------------------------------

// Destination of a linear layer's epilogue.  Any subset of the three outputs may be requested.
struct LinearOut {
  float* f32 = nullptr;          // [M, ldc] fp32
  int ldc = 0;
  __nv_bfloat16* hi = nullptr;   // split-bf16 planes [M, ld_split] (operand of the next tensor-core GEMM)
  __nv_bfloat16* lo = nullptr;
  int ld_split = 0;
  // optional: the output columns are [Q | K | V | skip] (each H*C wide) and Q / K / V are ALSO written as
  // the split-bf16 operand images of the dense-tile attention (attn_dense.cu), fusing pack_images_kernel
  // into the GEMM epilogue.  node_slot[row] >= 0 selects rows that belong to a dense tile.
  const int32_t* img_node_slot = nullptr;
  __nv_bfloat16* qimg = nullptr;
  __nv_bfloat16* kimg = nullptr;
  __nv_bfloat16* vimg = nullptr;
  int img_H = 0, img_C = 0, img_Cpad = 0, img_rows = 0;   // img_rows: rows [0, img_rows) have a node_slot entry
  // optional per-128-row-tile flags (bit 0: some row's fp32 Q is read later, bit 1: some row's fp32 K / V is):
  // tiles whose Q / K / V only feed the tensor-core attention skip those fp32 stores (HBM-write-bound GEMMs).
  const uint8_t* f32_tile_flags = nullptr;
  // optional (N == 32 tiles only): the 2-D head's last linear + sampler update fused behind the activation
  // (efficient_gat.py:144 final_mlp[2], spatial_diffusion.py:485-627): see HeadFinalArgs; nothing else is stored then
  const struct HeadFinalArgs* head = nullptr;
  // optional (with img_node_slot): folded last layer (fold.cu) -- the output columns are [Q | K | V'] with Q / K
  // img_H * img_C wide and V' img_H * img_Cv wide (its own image, padded head dim img_Cvpad); no skip part
  int img_Cv = 0, img_Cvpad = 0;
};

// y = act(a @ w^T + bias) on CUDA cores, exact fp32 FMA.  a:[M,lda] w:[N,ldw] (both K-contiguous).
cudaError_t launch_linear_simt(const float* a, int lda, const float* w, int ldw, const float* bias,
                               const LinearOut& out, int M, int N, int K, int act, cudaStream_t s);

// fp32 -> split bf16 planes (hi = bf16(x), lo = bf16(x - hi)); rows x cols with row strides.
// row_map (optional, device int32 [rows]): output row r is read from input row row_map[r]
cudaError_t launch_split_bf16(const float* x, int ldx, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld,
                              int rows, int cols, cudaStream_t s, const int32_t* row_map = nullptr);
// y[r, :] = x[row_map[r], :]
cudaError_t launch_gather_rows(const float* x, int ldx, float* y, int ldy, int rows, int cols, const int32_t* row_map,
                               cudaStream_t s);

// CSR-by-target multigraph attention (TransformerConv message/aggregate stage).
struct AttnCsrArgs {
  const float* qkvs;     // [n, ld] rows laid out [Q | K | V | skip], each H*C wide
  int ld;                // row stride of qkvs (= 4*H*C)
  const int32_t* rowptr; // [n_targets + 1] in-edge ranges per target
  const int32_t* col;    // [E] source node per in-edge
  const float* weight;   // optional [E] multiplicity of each in-edge (null = 1)
  int n_targets;         // targets processed (first n_targets rows)
  int H, C;
  const float* resid;    // optional [n_targets, ld_resid] added after skip (trunk residual), or null
  int ld_resid;
  int act;               // activation applied to (attn + skip [+ resid])
  LinearOut out;         // [n_targets, H*C]
  float* scores;         // optional [E, H] raw scaled scores at CSR positions (for alpha output)
  float* stats;          // optional [n_targets, H, 2] (max, sum) for alpha output
  // continuation of an online softmax started by the dense-tile kernel (null = start fresh):
  const float* init_acc;      // [n, H*C] un-normalised accumulator
  const float* init_stats;    // [n, H, 2] (m, l)
  const int32_t* init_slot;   // [n] >= 0 where the init state is valid
  const int32_t* node_list;   // optional: the n_targets target ids to process (null = 0 .. n_targets-1)
  int n_coop = 0;             // launch_attn_csr_vrows: the first n_coop rows of node_list get a whole CTA per head
  // heavy-row kernel only: sources with img_slot[j] >= 0 are read from the split-bf16 K / V operand images
  // (hi + lo) instead of the fp32 row, so the GEMM can skip the fp32 K / V stores of dense-tile rows
  const int32_t* img_slot = nullptr;
  const __nv_bfloat16* kimg = nullptr;
  const __nv_bfloat16* vimg = nullptr;
  int img_Cpad = 0;
  // launch_attn_csr_vrows only: riders of the same launch -- the per-layer gather of the plan's gx_n promoted extra sources
  // (fp32 K / V rows of gx_src -> split-bf16 image rows gx_slot; same work as launch_gather_extra for 32-channel heads)
  int gx_n = 0; const int32_t* gx_src = nullptr; const int32_t* gx_slot = nullptr;
  __nv_bfloat16* gx_kimg = nullptr; __nv_bfloat16* gx_vimg = nullptr;
};
cudaError_t launch_attn_csr(const AttnCsrArgs& a, cudaStream_t s);
// warp-per-node variant for low-degree targets (all heads at once, fully coalesced rows)
bool attn_csr_rows_supported(int H, int C);
cudaError_t launch_attn_csr_rows(const AttnCsrArgs& a, cudaStream_t s);
// CTA-per-(node, head) variant for rows with hundreds of in-edges (needs node_list)
cudaError_t launch_attn_csr_heavy(const AttnCsrArgs& a, cudaStream_t s);
// lane-per-edge CTA-per-(node, head) variant for 32-channel heads and rows OUTSIDE the dense tiles (no init state)
bool attn_csr_vrows_supported(int H, int C);
cudaError_t launch_attn_csr_vrows(const AttnCsrArgs& a, cudaStream_t s);

// alpha[eid[p], h] = exp(scores[p,h] - max) / (sum + 1e-16)
cudaError_t launch_alpha_normalize(const float* scores, const float* stats, const int32_t* rowptr,
                                   const int32_t* eid, int n_targets, int H, float* alpha, cudaStream_t s);

struct PrologueArgs {
  const float* x;        // [M, C_in]
  const int64_t* t;      // [M] or null -> t_uniform
  int t_uniform;
  const float* P;        // [M, Hm] hoisted feats @ W1[:, :Dv]^T + b1   (or null -> b1 only)
  const float* b1;       // [Hm]
  const float* pos_w0;   // [16, C_in]
  const float* pos_b0;   // [16]
  const float* pos_w2;   // [32, 16]
  const float* pos_b2;   // [32]
  const float* time_emb; // [T, 32]
  const float* w1pt_T;   // [64, Hm]  transposed W1[:, Dv:Dv+64]
  // table form (optional, both or neither): wc_T [16, Hm] = (W1[:, pos] W_p2)^T, tt [T, Hm] = time_emb W1[:, time]^T + W1[:, pos] b_p2
  const float* wc_T = nullptr; const float* tt = nullptr;
  int M, C_in, Hm, T, act;
  LinearOut out;         // [M, Hm]
  // optional: the engine's internal node order (DensePlan::ext_of_int): internal row r reads x / t of the caller's row
  // row_ext[r]; P and the outputs are in internal order
  const int32_t* row_ext = nullptr;
};
cudaError_t launch_prologue(const PrologueArgs& a, cudaStream_t s);

// sampler update modes for the head epilogue
enum StepMode : int { STEP_NONE = 0, STEP_DDPM = 1, STEP_DDIM = 2 };

// per-node timesteps: schedule tables on the device + the batch-wide minimum of t (for "(prev_timestep >= 0).all()")
struct StepTables {
  da_schedule sched;
  const int64_t* t = nullptr;    // [M] per node, the caller's order (null = uniform t: coefficients come in da_step_coef)
  const int32_t* tmin = nullptr; // device scalar: min over the nodes of t
};
__device__ __forceinline__ da_step_coef node_coef(const da_step_coef& base, const StepTables& tb, int ext) {
  da_step_coef c = base;
  int t = (int)tb.t[ext];
  t = t < 0 ? 0 : (t >= tb.sched.steps ? tb.sched.steps - 1 : t);
  c.t = t;
  c.beta_t = tb.sched.betas[t];
  c.sqrt_one_minus_acp = tb.sched.sqrt_one_minus_alphas_cumprod[t];
  c.sqrt_recip_alpha = tb.sched.sqrt_recip_alphas[t];
  c.posterior_variance = tb.sched.posterior_variance[t];
  c.acp = tb.sched.alphas_cumprod[t];
  c.sqrt_recip_acp = tb.sched.sqrt_recip_alphas_cumprod[t];
  c.sqrt_recipm1_acp = tb.sched.sqrt_recipm1_alphas_cumprod[t];
  const bool all_prev = *tb.tmin >= tb.sched.inference_ratio;   // (prev_timestep >= 0).all()
  c.has_prev = all_prev ? 1 : 0;
  c.acp_prev = all_prev ? tb.sched.alphas_cumprod[t - tb.sched.inference_ratio] : 1.f;
  return c;
}

struct HeadFinalArgs {
  const float* u;        // [M, Nh] hidden of the head after GELU
  int Nh;
  const float* w_b;      // 2D: [C_out, 32];  SE3: mlp_t.2 [3,256]
  const float* b_b;
  const float* w_r;      // SE3: mlp_r.2 [3,256]
  const float* b_r;
  int M, C_out, head_kind;
  int step_mode;
  da_step_coef coef;
  const float* x_in;     // [M, C] current sample (step modes)
  const float* noise;    // [M, C] or null
  float* out;            // [M, C_out] model output (STEP_NONE) or x_prev
  const int32_t* row_ext = nullptr;   // optional: internal row r reads / writes the caller's row row_ext[r] of x_in, noise, out
  StepTables tabs;                    // tabs.t != null: per-node schedule coefficients (gathered on the device)
};
cudaError_t launch_min_t(const int64_t* t, int n, int32_t* out, cudaStream_t s);

// Folded 2-D head (fold.cu): u = GELU(sum_heads partial + h (W_a W_2)^T + x3 (W_a W_skip)^T + bias), then `fin`
struct HeadFoldArgs {
  const float* partial;            // [M, H, 32] per-head normalised aggregates of the 32-channel folded values
  int H;
  const __nv_bfloat16* h_hi; const __nv_bfloat16* h_lo; int ld_h, Hm;     // trunk hidden h (split-bf16), first Hm columns
  const __nv_bfloat16* x_hi; const __nv_bfloat16* x_lo; int ld_x, hid;    // input of the last graph layer
  const __nv_bfloat16* w_hi; const __nv_bfloat16* w_lo;   // [32, Hm + hid] split-bf16: columns (W_a W_2) then (W_a W_skip)
  const float* bias;               // [32]
  HeadFinalArgs fin;               // final_mlp[2], sampler update, outputs (u / Nh unused: Nh must be 32)
};
cudaError_t launch_head_fold(const HeadFoldArgs& a, cudaStream_t s);
// fp64-accumulated product of fp32 matrices with free strides (weight folding, once per da_load_weights):
// out[i*ldo_r + j*ldo_c] = sum_t A[i*lda + t] * B[t*ldb_r + j*ldb_c] + add_scale * add[i*add_si + j*add_sj]
cudaError_t launch_matmul_f64(const float* A, int lda, const float* B, int ldb_r, int ldb_c, const float* add, int add_si,
                              int add_sj, float add_scale, float* out, int ldo_r, int ldo_c, int m, int n, int k, cudaStream_t s);
// hi[r, col0 + ids[r]] = 1 (one-hot columns of the virtual rows, see fold.cu)
cudaError_t launch_onehot_rows(__nv_bfloat16* hi, int ld, int col0, const int32_t* ids, int rows, cudaStream_t s);
cudaError_t launch_head_final(const HeadFinalArgs& a, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// Sampler updates (one value per thread), evaluated op by op in the reference's order with
// explicit round-to-nearest intrinsics so no FMA contraction changes the result.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ddpm_update(float x, float out, float noise, const da_step_coef& c) {
  // spatial_diffusion.py:495-510
  float mean = __fmul_rn(c.sqrt_recip_alpha, __fsub_rn(x, __fdiv_rn(__fmul_rn(c.beta_t, out), c.sqrt_one_minus_acp)));
  if (c.t_index == 0) return mean;
  return __fadd_rn(mean, __fmul_rn(sqrtf(c.posterior_variance), noise));
}

__device__ __forceinline__ float ddim_x0(float x, float out, const da_step_coef& c) {
  // spatial_diffusion.py:603-606
  if (c.pred == DA_PRED_START_X) return out;
  float beta = __fsub_rn(1.f, c.acp);
  return __fdiv_rn(__fsub_rn(x, __fmul_rn(sqrtf(beta), out)), sqrtf(c.acp));
}

__device__ __forceinline__ float ddim_update(float x, float out, float noise, const da_step_coef& c) {
  // spatial_diffusion.py:548-627 with _get_variance :528-546 and _predict_eps_from_xstart :629-632
  float x0 = ddim_x0(x, out, c);
  float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c.sqrt_recip_acp, x), x0), c.sqrt_recipm1_acp);
  float beta = __fsub_rn(1.f, c.acp), beta_prev = __fsub_rn(1.f, c.acp_prev);
  float variance = __fmul_rn(__fdiv_rn(beta_prev, beta), __fsub_rn(1.f, __fdiv_rn(c.acp, c.acp_prev)));
  float std_eta = __fmul_rn(c.eta, sqrtf(variance));
  float dir = __fmul_rn(sqrtf(__fsub_rn(__fsub_rn(1.f, c.acp_prev), __fmul_rn(std_eta, std_eta))), eps);
  float prev = __fadd_rn(__fmul_rn(sqrtf(c.acp_prev), x0), dir);
  if (c.eta > 0.f) prev = __fadd_rn(prev, __fmul_rn(std_eta, noise));
  return prev;
}

__device__ __forceinline__ float step_update(int mode, float x, float out, float noise, const da_step_coef& c) {
  if (mode == STEP_DDPM) return ddpm_update(x, out, noise, c);
  if (mode == STEP_DDIM) return ddim_update(x, out, noise, c);
  return out;
}
cudaError_t launch_sampler_update(const float* x_in, const float* model_out, float* x_out, int M, int C,
                                  int head_kind, int step_mode, const da_step_coef& coef,
                                  const float* noise, cudaStream_t s);

// Stream-ordered scratch allocations for the once-per-batch graph preparation: served from the device's default
// memory pool with an unlimited release threshold, so after the first batch no cudaMalloc / cudaFree (both
// device-synchronising and ~1 ms for 100 MB blocks) happens on the da_set_graph path.
inline cudaError_t tmp_alloc(void** p, size_t bytes, cudaStream_t s) {
  static bool pool_ready = false;
  if (!pool_ready) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    pool_ready = true;
  }
  return cudaMallocAsync(p, bytes ? bytes : 16, s);
}
template <typename T>
inline cudaError_t tmp_alloc(T** p, size_t bytes, cudaStream_t s) { return tmp_alloc(reinterpret_cast<void**>(p), bytes, s); }
inline void tmp_free(void* p, cudaStream_t s) { if (p) cudaFreeAsync(p, s); }
// copy on `s` and wait for `s` only: unlike cudaMemcpy it does not serialise with the legacy default stream, so a
// batch can be prepared on a side stream while another handle computes (GNN_Diffusion.prefetch)
inline cudaError_t copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, s);
  return e != cudaSuccess ? e : cudaStreamSynchronize(s);
}

// graph structure
struct CsrGraph {
  int32_t* rowptr = nullptr;  // [n + 1]
  int32_t* col = nullptr;     // [E]
  int32_t* eid = nullptr;     // [E] original edge index of each CSR slot (uncompressed build only)
  float* weight = nullptr;    // [E] multiplicity of each (target, source) pair (compressed build only)
  int64_t E = 0;
  int n = 0;
};
// Builds CSR by target with a stable radix sort.  Allocates with cudaMallocAsync-free plain cudaMalloc.
cudaError_t build_csr(const int64_t* src, const int64_t* dst, int64_t E, int n, CsrGraph* g, cudaStream_t s,
                      const char** err);
cudaError_t build_csr_compressed(const int64_t* src, const int64_t* dst, int64_t E, int n, CsrGraph* g, cudaStream_t s,
                                 const char** err);
void free_csr(CsrGraph* g, cudaStream_t s = nullptr);   // stream-ordered (cudaFreeAsync on s)

// greedy assignment metric (scope row N2): out [sum n, 3] int64 = (row, column, int64(distance)) in greedy order
cudaError_t launch_greedy_assign(const float* pos1, int ld1, const float* pos2, int ld2, const int32_t* graph_ptr,
                                 int n_graphs, int max_n, int64_t* out, cudaStream_t s);

// Exphander edge list on the device from per-graph permutations (scope row N3)
cudaError_t launch_expander_edges(const int32_t* perm, int n, int degree, int n_graphs, int64_t* src, int64_t* dst,
                                  cudaStream_t s);

// training (scope row N1)
cudaError_t launch_attn_backward(const float* qkvs, const float* dO, const CsrGraph& by_target, const CsrGraph& by_source,
                                 const float* stats, int n, int H, int C, float* dqkvs, float* delta, cudaStream_t s);
// per-graph shared-memory backward for batches whose edges all sit in the dense plan's bitmaps (train.cu)
bool attn_backward_dense_fits(int n_max, int C, int bm_words_max);
cudaError_t launch_attn_backward_dense(const float* qkvs, const float* dO, const void* graphs_dev, int n_graphs, int n_max,
                                       int bm_words_max, const uint32_t* bitmap, const float* stats, int n, int H, int C,
                                       float* dqkvs, float* delta, cudaStream_t s);
cudaError_t launch_adafactor(const da_adafactor_param* params_dev, int n, float eps1, float eps2, float clip, float weight_decay,
                             cudaStream_t s);
cudaError_t launch_linear_wgrad(const float* dY, const float* X, float* dW, float* db, int M, int N, int K, cudaStream_t s);

// NHWC fp32 convolution pieces of the EfficientNet-B0 patch encoder (scope row N4, conv.cu); bias + activation fused
cudaError_t launch_conv2d_nhwc(const float* x, const float* w, const float* b, float* y, int N, int H, int W, int Cin, int Cout,
                               int k, int stride, int pad, int act, cudaStream_t s);      // w [Cout][k][k][Cin]
cudaError_t launch_dwconv2d_nhwc(const float* x, const float* w, const float* b, float* y, int N, int H, int W, int C, int k,
                                 int stride, int pad, int act, cudaStream_t s);           // w [k][k][C]
cudaError_t launch_spatial_mean(const float* x, float* y, int ldy, int N, int HW, int C, cudaStream_t s);   // y[n, c] = mean over HW
cudaError_t launch_channel_scale(float* x, const float* gate, int ldg, int N, int HW, int C, cudaStream_t s);   // x[n, :, c] *= gate[n, c]
cudaError_t launch_normalize_to_nhwc(const float* x, const float* mean, const float* stdv, float* y, int N, int C, int HW, cudaStream_t s);
cudaError_t launch_add_inplace(float* y, const float* x, size_t n, cudaStream_t s);   // y += x (n % 4 == 0)

cudaError_t launch_fill_rows(float* dst, int ld, const float* table, const int32_t* ids, int rows, int cols,
                             cudaStream_t s);
// out[g, c] = max over rows seg_ptr[g] .. seg_ptr[g + 1] of x[row, c] (PointNet's global max pool, pointnet.py:40)
cudaError_t launch_segment_max(const float* x, int ld, const int32_t* seg_ptr, int n_seg, int cols, float* out,
                               cudaStream_t s);

}  // namespace da

// ---------------------------------------------------------------------------------------------
// Dense-tile attention plan (attn_mode = DA_ATTN_AUTO).
//
// Graphs that are large and dense enough are processed as masked dense attention on the tensor
// cores: their nodes are grouped in 128-row target tiles / 64-row source blocks, their in-graph
// edges (multiplicity one) become an adjacency bitmap, and Q / K / V^T are repacked per layer into
// split-bf16 "operand images" (the exact shared-memory layout tcgen05.mma consumes, so a tile is
// one contiguous bulk copy).  Every other edge (virtual-node wiring, duplicates, cross-graph
// edges, small or very sparse graphs) stays in a residual CSR that continues the same online
// softmax, so the union is exactly the reference's segment softmax over ALL in-edges.
// ---------------------------------------------------------------------------------------------
namespace da {

struct TileInfo {
  int32_t node0;     // first global node id of this 128-row target tile
  int32_t rows;      // valid rows (<= 128)
  int32_t gblock0;   // global index of the graph's first 64-row source block
  int32_t gn;        // nodes in the graph
  int32_t bm_words;  // uint32 words per bitmap row of this graph (multiple of 2)
  int32_t row0;      // local index (within the graph) of the tile's first row
  int64_t bm_off;    // word offset of the graph's bitmap
  // the 64-source blocks this tile has to visit, in order (DensePlan::blk_list + list_off): blocks whose bitmap words
  // are all zero for every row of the tile are skipped.  Entry = (block index within the graph << 1) | full, where
  // full = every (valid row, column) bit of the block is set (no masking needed).
  int32_t n_list;
  int32_t list_off;
};

struct DensePlan {
  int n_tiles = 0;           // 128-row tiles (image rows = n_tiles * 128)
  int n_dense_graphs = 0;
  int64_t n_dense_edges = 0;
  TileInfo* tiles = nullptr;       // [n_tiles] device
  int32_t* node_slot = nullptr;    // [n_total] device: image row (tile * 128 + r) of each node, -1 if not dense
  uint32_t* bitmap = nullptr;      // device
  size_t bitmap_words = 0;
  CsrGraph residual;               // CSR by target over the edges not in the bitmap
  // targets split by residual in-degree: "light" rows (<= HEAVY_MIN_DEG = 96 edges, plan.cu) take the warp-per-node kernel,
  // "heavy" rows (virtual nodes with hundreds of in-edges) the edge-parallel one.  Real nodes first.
  int32_t* light = nullptr; int n_light = 0, n_light_real = 0;
  int32_t* heavy = nullptr; int n_heavy = 0, n_heavy_real = 0;
  // rows finalised inside the dense kernel (in a dense tile, <= DA_FUSE_MAX_RESIDUAL residual in-edges), and the
  // light rows that are NOT (outside every tile, or more residual edges): the CSR rows kernel only sees the latter
  uint8_t* row_fused = nullptr;    // [n_total] device
  int n_fused = 0;
  // "promoted" residual edges: a multiplicity-one residual in-edge of a dense-tile row whose source is not a
  // column of the row's graph (virtual-node wiring, cross-graph edges, the second copy of a duplicate) becomes a
  // bitmap bit on an EXTRA column of that graph: a copy of the source's K / V rows is placed in one of the padding
  // image rows behind the graph's own nodes (gather_extra kernel, once per layer).  The tensor-core kernel then
  // covers those edges for free and the rows need no CSR continuation at all.
  int n_extra = 0;                 // extra columns over all graphs
  int32_t* x_src = nullptr;        // [n_extra] device: node whose K / V rows are copied
  int32_t* x_slot = nullptr;       // [n_extra] device: destination image row (tile * 128 + r)
  int64_t n_promoted_edges = 0;
  int32_t* light_nf = nullptr; int n_light_nf = 0, n_light_nf_real = 0;
  // true when no row handled by the CSR kernels (heavy rows, un-fused light rows) lies inside a dense tile: those
  // kernels then neither read the dense kernel's (acc, stats) nor race with its output rows, and can run next to it
  bool csr_rows_independent = false;
  // every real node sits in a dense tile and has no residual in-edge left (the folded last layer needs this)
  bool real_rows_clean = false;
  int32_t* csr_rows = nullptr; int n_csr_rows = 0, n_csr_rows_real = 0;   // heavy rows then un-fused light rows, one list
  // per 128-row tile of the node index space: bit 0 = a row has residual in-edges (its fp32 Q is read),
  // bit 1 = a row is a residual source (fp32 K / V read); [0] all targets, [1] last layer (real targets only)
  uint8_t* f32_tile_flags[2] = {nullptr, nullptr};
  // Internal node order.  The engine is free to number the nodes of a graph as it likes as long as it reads its inputs
  // and writes its outputs in the caller's order.  For sparse-ish dense-tile graphs the planner walks the graph greedily
  // along maximal neighbourhood overlap, which recovers the ring order of the reference's Exphander graphs (a random
  // permutation joined to its d/2 cyclic shifts, puzzle_dataset.py:133-152): in that order the adjacency is a band, a
  // 128-row tile only touches ~(128 + d) / 64 of the graph's source blocks, and the rest is skipped.
  // ext_of_int[r] = caller's node id of internal row r (null = identity).  Every node id inside the plan (bitmap rows /
  // columns, residual CSR, slots, row lists) is INTERNAL.
  int32_t* ext_of_int = nullptr;
  int n_reordered_graphs = 0;
  uint16_t* blk_list = nullptr;    // [n_tiles * max_blocks] (see TileInfo)
  int max_blocks = 0;
  int64_t n_blocks_total = 0, n_blocks_listed = 0, n_blocks_full = 0;
};
void free_plan(DensePlan* p, cudaStream_t s = nullptr);   // stream-ordered (cudaFreeAsync on s)
// Classifies the edges, fills the bitmap and builds the residual CSR.  Synchronous.  allow_reorder: the caller reads the
// plan's ext_of_int and feeds / drains the kernels in internal order (the fused engine); false keeps the caller's order.
cudaError_t build_dense_plan(const int64_t* src, const int64_t* dst, int64_t E, const int64_t* batch, int num_real,
                             int num_total, DensePlan* plan, cudaStream_t s, const char** err, bool allow_reorder = false);

// fp32 [Q | K | V | skip] rows -> split-bf16 operand images of the dense tiles
struct PackArgs {
  const float* qkvs; int ld;       // [n, 4*H*C]
  const int32_t* node_slot; int n;
  int H, C, Cpad;
  __nv_bfloat16* qimg; __nv_bfloat16* kimg; __nv_bfloat16* vimg;
  int Cv = 0, Cvpad = 0;   // gather_extra only, folded last layer: rows are [Q | K | V'] with V' H*Cv wide (0 = same as C)
};
cudaError_t launch_pack_images(const PackArgs& a, cudaStream_t s);
// copies the fp32 K / V rows of the plan's extra sources into their image rows (PackArgs: node_slot unused)
cudaError_t launch_gather_extra(const PackArgs& a, const int32_t* x_src, const int32_t* x_slot, int n_extra, cudaStream_t s);

struct AttnDenseArgs {
  const __nv_bfloat16* qimg; const __nv_bfloat16* kimg; const __nv_bfloat16* vimg;
  const TileInfo* tiles; int n_tiles;
  const uint32_t* bitmap;
  int H, C, Cpad;
  float* acc;    // [n, H*C] un-normalised sum_e exp(a_e - m) v_e of the bitmap edges
  float* stats;  // [n, H, 2] (m, l) of the bitmap edges, natural-log units
  long long* dbg;  // optional timing trace of CTA 0 (clock64 stamps); null in production
  // Fused finalisation (optional, row_fused != null): rows with row_fused[node] != 0 have at most
  // DA_FUSE_MAX_RESIDUAL residual in-edges; the dense kernel's epilogue continues the online softmax over them
  // (fp32 Q / K / V rows of `qkvs`), normalises, adds skip (+ resid), applies `act` and writes the layer
  // output itself -- no (acc, stats) round trip through HBM and no CSR continuation launch for those rows.
  const uint8_t* row_fused = nullptr;
  const float* qkvs = nullptr; int ld = 0;      // [n_rows, 4*H*C] rows [Q | K | V | skip]
  int n_rows = 0, n_rows_resid = 0, n_rows_out = 0;   // allocated rows of qkvs / resid / out.hi (tensor-map bounds)
  const int32_t* rowptr = nullptr; const int32_t* col = nullptr; const float* weight = nullptr;   // residual CSR
  const float* resid = nullptr; int ld_resid = 0;   // optional second addend (trunk residual), staged like skip when it fits
  int act = 0;
  LinearOut out;
  const uint16_t* blk_list = nullptr;   // DensePlan::blk_list (required)
  int stagger_ns = 0;                   // attn_hidden.cu: start delay of every CTA's second stream
};
// Folded last layer (fold.cu): scores from the C-channel Q / K images, aggregation of the Cv = 32-channel folded
// values V'; every row of every tile is finalised here (the planner guarantees no residual in-edge on a real row):
// partial[node, head, :] = sum_j alpha_ij V'_j
struct AttnFoldArgs {
  const __nv_bfloat16* qimg; const __nv_bfloat16* kimg; const __nv_bfloat16* vimg;
  const TileInfo* tiles; int n_tiles;
  const uint32_t* bitmap; const uint16_t* blk_list;
  int H, C, Cpad;      // score head dim (scale = 1 / sqrt(C)) and its padding in the Q / K images
  float* partial;      // [n, H, 32]
  int persistent;      // 1: one CTA per SM walks the (tile, head) items (needs 2 * Cpad + 192 <= 512 TMEM columns)
};
bool attn_dense_fold_supported(int Cpad);
cudaError_t launch_attn_dense_fold(const AttnFoldArgs& a, cudaStream_t s);
constexpr int DA_FUSE_MAX_RESIDUAL = 2;
// true when the fused epilogue's staging area (128 skip rows of C floats) fits the kernel's K ring
bool attn_dense_can_fuse(int C);
cudaError_t launch_attn_dense(const AttnDenseArgs& a, cudaStream_t s);
// Persistent two-stream form for the 32-channel hidden layers (attn_hidden.cu): every valid tile row must be finalised by
// the kernel without residual in-edges (DensePlan::real_rows_clean); same results as launch_attn_dense
bool attn_hidden_persist_supported(const AttnDenseArgs& a);
cudaError_t launch_attn_hidden_persist(const AttnDenseArgs& a, cudaStream_t s);
size_t dense_image_elems(int n_tiles, int H, int Cpad);  // elements of ONE of the q / k / v image buffers

}  // namespace da
