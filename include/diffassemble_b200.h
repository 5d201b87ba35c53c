/*
 * diffassemble_b200.h -- C ABI of the B200-native DiffAssemble denoiser + sampler step.
 *
 * This is the drop-in boundary for ONE hot path of IIT-PAVIS/DiffAssemble: the
 * per-timestep graph-transformer denoiser and the DDPM/DDIM update that drives it.
 * The reference has no FFI of its own (pure Python); each entry
 * point below cites the reference Python interface it replaces (paths relative to
 * the reference checkout).  The Python mirror of those interfaces lives in
 * diffassemble_b200/ and binds this header with ctypes.
 *
 * Conventions
 *   - plain C types only; every function returns 0 (DA_OK) or a negative da_status;
 *     nothing throws across the ABI.  da_last_error() gives a human-readable reason.
 *   - "device pointer" = CUDA device memory of the handle's device, contiguous,
 *     16-byte aligned, owned by the caller.  The library owns packed weights,
 *     graph structure (CSR by target) and all workspace.
 *   - every compute call is asynchronous on the cudaStream_t passed as `stream`
 *     (as void*; NULL = legacy default stream).  One handle per (device, stream);
 *     handles are not thread-safe.
 *   - random numbers stay with the caller (noise is passed in) so seeds match the
 *     reference call for call.
 *   - there is NO CPU fallback: every entry point fails with DA_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef DIFFASSEMBLE_B200_H
#define DIFFASSEMBLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DA_ABI_VERSION 1

typedef enum da_status {
  DA_OK = 0,
  DA_ERR_INVALID = -1,     /* bad argument / shape / state (e.g. forward before set_graph) */
  DA_ERR_CUDA = -2,        /* CUDA runtime / driver error, or no usable sm_100 device       */
  DA_ERR_UNSUPPORTED = -3, /* configuration outside what the kernels implement              */
  DA_ERR_MISSING = -4      /* a required weight was not loaded                              */
} da_status;

/* head_kind */
#define DA_HEAD_2D 0  /* Eff_GAT.final_mlp:  D -> 32 -> C_out          (efficient_gat.py:88-92,144)   */
#define DA_HEAD_SE3 1 /* Eff_GAT_3d.mlp_t / mlp_r -> [quat(4), t(3)]    (efficient_gat_3d.py:142-151,211-220) */
/* arch */
#define DA_ARCH_TRANSFORMER 0 /* Transformer_GNN: GELU after every layer but the last (Transformer_GNN.py:29-46) */
#define DA_ARCH_EXOPHORMER 1  /* Exophormer_GNN: no activation, virtual nodes       (exophormer_gnn.py:161-215) */
/* gemm_mode */
#define DA_GEMM_FP32_SIMT 0  /* exact fp32 FMA on CUDA cores (debug / anchor mode)                         */
#define DA_GEMM_BF16X3_UMMA 1 /* tcgen05 tensor cores, 3-pass split-bf16 (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi), fp32 accumulate in TMEM */
/* attn_mode */
#define DA_ATTN_CSR 0  /* every edge through the CSR-by-target warp kernel                    */
#define DA_ATTN_AUTO 1 /* per-graph dense bitmap tiles where the graph is dense enough, CSR for the rest */
/* model_mean_type (spatial_diffusion.py:60-67) */
#define DA_PRED_START_X 0
#define DA_PRED_EPSILON 1

typedef struct da_config {
  int32_t abi_version;  /* must be DA_ABI_VERSION */
  int32_t device;       /* CUDA device ordinal */
  int32_t feat_dim;     /* Dv: per-node encoder feature width (1088 for efficientnet_b0; 128 for pointnet) */
  int32_t in_channels;  /* C_in : pose width fed to pos_mlp (2, 4 with rotation, 7 in 3D) */
  int32_t out_channels; /* C_out: 2 / 4 (2D); 7 for the SE(3) head */
  int32_t heads;        /* H = 8 */
  int32_t hidden;       /* H*C of the inner layers = 256 */
  int32_t n_layers;     /* 4 */
  int32_t steps;        /* T, rows of time_emb */
  int32_t mlp_hidden;   /* 128 (Eff_GAT.mlp) / 256 (Eff_GAT_3d.mlp) */
  int32_t head_kind;    /* DA_HEAD_* */
  int32_t arch;         /* DA_ARCH_* */
  int32_t virt_nodes;   /* V (exophormer only; 0 = none) */
  int32_t gemm_mode;    /* DA_GEMM_* */
  int32_t attn_mode;    /* DA_ATTN_* */
  int32_t reserved[8];  /* zero */
} da_config;

typedef struct da_handle da_handle;

/* One named parameter tensor, fp32, row-major, in host OR device memory (UVA copy).
 * Names are the reference state_dict keys of the denoiser (prefix "model." stripped),
 * e.g. "time_emb.weight", "pos_mlp.0.weight", "mlp.2.bias",
 * "gnn_backbone.module_list.3.lin_key.weight", "gnn_backbone.virt_node_embedding.weight",
 * "final_mlp.0.weight", "mlp_t.2.bias" (SURVEY.md section 2.3f). */
typedef struct da_weight_desc {
  const char* name;
  const float* data;
  int64_t rows; /* weight: out_features (embedding: num_embeddings); bias: length */
  int64_t cols; /* weight: in_features  (embedding: dim);            bias: 1      */
} da_weight_desc;

/* Replaces Eff_GAT.__init__ / Eff_GAT_3d.__init__ (efficient_gat.py:23-112,
 * efficient_gat_3d.py:57-151): allocates the handle, no weights yet. */
int da_create(da_handle** out, const da_config* cfg);
void da_destroy(da_handle* h);
const char* da_last_error(const da_handle* h); /* h may be NULL: last da_create error */

/* Replaces nn.Module.load_state_dict for the denoiser (checkpoint -> packed device
 * weights).  May be called repeatedly; every call re-packs (QKV+skip concatenated per
 * layer, split-bf16 planes for the tensor-core path). */
int da_load_weights(da_handle* h, const da_weight_desc* w, int32_t n);

/* Replaces the (edge_index, batch) arguments of Eff_GAT.forward_with_feats
 * (efficient_gat.py:121-129) and, for arch=EXOPHORMER, consumes the extended
 * multigraph that exophormer_gnn.py:164-200 builds (the host mirror builds it once
 * per batch instead of once per step).
 *   edge_src/edge_dst : int64 device pointers, E entries: message j=src[e] -> i=dst[e];
 *                       a multiset (duplicates count twice in the softmax)
 *   batch             : int64 device pointer, num_real entries, non-decreasing graph ids
 *   num_real          : real nodes (rows of x / out)
 *   num_total         : num_real + virtual rows appended after the real ones
 *   virt_ids          : int32 device pointer, (num_total-num_real) rows of
 *                       virt_node_embedding to place in the virtual rows (NULL if none)
 * Builds CSR-by-target on the device (stable sort: deterministic summation order). */
int da_set_graph(da_handle* h, const int64_t* edge_src, const int64_t* edge_dst, int64_t E,
                 const int64_t* batch, int32_t num_real, int32_t num_total,
                 const int32_t* virt_ids, void* stream);

/* Replaces the patch_feats / pcd_feats argument (efficient_gat.py:127,
 * efficient_gat_3d.py:178): feats is fp32 [num_real, feat_dim] on the device, or NULL
 * for all-zero features (classifier-free "unconditional" pass,
 * spatial_diffusion.py:578-586).  Runs the step-invariant part of mlp[0] once
 * (feats @ W1[:, :Dv]^T + b1) so the per-step cost is the 64 pose+time columns only. */
int da_set_features(da_handle* h, const float* feats, void* stream);

/* Replaces Eff_GAT.forward_with_feats / Eff_GAT_3d.forward_with_feats
 * (efficient_gat.py:121-146, efficient_gat_3d.py:173-220).
 *   x   : fp32 [num_real, C_in]     t : int64 [num_real] (per node, as the reference)
 *   out : fp32 [num_real, C_out]
 *   alpha_last : NULL, or fp32 [E, H] attention weights of the LAST layer in the
 *                caller's edge order (exophormer_gnn.py:205-207); DA_ATTN_CSR only. */
int da_forward(da_handle* h, const float* x, const int64_t* t, float* out, float* alpha_last,
               void* stream);

/* Same forward, returning the attention weights of EVERY layer as the reference's Transformer_GNN / Exophormer_GNN do
 * (one (edge_index, alpha) tuple per layer, Transformer_GNN.py:29-46): alpha_all is fp32 [n_layers, E, H] in the caller's
 * edge order.  DA_ATTN_CSR only (the tensor-core tiles never materialise per-edge weights). */
int da_forward_attn(da_handle* h, const float* x, const int64_t* t, float* out, float* alpha_all, void* stream);

/* Scalar schedule coefficients of one sampler step, computed by the caller in fp32
 * exactly as the reference does from its registered buffers
 * (spatial_diffusion.py:289-321). */
typedef struct da_step_coef {
  int32_t t;                 /* timestep index (uniform over nodes in sampling, spatial_diffusion.py:665) */
  int32_t t_index;           /* == t in the reference loop; DDPM adds no noise when 0 */
  int32_t pred;              /* DA_PRED_* */
  int32_t has_prev;          /* DDIM: (t - inference_ratio) >= 0 */
  float beta_t;              /* betas[t] */
  float sqrt_one_minus_acp;  /* sqrt_one_minus_alphas_cumprod[t] */
  float sqrt_recip_alpha;    /* sqrt_recip_alphas[t] */
  float posterior_variance;  /* posterior_variance[t] */
  float acp;                 /* alphas_cumprod[t] */
  float acp_prev;            /* alphas_cumprod[t - ratio], or 1 when !has_prev */
  float sqrt_recip_acp;      /* sqrt_recip_alphas_cumprod[t] */
  float sqrt_recipm1_acp;    /* sqrt_recipm1_alphas_cumprod[t] */
  float eta;                 /* 0 (DDIM) or 1 */
  float cfg_w;               /* classifier_free_w (unused unless x_uncond given) */
} da_step_coef;

/* Replaces GNN_Diffusion.p_sample_ddpm (spatial_diffusion.py:485-510): one denoiser
 * forward fused with the posterior-mean update.  x_in / x_out fp32 [num_real, C]
 * (may alias); noise fp32 [num_real, C] or NULL when t_index == 0. */
int da_ddpm_step(da_handle* h, const float* x_in, float* x_out, const da_step_coef* c,
                 const float* noise, void* stream);

/* Replaces GNN_Diffusion.p_sample_ddim (spatial_diffusion.py:548-627) for
 * classifier_free_prob == 0: denoiser forward fused with the DDIM update; noise is
 * only read when eta > 0.  For head_kind == DA_HEAD_SE3 this is the R^3 + SO(3)
 * update of spatial_diffusion_3d_test_double_diffusion.py:595-685. */
int da_ddim_step(da_handle* h, const float* x_in, float* x_out, const da_step_coef* c,
                 const float* noise, void* stream);

/* Per-node timesteps (the reference gathers every schedule coefficient per node, `extract`, spatial_diffusion.py:173-176;
 * p_sample is called with a [nodes] tensor t).  The registered schedule buffers of the module are passed as device pointers
 * and gathered on the device, so a caller of p_sample never has to inspect t on the host. */
typedef struct da_schedule {
  const float* betas;                          /* all fp32 device arrays of length `steps` */
  const float* alphas_cumprod;
  const float* sqrt_one_minus_alphas_cumprod;
  const float* sqrt_recip_alphas;
  const float* posterior_variance;
  const float* sqrt_recip_alphas_cumprod;
  const float* sqrt_recipm1_alphas_cumprod;
  int32_t steps;
  int32_t inference_ratio;                     /* DDIM: prev_timestep = t - inference_ratio (spatial_diffusion.py:557) */
} da_schedule;
/* da_ddpm_step / da_ddim_step with t int64 [num_real] on the device.  t_index is the reference's Python int (DDPM adds no
 * noise when it is 0); DDIM's "(prev_timestep >= 0).all()" is evaluated on the device over all nodes, as the reference
 * evaluates it over the whole batch (spatial_diffusion.py:535,560). */
int da_ddpm_step_t(da_handle* h, const float* x_in, float* x_out, const int64_t* t, int32_t t_index,
                   const da_schedule* sched, const float* noise, void* stream);
int da_ddim_step_t(da_handle* h, const float* x_in, float* x_out, const int64_t* t, int32_t pred, float eta,
                   const da_schedule* sched, const float* noise, void* stream);

/* Sampler update alone on a given model output (used for classifier-free guidance,
 * where two forwards are blended first, spatial_diffusion.py:568-589). */
int da_ddim_update(da_handle* h, const float* x_in, const float* model_out, float* x_out,
                   const da_step_coef* c, const float* noise, void* stream);

/* Introspection for tests, roofline accounting and gpu_launches. */
size_t da_workspace_bytes(const da_handle* h);
int64_t da_launch_count(const da_handle* h); /* kernels launched by this handle so far */
int da_graph_stats(const da_handle* h, int64_t* n_dense_edges, int64_t* n_csr_edges,
                   int32_t* n_dense_graphs);
/* What the dense-tile planner (attn_mode = DA_ATTN_AUTO) made of the bound graph; out receives 8 values:
 * [0] 128-row tiles, [1] (tile, 64-source block) pairs, [2] of which are visited (hold at least one edge), [3] of which have
 * every bit set, [4] graphs whose nodes were renumbered internally by the ring walk (inputs / outputs stay in the caller's
 * order), [5] promoted extra sources, [6] rows finalised inside the tensor-core kernel, [7] rows served by the CSR kernels; with n >= 10 also
 * [8] 1 when every real row is finalised by the tensor-core kernel without residual in-edges, [9] 1 when steps on this batch
 * take the weight-folded path (csrc/fold.cu: first projection from the 128-wide trunk hidden, last layer aggregated on the
 * 32-channel values of final_mlp[0] folded into lin_value; efficient_gat.py:135-145); with n >= 11 also [10] the number of
 * hidden-layer attention launches that took the persistent two-stream kernel (csrc/attn_hidden.cu) on this handle so far. */
int da_graph_plan_info(const da_handle* h, int64_t* out, int32_t n);
/* Built-in CUDA-event profiler: with da_set_profiling(h, 1) every kernel launch is bracketed
 * by events on its own stream and accumulated per launch site ("tag").  da_get_profile fills
 * up to n_classes entries (ms and launch counts per tag), optionally resets, and returns the
 * number of tags; da_profile_tag_name(i) names tag i ("qkvs_gemm_first", "attn_last", ...). */
int da_set_profiling(da_handle* h, int32_t enable);
int da_get_profile(da_handle* h, double* ms_out, int64_t* launches_out, int32_t n_classes, int32_t reset);
const char* da_profile_tag_name(int32_t i);

/* Stand-alone operator entry points (unit-level parity tests and micro-benchmarks).
 * y[M,N] = act(a[M,K] @ w[N,K]^T + bias[N]); act: 0 none, 1 GELU(erf), 2 LeakyReLU(0.2), 3 ReLU, 4 SiLU, 5 sigmoid.
 * mode = DA_GEMM_*; all pointers device fp32. */
int da_op_linear(int32_t mode, const float* a, const float* w, const float* bias, float* y,
                 int32_t M, int32_t N, int32_t K, int32_t act, void* stream);
/* Fused Adafactor step (scope row N1).  Replaces transformers.optimization.Adafactor.step -- the optimizer the reference
 * instantiates with default arguments (spatial_diffusion.py:701-705) -- for fp32 parameters of rank <= 2:
 * factored second moments (sq_row [rows], sq_col [cols]) for matrices, a full one (sq [rows*cols]) for vectors
 * (then sq_row = sq_col = NULL), relative step, parameter scaling, update clipping, no first moment.
 * `params` is an array of n descriptors IN DEVICE MEMORY; beta2t = 1 - step^decay_rate and
 * rel_step = min(1e-2, 1/sqrt(step)) are computed by the caller per tensor; rms_out (optional) receives RMS(p). */
typedef struct da_adafactor_param {
  float* p;
  const float* g;
  float* sq_row;
  float* sq_col;
  float* sq;
  float* rms_out;
  int32_t rows, cols;
  float beta2t, rel_step;
} da_adafactor_param;
int da_adafactor_step(const da_adafactor_param* params, int32_t n, float eps1, float eps2, float clip_threshold,
                      float weight_decay, void* stream);
/* Segment-wise column maximum: out[g, c] = max_{seg_ptr[g] <= r < seg_ptr[g+1]} x[r, c]  (x fp32 [rows, ld], out fp32
 * [n_seg, cols]).  The global max pool of the PointNet fragment encoder (puzzle_diff/model/backbones/pointnet.py:39-40),
 * scope row N4; the point-wise layers of that encoder are da_op_linear calls with eval-mode BatchNorm folded in. */
int da_op_segment_max(const float* x, int32_t ld, const int32_t* seg_ptr, int32_t n_seg, int32_t cols, float* out,
                      void* stream);
/* Scope row N4, 2-D side -- the pieces of the EfficientNet-B0 patch encoder that are not plain GEMMs (Eff_GAT.visual_features,
 * efficient_gat.py:40-42,149-189; timm's efficientnet_b0 feature pyramid).  All tensors NHWC fp32 on the device, eval-mode
 * BatchNorm folded into w / bias by the caller, act as in da_op_linear plus 4 = SiLU, 5 = sigmoid.  The point-wise (1x1)
 * convolutions and the squeeze-excite FCs are da_op_linear calls over [N*H*W, C] rows.
 *   da_op_conv2d_nhwc   : dense k x k convolution, w [Cout][k][k][Cin] (the 3 -> 32 stem)
 *   da_op_dwconv2d_nhwc : depth-wise k x k convolution, w [k][k][C], C % 4 == 0
 *   da_op_spatial_mean  : y[n, c] = mean over the HW pixels (squeeze), y row stride ldy
 *   da_op_channel_scale : x[n, p, c] *= gate[n, c] in place (excite), gate row stride ldg */
int da_op_conv2d_nhwc(const float* x, const float* w, const float* bias, float* y, int32_t N, int32_t H, int32_t W, int32_t Cin,
                      int32_t Cout, int32_t k, int32_t stride, int32_t pad, int32_t act, void* stream);
int da_op_dwconv2d_nhwc(const float* x, const float* w, const float* bias, float* y, int32_t N, int32_t H, int32_t W, int32_t C,
                        int32_t k, int32_t stride, int32_t pad, int32_t act, void* stream);
int da_op_spatial_mean(const float* x, float* y, int32_t ldy, int32_t N, int32_t HW, int32_t C, void* stream);
int da_op_channel_scale(float* x, const float* gate, int32_t ldg, int32_t N, int32_t HW, int32_t C, void* stream);
/* y[n, h, w, c] = (x[n, c, h, w] - mean[c]) / std[c]: the reference's input normalisation (efficient_gat.py:150) fused with the
 * NCHW -> NHWC layout change;  da_op_add_inplace: y += x over n floats (the residual of the repeated MBConv blocks). */
int da_op_normalize_to_nhwc(const float* x, const float* mean, const float* stdv, float* y, int32_t N, int32_t C, int32_t HW, void* stream);
int da_op_add_inplace(float* y, const float* x, int64_t n, void* stream);
/* Same operator with caller-provided scratch (split-bf16 operand planes): no allocation, no synchronisation. */
size_t da_op_linear_workspace_bytes(int32_t mode, int32_t M, int32_t N, int32_t K);
int da_op_linear_ws(int32_t mode, const float* a, const float* w, const float* bias, float* y, int32_t M, int32_t N,
                    int32_t K, int32_t act, void* workspace, size_t workspace_bytes, void* stream);
/* TransformerConv attention stage on precomputed projections (section 2.3c of SURVEY.md):
 * qkvs fp32 [n, 4*H*C] laid out [Q | K | V | skip]; edges as in da_set_graph;
 * y[n, H*C] = softmax-aggregate + skip. */
int da_op_graph_attention(const float* qkvs, const int64_t* edge_src, const int64_t* edge_dst,
                          int64_t E, int32_t n, int32_t H, int32_t C, float* y, float* alpha,
                          void* stream);

/* Same stage through the tensor-core dense-tile path (DA_ATTN_AUTO): `batch` (int64 [n], PyG
 * collation order) selects the per-graph bitmap tiles; edges the bitmap cannot hold (duplicates,
 * cross-graph, small or sparse graphs) run through the residual CSR.  *n_dense_edges (optional)
 * receives how many edges took the tensor-core path. */
int da_op_graph_attention_dense(const float* qkvs, const int64_t* edge_src, const int64_t* edge_dst,
                                int64_t E, const int64_t* batch, int32_t n, int32_t H, int32_t C,
                                float* y, int64_t* n_dense_edges, void* stream);

/* Scope row N2 -- replaces greedy_cost_assignment (spatial_diffusion.py:179-216), the metric step right
 * after the sampling loop: for each puzzle g (nodes graph_ptr[g] .. graph_ptr[g+1]) repeatedly assign the
 * globally closest unassigned (piece, grid cell) pair.  pos1 / pos2: fp32 device rows of at least 2 columns
 * (x, y) with row strides ld1 / ld2 (so `img[:, :2]` of an [N, 4] sample can be passed in place);
 * graph_ptr: int32 device [n_graphs + 1]; out: int64 device [N, 3] = (piece, cell, int64(distance)) per
 * puzzle in greedy order, indices local to the puzzle.  One CTA per puzzle, no host synchronisation. */
int da_greedy_cost_assignment(const float* pos1, int32_t ld1, const float* pos2, int32_t ld2,
                              const int32_t* graph_ptr, int32_t n_graphs, int32_t max_nodes, int64_t* out,
                              void* stream);

/* Scope row N3 -- replaces generate_random_regular_graph + PyG collation (puzzle_dataset.py:115-152) for a
 * batch of equally sized graphs: perm is int32 device [n_graphs, n] (one random permutation per graph, drawn by
 * the caller with the reference's numpy Generator); writes edge_src / edge_dst (int64 device,
 * n_graphs * n * degree entries each) in the reference's edge order with node ids offset by g * n. */
int da_expander_edge_index(const int32_t* perm, int32_t n, int32_t degree, int32_t n_graphs, int64_t* edge_src,
                           int64_t* edge_dst, void* stream);

/* ---- Scope row N1 (training step, spatial_diffusion.py:432-483,707-722): operator-level entry points the
 * host-side autograd functions call.  A da_graph holds the CSR-by-target and CSR-by-source forms of one batch's
 * edge multiset (built once per batch). */
typedef struct da_graph da_graph;
int da_graph_create(da_graph** out, const int64_t* edge_src, const int64_t* edge_dst, int64_t E, int32_t n, void* stream);
/* optional: also build the dense-tile plan of the batch (`batch` int64 [n], PyG collation order; same edge arrays as
 * da_graph_create).  When every edge lands in a bitmap (the dense puzzle graphs of the training configs) the forward
 * below runs on the tensor-core attention kernel; otherwise nothing changes. */
int da_graph_set_batch(da_graph* g, const int64_t* edge_src, const int64_t* edge_dst, const int64_t* batch, void* stream);
void da_graph_destroy(da_graph* g);
/* forward of the TransformerConv attention stage, also returning the per-(node, head) softmax statistics
 * stats[n, H, 2] = (max, sum) that the backward needs;  y[n, H*C] = aggregate + skip */
int da_op_graph_attention_fwd(const da_graph* g, const float* qkvs, int32_t H, int32_t C, float* y, float* stats,
                              void* stream);
/* backward: dy[n, H*C] -> dqkvs[n, 4*H*C] = [dQ | dK | dV | dskip];  delta_ws: fp32 workspace [n, H] */
int da_op_graph_attention_bwd(const da_graph* g, const float* qkvs, const float* stats, const float* dy, int32_t H,
                              int32_t C, float* dqkvs, float* delta_ws, void* stream);
/* weight / bias gradient of y = x @ w^T + b:  dw[N, K] = dy^T @ x,  db[N] = column sums of dy (db may be NULL) */
int da_op_linear_wgrad(const float* dy, const float* x, float* dw, float* db, int32_t M, int32_t N, int32_t K,
                       void* stream);

int da_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DIFFASSEMBLE_B200_H */
