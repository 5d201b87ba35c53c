// C-ABI entry points and per-step orchestration (include/diffassemble_b200.h).
#include <map>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>

#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

using namespace da;

namespace da {
bool pdl_enabled() {
  const char* e = getenv("DA_NO_PDL");   // read per launch (nanoseconds), so that a test can flip it inside one process
  return !(e != nullptr && e[0] == '1');
}
}  // namespace da

namespace {

thread_local std::string g_create_error;

// launch-site tags for the built-in CUDA-event profiler
enum Tag : int {
  TAG_HOIST_GEMM = 0, TAG_PROLOGUE, TAG_MLP2_GEMM, TAG_QKVS_GEMM_FIRST, TAG_QKVS_GEMM_MID, TAG_QKVS_GEMM_LAST,
  TAG_ATTN_HIDDEN, TAG_ATTN_LAST, TAG_HEAD_GEMM, TAG_HEAD_FINAL, TAG_OTHER, TAG_PACK_HIDDEN, TAG_PACK_LAST,
  TAG_ATTN_DENSE_HIDDEN, TAG_ATTN_DENSE_LAST, TAG_COUNT
};
const char* kTagNames[TAG_COUNT] = {"hoist_gemm", "prologue", "mlp2_gemm", "qkvs_gemm_first", "qkvs_gemm_mid",
                                    "qkvs_gemm_last", "attn_hidden", "attn_last", "head_gemm", "head_final", "other",
                                    "pack_hidden", "pack_last", "attn_dense_hidden", "attn_dense_last"};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, need ? need : 16);
    if (e == cudaSuccess) bytes = need ? need : 16;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Linear {  // y = x @ w^T + b ; w [N, K] row-major
  DevBuf w, b, w_hi, w_lo;
  int N = 0, K = 0;
  void release() { w.release(); b.release(); w_hi.release(); w_lo.release(); }
};

struct ProfEvent { cudaEvent_t a, b; int tag; };

}  // namespace

struct da_handle {
  da_config cfg{};
  int D = 0, L = 0, Nh = 0;
  std::string err;
  bool weights_loaded = false, graph_set = false, feats_set = false, feats_zero = false;
  // packed weights
  Linear hoist, mlp2, head1;
  std::vector<Linear> layer;
  DevBuf pos_w0, pos_b0, pos_w2, pos_b2, time_emb, w1pt_T, b1, headb_w, headb_b, headr_w, headr_b, virt_emb;
  DevBuf pro_wc, pro_tt;   // table form of the prologue (pointwise.cu): composed pos_mlp[2] / time_emb products
  bool no_pro_table = getenv("DA_NO_PRO_TABLE") != nullptr && getenv("DA_NO_PRO_TABLE")[0] == '1';
  // graph
  CsrGraph csr;        // every edge (attn_mode = CSR)
  DensePlan plan;      // bitmap tiles + residual CSR (attn_mode = AUTO)
  bool use_plan = false;
  int num_real = 0, num_total = 0;
  int dbg_layer = -1;
  // development switches (environment, read once per handle): DA_NO_FUSE=1 keeps the un-fused
  // dense -> (acc, stats) -> CSR-continuation pipeline for A/B measurements
  bool no_fuse = getenv("DA_NO_FUSE") != nullptr && getenv("DA_NO_FUSE")[0] == '1';
  bool no_side = getenv("DA_NO_SIDE") != nullptr && getenv("DA_NO_SIDE")[0] == '1';
  bool no_vrows = getenv("DA_NO_VROWS") != nullptr && getenv("DA_NO_VROWS")[0] == '1';
  bool no_head_fuse = getenv("DA_NO_HEAD_FUSE") != nullptr && getenv("DA_NO_HEAD_FUSE")[0] == '1';
  // Weight folding of the 2-D denoiser (fold.cu).  fold_cfg: the configuration allows it (decided in da_create; the trunk
  // hidden h is then kept with row stride Kf = mlp_hidden + 64 one-hot columns for the virtual rows); fold_ready: the
  // folded weights exist (da_load_weights).  A step takes the folded path when additionally every real row of the bound
  // batch is finalised by the dense-tile kernel (DensePlan::real_rows_clean) -- otherwise the unfolded pipeline runs.
  bool no_fold = getenv("DA_NO_FOLD") != nullptr && getenv("DA_NO_FOLD")[0] == '1';
  bool fold_persist = !(getenv("DA_FOLD_PERSIST") != nullptr && getenv("DA_FOLD_PERSIST")[0] == '0');
  // hidden layers on the persistent two-stream kernel (attn_hidden.cu) whenever its preconditions hold; "0": one CTA per (tile, head)
  bool hidden_persist = !(getenv("DA_HIDDEN_PERSIST") != nullptr && getenv("DA_HIDDEN_PERSIST")[0] == '0');
  bool no_gather_ride = getenv("DA_NO_GATHER_RIDE") != nullptr && getenv("DA_NO_GATHER_RIDE")[0] == '1';
  int hidden_stagger_ns = getenv("DA_HIDDEN_STAGGER_NS") != nullptr ? atoi(getenv("DA_HIDDEN_STAGGER_NS")) : 8000;
  bool fold_cfg = false, fold_ready = false;
  int Kf = 0;
  Linear fold0, fold3;
  DevBuf fold_wt, fold_wt_hi, fold_wt_lo, fold_b, fold_g, vimg_f, partial;
  // side stream: the CSR kernels of rows outside every dense tile (virtual nodes) run next to the dense kernel
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  void* dbg_trace = nullptr;   // development aid: clock64 trace buffer for the dense attention kernel
  DevBuf qimg, kimg, vimg, qimg_l, kimg_l, vimg_l, dacc, dstats;   // operand images: hidden layers / last layer
  // activations / workspace
  DevBuf P, hbuf, combined, qkvs, xa, xb, r, u, model_out, scores, stats;
  // split-bf16 operand planes for the tensor-core path
  DevBuf tmin;         // per-node-t sampler steps: min over the nodes of t (device scalar)
  DevBuf feats_perm;   // exact-fp32 mode: features gathered into the planner's internal node order
  DevBuf feats_sp_hi, feats_sp_lo, h_hi, h_lo, comb_hi, comb_lo, xa_hi, xa_lo, xb_hi, xb_lo, r_hi, r_lo;
  int64_t launches = 0;
  int64_t persist_launches = 0;   // hidden layers that ran on attn_hidden.cu's persistent kernel (da_graph_plan_info [10])
  bool profiling = false;
  std::vector<ProfEvent> prof_events;
  double prof_ms[TAG_COUNT] = {0};
  int64_t prof_n[TAG_COUNT] = {0};

  int fail(da_status st, const std::string& m) { err = m; return (int)st; }
  int cuda_fail(cudaError_t e, const char* where) {
    err = std::string(where) + ": " + cudaGetErrorString(e);
    return (int)DA_ERR_CUDA;
  }
  size_t workspace_bytes() const {
    const DevBuf* all[] = {&P, &hbuf, &combined, &qkvs, &xa, &xb, &r, &u, &model_out, &scores, &stats, &qimg, &kimg, &vimg, &feats_perm,
                           &qimg_l, &kimg_l, &vimg_l, &dacc, &dstats,
                           &feats_sp_hi, &feats_sp_lo, &h_hi, &h_lo, &comb_hi, &comb_lo, &xa_hi, &xa_lo,
                           &xb_hi, &xb_lo, &r_hi, &r_lo};
    size_t s = 0;
    for (auto* b : all) s += b->bytes;
    return s;
  }
};

namespace {

struct Scoped {  // records profiler events around one launch
  da_handle* h; cudaStream_t s; int tag; ProfEvent ev{};
  Scoped(da_handle* h_, cudaStream_t s_, int tag_) : h(h_), s(s_), tag(tag_) {
    h->launches++;
    if (h->profiling) {
      cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); ev.tag = tag;
      cudaEventRecord(ev.a, s);
    }
  }
  ~Scoped() {
    if (h->profiling) { cudaEventRecord(ev.b, s); h->prof_events.push_back(ev); }
  }
};

void drain_profile(da_handle* h) {
  for (auto& e : h->prof_events) {
    cudaEventSynchronize(e.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) { h->prof_ms[e.tag] += ms; h->prof_n[e.tag]++; }
    cudaEventDestroy(e.a); cudaEventDestroy(e.b);
  }
  h->prof_events.clear();
}

int layer_in(const da_handle* h, int l) { return l == 0 ? h->D : h->cfg.hidden; }
int layer_hc(const da_handle* h, int l) { return l == h->L - 1 ? h->D : h->cfg.hidden; }

bool use_umma(const da_handle* h) { return h->cfg.gemm_mode == DA_GEMM_BF16X3_UMMA; }

// One linear layer through whichever GEMM path the handle is configured for.
// a_f32 is used by the SIMT path, (a_hi, a_lo) by the tensor-core path.
cudaError_t run_linear(da_handle* h, const Linear& lin, const float* a_f32, int lda, const __nv_bfloat16* a_hi,
                       const __nv_bfloat16* a_lo, int ld_sp, int M, int act, const LinearOut& out, int tag,
                       cudaStream_t s) {
  Scoped sc(h, s, tag);
  if (use_umma(h)) {
    return launch_linear_umma(a_hi, a_lo, ld_sp, lin.w_hi.as<__nv_bfloat16>(), lin.w_lo.as<__nv_bfloat16>(), lin.K,
                              lin.b.as<float>(), out, M, lin.N, lin.K, act, s);
  }
  return launch_linear_simt(a_f32, lda, lin.w.as<float>(), lin.K, lin.b.as<float>(), out, M, lin.N, lin.K, act, s);
}

cudaError_t upload(DevBuf& dst, const float* src, size_t n) {
  cudaError_t e = dst.ensure(n * sizeof(float));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(dst.p, src, n * sizeof(float), cudaMemcpyDefault);
}

struct WView { const float* data; int64_t rows, cols; };

int forward_impl(da_handle* h, const float* x, const int64_t* t_arr, int t_uniform, float* out, float* alpha_last,
                 int step_mode, const da_step_coef* coef, const float* x_in, const float* noise, cudaStream_t s,
                 const da_schedule* sched = nullptr, float* alpha_all = nullptr) {
  if (alpha_all && !alpha_last) alpha_last = alpha_all + (size_t)(h->L - 1) * (size_t)h->csr.E * h->cfg.heads;   // [L, E, H]
  if (!h->weights_loaded) return h->fail(DA_ERR_MISSING, "weights not loaded (da_load_weights)");
  if (!h->graph_set) return h->fail(DA_ERR_INVALID, "graph not set (da_set_graph)");
  if (!h->feats_set) return h->fail(DA_ERR_INVALID, "features not set (da_set_features)");
  if (alpha_last && h->cfg.attn_mode != DA_ATTN_CSR)
    return h->fail(DA_ERR_UNSUPPORTED, "attention weights are only produced with attn_mode = DA_ATTN_CSR");
  const da_config& c = h->cfg;
  const int Mr = h->num_real, Mt = h->num_total, D = h->D, Hm = c.mlp_hidden, L = h->L;
  const bool umma = use_umma(h);
  const bool fold = h->fold_ready && umma && h->use_plan && h->plan.n_tiles > 0 && h->plan.real_rows_clean && !alpha_last && !alpha_all;
  const int ld_h = h->fold_cfg ? h->Kf : Hm;   // row stride of the split-bf16 trunk hidden
  cudaError_t ce;
#define DA_CK(call, where) do { ce = (call); if (ce != cudaSuccess) return h->cuda_fail(ce, where); } while (0)

  // 1. prologue -> h [Mr, Hm]
  {
    PrologueArgs a{};
    a.x = x; a.t = t_arr; a.t_uniform = t_uniform;
    a.P = h->feats_zero ? nullptr : h->P.as<float>();
    a.b1 = h->b1.as<float>();
    a.pos_w0 = h->pos_w0.as<float>(); a.pos_b0 = h->pos_b0.as<float>();
    a.pos_w2 = h->pos_w2.as<float>(); a.pos_b2 = h->pos_b2.as<float>();
    a.time_emb = h->time_emb.as<float>(); a.w1pt_T = h->w1pt_T.as<float>();
    if (!h->no_pro_table && h->pro_wc.p && h->pro_tt.p) { a.wc_T = h->pro_wc.as<float>(); a.tt = h->pro_tt.as<float>(); }
    a.M = Mr; a.C_in = c.in_channels; a.Hm = Hm; a.T = c.steps;
    a.act = (c.head_kind == DA_HEAD_SE3) ? ACT_LRELU : ACT_GELU;
    a.row_ext = h->use_plan ? h->plan.ext_of_int : nullptr;   // internal node order of the planner (x / t are the caller's)
    if (umma) { a.out.hi = h->h_hi.as<__nv_bfloat16>(); a.out.lo = h->h_lo.as<__nv_bfloat16>(); a.out.ld_split = ld_h; }
    else { a.out.f32 = h->hbuf.as<float>(); a.out.ldc = Hm; }
    Scoped sc(h, s, TAG_PROLOGUE);
    DA_CK(launch_prologue(a, s), "prologue");
  }
  // 2. combined = act2(h @ W2^T + b2) -> [Mr, D] (fp32 kept for the trunk residual); folded away on the folded path
  if (!fold) {
    LinearOut o; o.f32 = h->combined.as<float>(); o.ldc = D;
    if (umma) { o.hi = h->comb_hi.as<__nv_bfloat16>(); o.lo = h->comb_lo.as<__nv_bfloat16>(); o.ld_split = D; }
    DA_CK(run_linear(h, h->mlp2, h->hbuf.as<float>(), Hm, h->h_hi.as<__nv_bfloat16>(), h->h_lo.as<__nv_bfloat16>(), ld_h,
                     Mr, (c.head_kind == DA_HEAD_SE3) ? ACT_LRELU : ACT_NONE, o, TAG_MLP2_GEMM, s),
          "mlp2 gemm");
  }
  // 3. graph-transformer layers
  const float* xin = h->combined.as<float>();
  const __nv_bfloat16* xin_hi = h->comb_hi.as<__nv_bfloat16>();
  const __nv_bfloat16* xin_lo = h->comb_lo.as<__nv_bfloat16>();
  int ld_in = D;
  if (fold) {   // first projection straight from h (and the one-hot columns of the virtual rows): K = Kf instead of D
    xin = nullptr; xin_hi = h->h_hi.as<__nv_bfloat16>(); xin_lo = h->h_lo.as<__nv_bfloat16>(); ld_in = h->Kf;
  }
  for (int l = 0; l < L; ++l) {
    const int HC = layer_hc(h, l), C = HC / c.heads;
    const bool last = (l == L - 1);
    const bool dense = h->use_plan && h->plan.n_tiles > 0;
    const int Cpad = (C + 15) / 16 * 16;
    if (fold && last) {
      // folded last layer: [Q | K | V'] projection, scores on the C-channel images, aggregation of the 32-channel V'
      const int N3 = h->fold3.N;
      LinearOut o; o.f32 = h->qkvs.as<float>(); o.ldc = N3;
      o.img_node_slot = h->plan.node_slot; o.qimg = h->qimg_l.as<__nv_bfloat16>(); o.kimg = h->kimg_l.as<__nv_bfloat16>();
      o.vimg = h->vimg_f.as<__nv_bfloat16>();
      o.img_H = c.heads; o.img_C = C; o.img_Cpad = Cpad; o.img_rows = Mr; o.img_Cv = 32; o.img_Cvpad = 32;
      o.f32_tile_flags = h->plan.f32_tile_flags[1];   // fp32 K / V' rows only where the gather below reads them
      DA_CK(run_linear(h, h->fold3, nullptr, ld_in, xin_hi, xin_lo, ld_in, Mt, ACT_NONE, o, TAG_QKVS_GEMM_LAST, s), "folded qkv gemm");
      if (h->plan.n_extra > 0) {
        PackArgs pa{};
        pa.qkvs = h->qkvs.as<float>(); pa.ld = N3; pa.node_slot = h->plan.node_slot; pa.n = Mr;
        pa.H = c.heads; pa.C = C; pa.Cpad = Cpad; pa.Cv = 32; pa.Cvpad = 32;
        pa.qimg = o.qimg; pa.kimg = o.kimg; pa.vimg = o.vimg;
        Scoped sc(h, s, TAG_PACK_LAST);
        DA_CK(launch_gather_extra(pa, h->plan.x_src, h->plan.x_slot, h->plan.n_extra, s), "gather extra sources");
      }
      AttnFoldArgs fa{};
      fa.qimg = o.qimg; fa.kimg = o.kimg; fa.vimg = o.vimg;
      fa.tiles = h->plan.tiles; fa.n_tiles = h->plan.n_tiles; fa.bitmap = h->plan.bitmap; fa.blk_list = h->plan.blk_list;
      fa.H = c.heads; fa.C = C; fa.Cpad = Cpad; fa.partial = h->partial.as<float>(); fa.persistent = h->fold_persist ? 1 : 0;
      Scoped sc(h, s, TAG_ATTN_DENSE_LAST);
      DA_CK(launch_attn_dense_fold(fa, s), "folded dense attention");
      break;
    }
    __nv_bfloat16* qimg = (last ? h->qimg_l : h->qimg).as<__nv_bfloat16>();
    __nv_bfloat16* kimg = (last ? h->kimg_l : h->kimg).as<__nv_bfloat16>();
    __nv_bfloat16* vimg = (last ? h->vimg_l : h->vimg).as<__nv_bfloat16>();
    // rows of dense tiles with at most DA_FUSE_MAX_RESIDUAL residual in-edges are finalised by the dense kernel itself
    const bool fuse = dense && attn_csr_rows_supported(c.heads, C) && !alpha_last && h->plan.n_fused > 0 && !h->no_fuse &&
                      attn_dense_can_fuse(C);
    {
      LinearOut o; o.f32 = h->qkvs.as<float>(); o.ldc = 4 * HC;
      if (dense && umma) {  // Q / K / V operand images straight from the GEMM epilogue
        o.img_node_slot = h->plan.node_slot; o.qimg = qimg; o.kimg = kimg; o.vimg = vimg;
        o.img_H = c.heads; o.img_C = C; o.img_Cpad = Cpad; o.img_rows = Mr;
        // the flags assume the fused rows take Q from TMEM (plan.cu)
        if (attn_csr_rows_supported(c.heads, C) && !alpha_last && (fuse || h->plan.n_fused == 0))
          o.f32_tile_flags = h->plan.f32_tile_flags[last ? 1 : 0];
      }
      int tag = l == 0 ? TAG_QKVS_GEMM_FIRST : (last ? TAG_QKVS_GEMM_LAST : TAG_QKVS_GEMM_MID);
      DA_CK(run_linear(h, (fold && l == 0) ? h->fold0 : h->layer[l], xin, ld_in, xin_hi, xin_lo, ld_in, Mt, ACT_NONE, o, tag, s), "qkvs gemm");
    }
    const CsrGraph& csr = h->use_plan ? h->plan.residual : h->csr;
    bool gather_pending = false;
    PackArgs gather_args{};
    if (dense) {  // bitmap edges on the tensor cores; the CSR kernel below continues with the residual edges
      PackArgs pa{};
      pa.qkvs = h->qkvs.as<float>(); pa.ld = 4 * HC; pa.node_slot = h->plan.node_slot; pa.n = Mr;
      pa.H = c.heads; pa.C = C; pa.Cpad = Cpad;
      pa.qimg = qimg; pa.kimg = kimg; pa.vimg = vimg;
      if (!umma) {  // exact-fp32 GEMM mode: separate repack pass
        Scoped sc(h, s, last ? TAG_PACK_LAST : TAG_PACK_HIDDEN);
        DA_CK(launch_pack_images(pa, s), "pack images");
      }
      // copies of the promoted residual sources' K / V rows into the padding image rows: a launch of its own, unless it
      // can ride with the rows-outside-the-tiles launch of the persistent path (decided below)
      gather_pending = h->plan.n_extra > 0;
      gather_args = pa;
    }
    AttnCsrArgs a{};
    a.qkvs = h->qkvs.as<float>(); a.ld = 4 * HC;
    a.rowptr = csr.rowptr; a.col = csr.col; a.weight = csr.weight;
    if (dense) { a.init_acc = h->dacc.as<float>(); a.init_stats = h->dstats.as<float>(); a.init_slot = h->plan.node_slot; }
    a.H = c.heads; a.C = C;
    a.act = (!last && c.arch == DA_ARCH_TRANSFORMER) ? ACT_GELU : ACT_NONE;
    if (last) {
      a.n_targets = alpha_last ? Mt : Mr;
      a.resid = h->combined.as<float>(); a.ld_resid = D;  // trunk residual feats + combined (efficient_gat.py:145)
      if (umma) { a.out.hi = h->r_hi.as<__nv_bfloat16>(); a.out.lo = h->r_lo.as<__nv_bfloat16>(); a.out.ld_split = D; }
      else { a.out.f32 = h->r.as<float>(); a.out.ldc = D; }
    } else {
      a.n_targets = Mt;
      DevBuf& yb = (l & 1) ? h->xb : h->xa;
      DevBuf& yh = (l & 1) ? h->xb_hi : h->xa_hi;
      DevBuf& yl = (l & 1) ? h->xb_lo : h->xa_lo;
      if (umma) { a.out.hi = yh.as<__nv_bfloat16>(); a.out.lo = yl.as<__nv_bfloat16>(); a.out.ld_split = HC; }
      else { a.out.f32 = yb.as<float>(); a.out.ldc = HC; }
      xin = yb.as<float>(); xin_hi = yh.as<__nv_bfloat16>(); xin_lo = yl.as<__nv_bfloat16>(); ld_in = HC;
    }
    float* alpha_l = last ? alpha_last : (alpha_all ? alpha_all + (size_t)l * (size_t)h->csr.E * c.heads : nullptr);
    if (alpha_l) {   // attention weights of this layer in the caller's edge order (CSR mode only, checked above)
      DA_CK(h->scores.ensure((size_t)(h->csr.E > 0 ? h->csr.E : 1) * c.heads * sizeof(float)), "alloc scores");
      DA_CK(h->stats.ensure((size_t)Mt * c.heads * 2 * sizeof(float)), "alloc stats");
      a.scores = h->scores.as<float>(); a.stats = h->stats.as<float>();
    }
    const bool rows_path = h->use_plan && attn_csr_rows_supported(c.heads, C) && !a.scores;
    AttnCsrArgs hv = a, lt = a;
    if (rows_path) {
      // residual edges: low-degree rows on the warp-per-node kernel, heavy rows (virtual nodes) edge-parallel
      const DensePlan& pl = h->plan;
      hv.node_list = pl.heavy; hv.n_targets = last ? pl.n_heavy_real : pl.n_heavy;
      if (dense && umma) { hv.img_slot = pl.node_slot; hv.kimg = kimg; hv.vimg = vimg; hv.img_Cpad = Cpad; }
      if (fuse) { lt.node_list = pl.light_nf; lt.n_targets = last ? pl.n_light_nf_real : pl.n_light_nf; }
      else { lt.node_list = pl.light; lt.n_targets = last ? pl.n_light_real : pl.n_light; }
    }
    // hidden layers (32-channel heads), every CSR-served row outside the tiles: one lane-per-edge launch
    const bool vrows = rows_path && dense && fuse && !last && h->plan.csr_rows_independent && attn_csr_vrows_supported(c.heads, C) &&
                       !(dense && umma && Cpad != 32) && !h->no_vrows;
    bool gather_rides = false;
    auto launch_rows = [&](cudaStream_t st) -> cudaError_t {
      cudaError_t e = cudaSuccess;
      if (vrows) {
        if (h->plan.n_csr_rows > 0) {
          AttnCsrArgs vr = hv;
          vr.node_list = h->plan.csr_rows; vr.n_targets = h->plan.n_csr_rows; vr.n_coop = h->plan.n_heavy;   // hubs first
          vr.init_acc = nullptr; vr.init_stats = nullptr; vr.init_slot = nullptr;
          if (gather_rides) {
            vr.gx_n = h->plan.n_extra; vr.gx_src = h->plan.x_src; vr.gx_slot = h->plan.x_slot;
            vr.gx_kimg = gather_args.kimg; vr.gx_vimg = gather_args.vimg;
          }
          Scoped sc(h, st, TAG_ATTN_HIDDEN);
          e = launch_attn_csr_vrows(vr, st);
        }
        return e;
      }
      if (hv.n_targets > 0) {
        Scoped sc(h, st, last ? TAG_ATTN_LAST : TAG_ATTN_HIDDEN);
        e = launch_attn_csr_heavy(hv, st);
      }
      if (e == cudaSuccess && lt.n_targets > 0) {
        Scoped sc(h, st, last ? TAG_ATTN_LAST : TAG_ATTN_HIDDEN);
        e = launch_attn_csr_rows(lt, st);
      }
      return e;
    };
    // With every in-tile row finalised by the dense kernel, the CSR kernels only serve rows outside the tiles
    // (virtual nodes): they depend on the GEMM alone and run on a side stream next to the dense kernel.
    // Persistent form of the hidden-layer kernel (one CTA per SM, every register of it): nothing can run next to it, so the
    // rows outside the tiles go first on the same stream.
    AttnDenseArgs da_{};
    if (dense) {
      da_.qimg = qimg; da_.kimg = kimg; da_.vimg = vimg;
      da_.tiles = h->plan.tiles; da_.n_tiles = h->plan.n_tiles; da_.bitmap = h->plan.bitmap; da_.blk_list = h->plan.blk_list;
      da_.H = c.heads; da_.C = C; da_.Cpad = Cpad;
      da_.acc = h->dacc.as<float>(); da_.stats = h->dstats.as<float>();
      da_.dbg = (l == h->dbg_layer) ? (long long*)h->dbg_trace : nullptr;
      da_.stagger_ns = h->hidden_stagger_ns;
      if (fuse) {
        da_.row_fused = h->plan.row_fused;
        da_.qkvs = a.qkvs; da_.ld = a.ld; da_.n_rows = Mt; da_.n_rows_resid = Mt;   // combined has Mt rows too
        da_.n_rows_out = last ? Mr : Mt;   // r_hi / r_lo are [Mr, D], xa / xb planes [Mt, 256]
        da_.rowptr = csr.rowptr; da_.col = csr.col; da_.weight = csr.weight;
        da_.resid = a.resid; da_.ld_resid = a.ld_resid; da_.act = a.act; da_.out = a.out;
      }
    }
    const bool persist = rows_path && dense && fuse && umma && !last && h->hidden_persist && h->plan.real_rows_clean &&
                         h->plan.csr_rows_independent && attn_hidden_persist_supported(da_);
    const bool side_rows = rows_path && dense && fuse && h->plan.csr_rows_independent && !h->no_side && !persist &&
                           (hv.n_targets > 0 || lt.n_targets > 0);
    if (side_rows) {
      if (!h->side) {
        // highest priority: its few CTAs take the next free slots instead of queueing behind the dense grid's 2048
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        DA_CK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_hi), "side stream");
        DA_CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming), "side stream");
        DA_CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming), "side stream");
      }
      DA_CK(cudaEventRecord(h->ev_fork, s), "fork");   // the GEMM (and the gather) of this layer
    }
    gather_rides = gather_pending && persist && vrows && h->plan.n_csr_rows > 0 && umma && Cpad == 32 && !h->no_gather_ride;
    if (gather_pending && !gather_rides) {
      Scoped sc(h, s, last ? TAG_PACK_LAST : TAG_PACK_HIDDEN);
      DA_CK(launch_gather_extra(gather_args, h->plan.x_src, h->plan.x_slot, h->plan.n_extra, s), "gather extra sources");
    }
    if (persist && rows_path) DA_CK(launch_rows(s), "graph attention (rows outside the dense tiles)");
    if (dense) {
      Scoped sc(h, s, last ? TAG_ATTN_DENSE_LAST : TAG_ATTN_DENSE_HIDDEN);
      if (persist) { h->persist_launches++; DA_CK(launch_attn_hidden_persist(da_, s), "dense attention (persistent)"); }
      else DA_CK(launch_attn_dense(da_, s), "dense attention");
    }
    if (side_rows) {   // launched AFTER the dense kernel so that its CTAs fill the slots the dense grid leaves in its tail
      DA_CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0), "fork");
      DA_CK(launch_rows(h->side), "graph attention (rows outside the dense tiles)");
      DA_CK(cudaEventRecord(h->ev_join, h->side), "join");
    }
    if (side_rows) {
      DA_CK(cudaStreamWaitEvent(s, h->ev_join, 0), "join");
    } else if (rows_path) {
      if (!persist) DA_CK(launch_rows(s), "graph attention (residual rows)");
    } else {
      Scoped sc(h, s, last ? TAG_ATTN_LAST : TAG_ATTN_HIDDEN);
      DA_CK(launch_attn_csr(a, s), "graph attention");
    }
    if (alpha_l) {
      Scoped sc(h, s, TAG_OTHER);
      DA_CK(launch_alpha_normalize(h->scores.as<float>(), h->stats.as<float>(), h->csr.rowptr, h->csr.eid, Mt,
                                   c.heads, alpha_l, s), "alpha normalize");
    }
  }
  // 4. head: u = GELU(r @ Wa^T + ba) ; then final linear(s) + pose map + sampler update
  {
    HeadFinalArgs a{};
    a.u = h->u.as<float>(); a.Nh = h->Nh;
    a.w_b = h->headb_w.as<float>(); a.b_b = h->headb_b.as<float>();
    a.w_r = h->headr_w.as<float>(); a.b_r = h->headr_b.as<float>();
    a.M = Mr; a.C_out = c.out_channels; a.head_kind = c.head_kind;
    a.step_mode = step_mode;
    if (coef) a.coef = *coef;
    a.x_in = x_in; a.noise = noise; a.out = out;
    a.row_ext = h->use_plan ? h->plan.ext_of_int : nullptr;
    if (sched != nullptr && t_arr != nullptr && step_mode != STEP_NONE) {   // per-node schedule coefficients
      DA_CK(h->tmin.ensure(sizeof(int32_t)), "alloc");
      h->launches++;
      DA_CK(launch_min_t(t_arr, Mr, h->tmin.as<int32_t>(), s), "min t");
      a.tabs.sched = *sched; a.tabs.t = t_arr; a.tabs.tmin = h->tmin.as<int32_t>();
    }
    // 2-D head on the tensor-core path: final_mlp[2] + the sampler update ride in the GEMM epilogue (one launch)
    if (fold) {   // h (W_a W_2)^T + x3 (W_a W_skip)^T + per-head aggregates -> GELU -> final_mlp[2] -> sampler update
      HeadFoldArgs f{};
      f.partial = h->partial.as<float>(); f.H = c.heads;
      f.h_hi = h->h_hi.as<__nv_bfloat16>(); f.h_lo = h->h_lo.as<__nv_bfloat16>(); f.ld_h = h->Kf; f.Hm = Hm;
      f.x_hi = xin_hi; f.x_lo = xin_lo; f.ld_x = ld_in; f.hid = c.hidden;
      f.w_hi = h->fold_wt_hi.as<__nv_bfloat16>(); f.w_lo = h->fold_wt_lo.as<__nv_bfloat16>(); f.bias = h->fold_b.as<float>();
      f.fin = a;
      Scoped sc(h, s, TAG_HEAD_FINAL);
      DA_CK(launch_head_fold(f, s), "folded head");
      return DA_OK;
    }
    const bool fused_head = umma && c.head_kind == DA_HEAD_2D && h->Nh == 32 && !h->no_head_fuse;
    LinearOut o;
    if (fused_head) o.head = &a; else { o.f32 = h->u.as<float>(); o.ldc = h->Nh; }
    DA_CK(run_linear(h, h->head1, h->r.as<float>(), D, h->r_hi.as<__nv_bfloat16>(), h->r_lo.as<__nv_bfloat16>(), D, Mr,
                     ACT_GELU, o, TAG_HEAD_GEMM, s), "head gemm");
    if (!fused_head) {
      Scoped sc(h, s, TAG_HEAD_FINAL);
      DA_CK(launch_head_final(a, s), "head final");
    }
  }
#undef DA_CK
  return DA_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

int da_abi_version(void) { return DA_ABI_VERSION; }

const char* da_last_error(const da_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int da_create(da_handle** out, const da_config* cfg) {
  if (!out || !cfg) { g_create_error = "null argument"; return DA_ERR_INVALID; }
  *out = nullptr;
  if (cfg->abi_version != DA_ABI_VERSION) { g_create_error = "abi_version mismatch"; return DA_ERR_INVALID; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0") +
                     " (this library has no CPU fallback)";
    return DA_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return DA_ERR_INVALID; }
  cudaDeviceProp prop{};
  ce = cudaGetDeviceProperties(&prop, cfg->device);
  if (ce != cudaSuccess) { g_create_error = cudaGetErrorString(ce); return DA_ERR_CUDA; }
  if (prop.major != 10) {
    g_create_error = "device is not sm_100 (B200); kernels are built for sm_100a only";
    return DA_ERR_CUDA;
  }
  const da_config& c = *cfg;
  if (c.heads <= 0 || c.hidden <= 0 || c.hidden % c.heads || c.n_layers < 2 || c.feat_dim <= 0 || c.steps <= 0 ||
      c.in_channels <= 0 || c.in_channels > 8 || c.out_channels <= 0 || c.mlp_hidden <= 0) {
    g_create_error = "invalid configuration value";
    return DA_ERR_INVALID;
  }
  const int D = c.feat_dim + 64;  // efficient_gat.py:48 (Dv + 32 + 32)
  if (D % c.heads) { g_create_error = "combined feature dim must be divisible by heads"; return DA_ERR_INVALID; }
  if (c.head_kind == DA_HEAD_SE3 && c.out_channels != 7) { g_create_error = "SE3 head has 7 output channels"; return DA_ERR_INVALID; }
  if (c.head_kind != DA_HEAD_SE3 && c.head_kind != DA_HEAD_2D) { g_create_error = "unknown head_kind"; return DA_ERR_INVALID; }
  if (c.gemm_mode == DA_GEMM_BF16X3_UMMA) {
    // tensor-core tiles: K in multiples of 64 bf16 (one 128-byte swizzle row), N in multiples of 16
    if (c.feat_dim % 64 || c.mlp_hidden % 64 || c.hidden % 64 || D % 64) {
      g_create_error = "DA_GEMM_BF16X3_UMMA needs feat_dim, mlp_hidden, hidden and D to be multiples of 64";
      return DA_ERR_UNSUPPORTED;
    }
  } else if (c.gemm_mode != DA_GEMM_FP32_SIMT) { g_create_error = "unknown gemm_mode"; return DA_ERR_INVALID; }
  if ((D % 4) || (c.feat_dim % 4) || (c.mlp_hidden % 4) || (c.hidden % 4)) {
    g_create_error = "feature widths must be multiples of 4";
    return DA_ERR_UNSUPPORTED;
  }
  if ((D / c.heads + 31) / 32 > 13) { g_create_error = "head dim above 416 not supported"; return DA_ERR_UNSUPPORTED; }
  cudaSetDevice(c.device);
  da_handle* h = new da_handle();
  h->cfg = c;
  h->D = D;
  h->L = c.n_layers;
  h->Nh = (c.head_kind == DA_HEAD_SE3) ? 512 : 32;
  h->layer.resize(h->L);
  {
    const int c_last = D / c.heads, cpad_last = (c_last + 15) / 16 * 16;
    const bool virt = c.arch == DA_ARCH_EXOPHORMER && c.virt_nodes > 0;
    h->fold_cfg = !h->no_fold && c.gemm_mode == DA_GEMM_BF16X3_UMMA && c.attn_mode == DA_ATTN_AUTO && c.head_kind == DA_HEAD_2D &&
                  h->Nh == 32 && c.virt_nodes <= 64 && c_last % 8 == 0 && (c.heads * c_last) % 32 == 0 &&
                  (2 * D + c.heads * 32) % 128 == 0 && c.hidden % 64 == 0 && c.mlp_hidden % 16 == 0 && attn_dense_fold_supported(cpad_last);
    h->Kf = c.mlp_hidden + (h->fold_cfg && virt ? 64 : 0);
  }
  *out = h;
  return DA_OK;
}

void da_destroy(da_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  drain_profile(h);
  h->hoist.release(); h->mlp2.release(); h->head1.release(); h->fold0.release(); h->fold3.release();
  h->fold_wt.release(); h->fold_wt_hi.release(); h->fold_wt_lo.release(); h->fold_b.release(); h->fold_g.release(); h->vimg_f.release(); h->partial.release();
  for (auto& l : h->layer) l.release();
  DevBuf* all[] = {&h->pos_w0, &h->pos_b0, &h->pos_w2, &h->pos_b2, &h->time_emb, &h->w1pt_T, &h->b1, &h->headb_w,
                   &h->headb_b, &h->headr_w, &h->headr_b, &h->virt_emb, &h->P, &h->hbuf, &h->combined, &h->qkvs,
                   &h->xa, &h->xb, &h->r, &h->u, &h->model_out, &h->scores, &h->stats, &h->feats_sp_hi,
                   &h->feats_sp_lo, &h->h_hi, &h->h_lo, &h->comb_hi, &h->comb_lo, &h->xa_hi, &h->xa_lo, &h->xb_hi,
                   &h->xb_lo, &h->r_hi, &h->r_lo, &h->qimg, &h->kimg, &h->vimg, &h->qimg_l, &h->kimg_l, &h->vimg_l, &h->dacc, &h->dstats,
                   &h->feats_perm, &h->tmin, &h->pro_wc, &h->pro_tt};
  for (auto* b : all) b->release();
  free_csr(&h->csr);
  free_plan(&h->plan);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
}

int da_load_weights(da_handle* h, const da_weight_desc* w, int32_t n) {
  if (!h || !w || n <= 0) return h ? h->fail(DA_ERR_INVALID, "null weights") : DA_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  const da_config& c = h->cfg;
  std::map<std::string, WView> m;
  for (int i = 0; i < n; ++i) {
    if (!w[i].name || !w[i].data) return h->fail(DA_ERR_INVALID, "weight descriptor with null name/data");
    m[w[i].name] = WView{w[i].data, w[i].rows, w[i].cols};
  }
  std::string missing;
  auto need = [&](const std::string& name, int64_t rows, int64_t cols) -> const WView* {
    auto it = m.find(name);
    if (it == m.end()) { missing = "missing weight '" + name + "'"; return nullptr; }
    if (it->second.rows != rows || it->second.cols != cols) {
      char buf[256];
      snprintf(buf, sizeof buf, "weight '%s' has shape [%lld,%lld], expected [%lld,%lld]", name.c_str(),
               (long long)it->second.rows, (long long)it->second.cols, (long long)rows, (long long)cols);
      missing = buf;
      return nullptr;
    }
    return &it->second;
  };
  // Stage everything on the host first (weights may live in host or device memory).
  auto fetch = [&](const WView* v, std::vector<float>& dst) -> cudaError_t {
    dst.resize((size_t)v->rows * v->cols);
    return cudaMemcpy(dst.data(), v->data, dst.size() * sizeof(float), cudaMemcpyDefault);
  };
  cudaError_t ce;
#define DA_NEED(var, name, r, cc) const WView* var = need(name, r, cc); if (!var) return h->fail(DA_ERR_MISSING, missing)
#define DA_CK(call) do { ce = (call); if (ce != cudaSuccess) return h->cuda_fail(ce, "da_load_weights"); } while (0)
  const int D = h->D, Dv = c.feat_dim, Hm = c.mlp_hidden, Cin = c.in_channels;
  std::vector<float> tmp, tmp2;
  auto up_direct = [&](DevBuf& dst, const WView* v) -> cudaError_t {
    cudaError_t e = fetch(v, tmp);
    if (e != cudaSuccess) return e;
    return upload(dst, tmp.data(), tmp.size());
  };
  auto finish_linear = [&](Linear& lin, const std::vector<float>& wv, const std::vector<float>& bv, int N, int K) -> cudaError_t {
    lin.N = N; lin.K = K;
    cudaError_t e = upload(lin.w, wv.data(), wv.size());
    if (e != cudaSuccess) return e;
    e = upload(lin.b, bv.data(), bv.size());
    if (e != cudaSuccess) return e;
    if (use_umma(h)) {
      e = lin.w_hi.ensure(wv.size() * sizeof(__nv_bfloat16)); if (e != cudaSuccess) return e;
      e = lin.w_lo.ensure(wv.size() * sizeof(__nv_bfloat16)); if (e != cudaSuccess) return e;
      e = launch_split_bf16(lin.w.as<float>(), K, lin.w_hi.as<__nv_bfloat16>(), lin.w_lo.as<__nv_bfloat16>(), K, N, K, 0);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };

  std::vector<float> pw2, pb2, temb;   // kept for the composed prologue tables below
  {
    DA_NEED(v0, "pos_mlp.0.weight", 16, Cin); DA_CK(up_direct(h->pos_w0, v0));
    DA_NEED(v1, "pos_mlp.0.bias", 16, 1); DA_CK(up_direct(h->pos_b0, v1));
    DA_NEED(v2, "pos_mlp.2.weight", 32, 16); DA_CK(up_direct(h->pos_w2, v2)); DA_CK(fetch(v2, pw2));
    DA_NEED(v3, "pos_mlp.2.bias", 32, 1); DA_CK(up_direct(h->pos_b2, v3)); DA_CK(fetch(v3, pb2));
    DA_NEED(v4, "time_emb.weight", c.steps, 32); DA_CK(up_direct(h->time_emb, v4)); DA_CK(fetch(v4, temb));
  }
  {  // mlp[0]: split into the step-invariant feature block and the 64 pose+time columns
    DA_NEED(vw, "mlp.0.weight", Hm, D);
    DA_NEED(vb, "mlp.0.bias", Hm, 1);
    DA_CK(fetch(vw, tmp));
    std::vector<float> wf((size_t)Hm * Dv), wpt((size_t)64 * Hm), bias;
    for (int n_ = 0; n_ < Hm; ++n_) {
      memcpy(&wf[(size_t)n_ * Dv], &tmp[(size_t)n_ * D], sizeof(float) * Dv);
      for (int k = 0; k < 64; ++k) wpt[(size_t)k * Hm + n_] = tmp[(size_t)n_ * D + Dv + k];
    }
    DA_CK(fetch(vb, bias));
    DA_CK(finish_linear(h->hoist, wf, bias, Hm, Dv));
    DA_CK(upload(h->w1pt_T, wpt.data(), wpt.size()));
    DA_CK(upload(h->b1, bias.data(), bias.size()));
    {  // table form of the prologue (pointwise.cu: prologue_table_kernel), composed in fp64
      std::vector<float> wc((size_t)16 * Hm), ttab((size_t)c.steps * Hm);
      for (int n_ = 0; n_ < Hm; ++n_) {
        const float* wrow = &tmp[(size_t)n_ * D + Dv];   // [pos 32 | time 32] columns of W1 row n_
        for (int u = 0; u < 16; ++u) {
          double acc = 0.0;
          for (int v = 0; v < 32; ++v) acc += (double)wrow[v] * (double)pw2[(size_t)v * 16 + u];
          wc[(size_t)u * Hm + n_] = (float)acc;
        }
        double bc = 0.0;
        for (int v = 0; v < 32; ++v) bc += (double)wrow[v] * (double)pb2[v];
        for (int t_ = 0; t_ < c.steps; ++t_) {
          double acc = bc;
          for (int v = 0; v < 32; ++v) acc += (double)wrow[32 + v] * (double)temb[(size_t)t_ * 32 + v];
          ttab[(size_t)t_ * Hm + n_] = (float)acc;
        }
      }
      DA_CK(upload(h->pro_wc, wc.data(), wc.size()));
      DA_CK(upload(h->pro_tt, ttab.data(), ttab.size()));
    }
  }
  {
    DA_NEED(vw, "mlp.2.weight", D, Hm);
    DA_NEED(vb, "mlp.2.bias", D, 1);
    DA_CK(fetch(vw, tmp)); DA_CK(fetch(vb, tmp2));
    DA_CK(finish_linear(h->mlp2, tmp, tmp2, D, Hm));
  }
  for (int l = 0; l < h->L; ++l) {  // [Q | K | V | skip] concatenated along the output dim
    const int in = layer_in(h, l), HC = layer_hc(h, l);
    std::vector<float> wcat((size_t)4 * HC * in), bcat((size_t)4 * HC);
    const char* parts[4] = {"lin_query", "lin_key", "lin_value", "lin_skip"};
    for (int p = 0; p < 4; ++p) {
      std::string base = "gnn_backbone.module_list." + std::to_string(l) + "." + parts[p];
      DA_NEED(vw, base + ".weight", HC, in);
      DA_NEED(vb, base + ".bias", HC, 1);
      DA_CK(fetch(vw, tmp)); DA_CK(fetch(vb, tmp2));
      memcpy(&wcat[(size_t)p * HC * in], tmp.data(), tmp.size() * sizeof(float));
      memcpy(&bcat[(size_t)p * HC], tmp2.data(), tmp2.size() * sizeof(float));
    }
    DA_CK(finish_linear(h->layer[l], wcat, bcat, 4 * HC, in));
  }
  if (c.arch == DA_ARCH_EXOPHORMER && c.virt_nodes > 0) {
    DA_NEED(v, "gnn_backbone.virt_node_embedding.weight", c.virt_nodes, D);
    DA_CK(up_direct(h->virt_emb, v));
  }
  if (c.head_kind == DA_HEAD_2D) {
    DA_NEED(vw, "final_mlp.0.weight", 32, D);
    DA_NEED(vb, "final_mlp.0.bias", 32, 1);
    DA_CK(fetch(vw, tmp)); DA_CK(fetch(vb, tmp2));
    DA_CK(finish_linear(h->head1, tmp, tmp2, 32, D));
    DA_NEED(v2, "final_mlp.2.weight", c.out_channels, 32); DA_CK(up_direct(h->headb_w, v2));
    DA_NEED(v3, "final_mlp.2.bias", c.out_channels, 1); DA_CK(up_direct(h->headb_b, v3));
  } else {
    DA_NEED(vtw, "mlp_t.0.weight", 256, D);
    DA_NEED(vtb, "mlp_t.0.bias", 256, 1);
    DA_NEED(vrw, "mlp_r.0.weight", 256, D);
    DA_NEED(vrb, "mlp_r.0.bias", 256, 1);
    std::vector<float> wcat((size_t)512 * D), bcat(512);
    DA_CK(fetch(vtw, tmp)); memcpy(wcat.data(), tmp.data(), tmp.size() * sizeof(float));
    DA_CK(fetch(vrw, tmp)); memcpy(wcat.data() + (size_t)256 * D, tmp.data(), tmp.size() * sizeof(float));
    DA_CK(fetch(vtb, tmp)); memcpy(bcat.data(), tmp.data(), 256 * sizeof(float));
    DA_CK(fetch(vrb, tmp)); memcpy(bcat.data() + 256, tmp.data(), 256 * sizeof(float));
    DA_CK(finish_linear(h->head1, wcat, bcat, 512, D));
    DA_NEED(v2, "mlp_t.2.weight", 3, 256); DA_CK(up_direct(h->headb_w, v2));
    DA_NEED(v3, "mlp_t.2.bias", 3, 1); DA_CK(up_direct(h->headb_b, v3));
    DA_NEED(v4, "mlp_r.2.weight", 3, 256); DA_CK(up_direct(h->headr_w, v4));
    DA_NEED(v5, "mlp_r.2.bias", 3, 1); DA_CK(up_direct(h->headr_b, v5));
  }
  h->fold_ready = false;
  if (h->fold_cfg) {
    // folded products in fp64 on the device (fold.cu): first projection from h, last layer's V' / skip / trunk residual
    // pulled through final_mlp[0]
    const int hid = c.hidden, H = c.heads, Cl = D / H, L = h->L, Kf = h->Kf, N0 = 4 * hid, N3 = 2 * D + H * 32;
    const float* W0 = h->layer[0].w.as<float>(); const float* b0 = h->layer[0].b.as<float>();
    const float* W2 = h->mlp2.w.as<float>(); const float* b2 = h->mlp2.b.as<float>();
    const float* W3 = h->layer[L - 1].w.as<float>(); const float* b3 = h->layer[L - 1].b.as<float>();
    const float* Wa = h->head1.w.as<float>(); const float* ba = h->head1.b.as<float>();
    auto alloc_linear = [&](Linear& lin, int N, int K) -> cudaError_t {
      lin.N = N; lin.K = K;
      cudaError_t e = lin.w.ensure((size_t)N * K * sizeof(float)); if (e != cudaSuccess) return e;
      e = lin.b.ensure((size_t)N * sizeof(float)); if (e != cudaSuccess) return e;
      e = lin.w_hi.ensure((size_t)N * K * sizeof(__nv_bfloat16)); if (e != cudaSuccess) return e;
      e = lin.w_lo.ensure((size_t)N * K * sizeof(__nv_bfloat16)); if (e != cudaSuccess) return e;
      return cudaMemset(lin.w.p, 0, (size_t)N * K * sizeof(float));
    };
    DA_CK(alloc_linear(h->fold0, N0, Kf));
    DA_CK(alloc_linear(h->fold3, N3, hid));
    DA_CK(h->fold_g.ensure((size_t)N0 * sizeof(float)));
    DA_CK(h->fold_wt.ensure((size_t)(Hm + hid) * 32 * sizeof(float)));
    DA_CK(h->fold_b.ensure(32 * sizeof(float)));
    float* F0 = h->fold0.w.as<float>(); float* g = h->fold_g.as<float>();
    DA_CK(launch_matmul_f64(W0, D, W2, Hm, 1, nullptr, 0, 0, 0.f, F0, Kf, 1, N0, Hm, D, 0));               // W_0 W_2
    DA_CK(launch_matmul_f64(W0, D, b2, 1, 0, nullptr, 0, 0, 0.f, g, 1, 0, N0, 1, D, 0));                    // g = W_0 b_2
    DA_CK(launch_matmul_f64(W0, D, b2, 1, 0, b0, 1, 0, 1.f, h->fold0.b.as<float>(), 1, 0, N0, 1, D, 0));    // W_0 b_2 + b_0
    if (Kf > Hm && c.virt_nodes > 0)   // one-hot columns of the virtual rows: W_0 (virt_emb[v] - b_2)
      DA_CK(launch_matmul_f64(W0, D, h->virt_emb.as<float>(), 1, D, g, 1, 0, -1.f, F0 + Hm, Kf, 1, N0, c.virt_nodes, D, 0));
    float* F3 = h->fold3.w.as<float>(); float* fb3 = h->fold3.b.as<float>();
    DA_CK(cudaMemcpy(F3, W3, (size_t)2 * D * hid * sizeof(float), cudaMemcpyDeviceToDevice));   // Q and K rows unchanged
    DA_CK(cudaMemcpy(fb3, b3, (size_t)2 * D * sizeof(float), cudaMemcpyDeviceToDevice));
    for (int hh = 0; hh < H; ++hh) {   // V'_h = W_a[:, head block] W_v[head block, :]
      DA_CK(launch_matmul_f64(Wa + hh * Cl, D, W3 + (size_t)(2 * D + hh * Cl) * hid, hid, 1, nullptr, 0, 0, 0.f,
                              F3 + (size_t)(2 * D + hh * 32) * hid, hid, 1, 32, hid, Cl, 0));
      DA_CK(launch_matmul_f64(Wa + hh * Cl, D, b3 + 2 * D + hh * Cl, 1, 0, nullptr, 0, 0, 0.f, fb3 + 2 * D + hh * 32, 1, 0, 32, 1, Cl, 0));
    }
    float* wt = h->fold_wt.as<float>(); float* fb = h->fold_b.as<float>();
    const int Kt = Hm + hid;   // [32, Kt] = [W_a W_2 | W_a W_skip]
    DA_CK(h->fold_wt_hi.ensure((size_t)Kt * 32 * sizeof(__nv_bfloat16)));
    DA_CK(h->fold_wt_lo.ensure((size_t)Kt * 32 * sizeof(__nv_bfloat16)));
    DA_CK(launch_matmul_f64(Wa, D, W2, Hm, 1, nullptr, 0, 0, 0.f, wt, Kt, 1, 32, Hm, D, 0));
    DA_CK(launch_matmul_f64(Wa, D, W3 + (size_t)3 * D * hid, hid, 1, nullptr, 0, 0, 0.f, wt + Hm, Kt, 1, 32, hid, D, 0));
    DA_CK(launch_split_bf16(wt, Kt, h->fold_wt_hi.as<__nv_bfloat16>(), h->fold_wt_lo.as<__nv_bfloat16>(), Kt, 32, Kt, 0));
    DA_CK(launch_matmul_f64(Wa, D, b3 + 3 * D, 1, 0, ba, 1, 0, 1.f, fb, 1, 0, 32, 1, D, 0));    // b_a + W_a b_skip
    DA_CK(launch_matmul_f64(Wa, D, b2, 1, 0, fb, 1, 0, 1.f, fb, 1, 0, 32, 1, D, 0));            // ... + W_a b_2
    DA_CK(launch_split_bf16(F0, Kf, h->fold0.w_hi.as<__nv_bfloat16>(), h->fold0.w_lo.as<__nv_bfloat16>(), Kf, N0, Kf, 0));
    DA_CK(launch_split_bf16(F3, hid, h->fold3.w_hi.as<__nv_bfloat16>(), h->fold3.w_lo.as<__nv_bfloat16>(), hid, N3, hid, 0));
    h->fold_ready = true;
  }
  DA_CK(cudaDeviceSynchronize());
#undef DA_NEED
#undef DA_CK
  h->weights_loaded = true;
  h->feats_set = false;  // the hoisted product depends on the weights
  return DA_OK;
}

int da_set_graph(da_handle* h, const int64_t* edge_src, const int64_t* edge_dst, int64_t E, const int64_t* batch,
                 int32_t num_real, int32_t num_total, const int32_t* virt_ids, void* stream) {
  if (!h) return DA_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  if (num_real <= 0 || num_total < num_real || E < 0) return h->fail(DA_ERR_INVALID, "bad graph sizes");
  if (E > 0 && (!edge_src || !edge_dst)) return h->fail(DA_ERR_INVALID, "null edge arrays");
  if (num_total > num_real && (!virt_ids || h->cfg.arch != DA_ARCH_EXOPHORMER))
    return h->fail(DA_ERR_INVALID, "virtual rows need arch = EXOPHORMER and virt_ids");
  if (num_total > num_real && !h->weights_loaded)
    return h->fail(DA_ERR_MISSING, "load weights before da_set_graph when virtual nodes are used");
  const da_config& c = h->cfg;
  const int D = h->D, Hm = c.mlp_hidden, hid = c.hidden;
  const char* why = "";
  cudaError_t ce;
  // dense tensor-core tiles need head dims that are multiples of 8 and small enough for shared memory
  const int c_hid = hid / c.heads, c_last = D / c.heads;
  const int cpad_max = ((c_hid > c_last ? c_hid : c_last) + 15) / 16 * 16;
  h->use_plan = (c.attn_mode == DA_ATTN_AUTO) && batch != nullptr && (c_hid % 8 == 0) && (c_last % 8 == 0) && cpad_max <= 144;
  free_csr(&h->csr, s);
  free_plan(&h->plan, s);
  if (h->use_plan) ce = build_dense_plan(edge_src, edge_dst, E, batch, num_real, num_total, &h->plan, s, &why, /*allow_reorder=*/true);
  else ce = build_csr(edge_src, edge_dst, E, num_total, &h->csr, s, &why);
  if (ce != cudaSuccess) {
    if (ce == cudaErrorInvalidValue && why[0]) return h->fail(DA_ERR_INVALID, why);
    return h->cuda_fail(ce, why);
  }
  h->num_real = num_real; h->num_total = num_total;
  const size_t Mt = num_total, Mr = num_real;
  const bool umma = use_umma(h);
#define DA_CK(call) do { ce = (call); if (ce != cudaSuccess) return h->cuda_fail(ce, "da_set_graph alloc"); } while (0)
  DA_CK(h->P.ensure(Mr * Hm * sizeof(float)));
  DA_CK(h->combined.ensure(Mt * D * sizeof(float)));
  DA_CK(h->qkvs.ensure(Mt * 4 * (size_t)(D > hid ? D : hid) * sizeof(float)));
  DA_CK(h->u.ensure(Mr * h->Nh * sizeof(float)));
  DA_CK(h->model_out.ensure(Mr * c.out_channels * sizeof(float)));
  if (umma) {
    const size_t b2 = sizeof(__nv_bfloat16);
    if (h->fold_cfg) {
      // trunk hidden with row stride Kf and one row per node INCLUDING the virtual ones: columns >= Hm are the one-hot
      // selectors of the folded first projection (zero on real rows), see fold.cu
      const size_t hb = Mt * (size_t)h->Kf * b2;
      DA_CK(h->h_hi.ensure(hb)); DA_CK(h->h_lo.ensure(hb));
      DA_CK(cudaMemsetAsync(h->h_hi.p, 0, hb, s)); DA_CK(cudaMemsetAsync(h->h_lo.p, 0, hb, s));
      if (num_total > num_real && h->Kf > Hm) {
        h->launches++;
        DA_CK(launch_onehot_rows(h->h_hi.as<__nv_bfloat16>() + Mr * (size_t)h->Kf, h->Kf, Hm, virt_ids, num_total - num_real, s));
      }
      DA_CK(h->partial.ensure(Mr * (size_t)c.heads * 32 * sizeof(float)));
    } else {
      DA_CK(h->h_hi.ensure(Mr * Hm * b2)); DA_CK(h->h_lo.ensure(Mr * Hm * b2));
    }
    DA_CK(h->comb_hi.ensure(Mt * D * b2)); DA_CK(h->comb_lo.ensure(Mt * D * b2));
    DA_CK(h->xa_hi.ensure(Mt * hid * b2)); DA_CK(h->xa_lo.ensure(Mt * hid * b2));
    DA_CK(h->xb_hi.ensure(Mt * hid * b2)); DA_CK(h->xb_lo.ensure(Mt * hid * b2));
    DA_CK(h->r_hi.ensure(Mt * D * b2)); DA_CK(h->r_lo.ensure(Mt * D * b2));
  } else {
    DA_CK(h->hbuf.ensure(Mr * Hm * sizeof(float)));
    DA_CK(h->xa.ensure(Mt * hid * sizeof(float)));
    DA_CK(h->xb.ensure(Mt * hid * sizeof(float)));
    DA_CK(h->r.ensure(Mt * D * sizeof(float)));
  }
  if (h->use_plan && h->plan.n_tiles > 0) {
    // padded rows / channels of the operand images must read as zeros: clear once per graph; the
    // epilogues only ever write real rows and real channels (hidden and last layers own separate images)
    const size_t img_h = dense_image_elems(h->plan.n_tiles, c.heads, (c_hid + 15) / 16 * 16) * sizeof(__nv_bfloat16);
    const size_t img_l = dense_image_elems(h->plan.n_tiles, c.heads, (c_last + 15) / 16 * 16) * sizeof(__nv_bfloat16);
    DevBuf* imgs[6] = {&h->qimg, &h->kimg, &h->vimg, &h->qimg_l, &h->kimg_l, &h->vimg_l};
    for (int i = 0; i < 6; ++i) {
      DA_CK(imgs[i]->ensure(i < 3 ? img_h : img_l));
      DA_CK(cudaMemsetAsync(imgs[i]->p, 0, imgs[i]->bytes, s));
    }
    if (h->fold_cfg) {
      DA_CK(h->vimg_f.ensure(dense_image_elems(h->plan.n_tiles, c.heads, 32) * sizeof(__nv_bfloat16)));
      DA_CK(cudaMemsetAsync(h->vimg_f.p, 0, h->vimg_f.bytes, s));
    }
    DA_CK(h->dacc.ensure(Mr * (size_t)(D > hid ? D : hid) * sizeof(float)));
    DA_CK(h->dstats.ensure(Mr * (size_t)c.heads * 2 * sizeof(float)));
  }
  if (num_total > num_real) {  // virtual rows: constant embeddings appended after the real nodes (exophormer_gnn.py:169-178)
    const int nv = num_total - num_real;
    float* dst = h->combined.as<float>() + Mr * D;
    h->launches++;
    DA_CK(launch_fill_rows(dst, D, h->virt_emb.as<float>(), virt_ids, nv, D, s));
    if (umma) {
      h->launches++;
      DA_CK(launch_split_bf16(dst, D, h->comb_hi.as<__nv_bfloat16>() + Mr * D, h->comb_lo.as<__nv_bfloat16>() + Mr * D, D, nv, D, s));
    }
  }
#undef DA_CK
  h->graph_set = true;
  h->feats_set = false;
  return DA_OK;
}

int da_set_features(da_handle* h, const float* feats, void* stream) {
  if (!h) return DA_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  if (!h->weights_loaded) return h->fail(DA_ERR_MISSING, "weights not loaded");
  if (!h->graph_set) return h->fail(DA_ERR_INVALID, "graph not set");
  cudaStream_t s = (cudaStream_t)stream;
  if (!feats) { h->feats_zero = true; h->feats_set = true; return DA_OK; }
  h->feats_zero = false;
  const int Mr = h->num_real, Dv = h->cfg.feat_dim, Hm = h->cfg.mlp_hidden;
  cudaError_t ce;
  // the planner may have renumbered the nodes inside each graph: the hoisted product P is kept in internal order
  const int32_t* row_ext = h->use_plan ? h->plan.ext_of_int : nullptr;
  if (use_umma(h)) {
    const size_t b2 = sizeof(__nv_bfloat16);
    ce = h->feats_sp_hi.ensure((size_t)Mr * Dv * b2); if (ce != cudaSuccess) return h->cuda_fail(ce, "alloc");
    ce = h->feats_sp_lo.ensure((size_t)Mr * Dv * b2); if (ce != cudaSuccess) return h->cuda_fail(ce, "alloc");
    Scoped sc(h, s, TAG_OTHER);
    ce = launch_split_bf16(feats, Dv, h->feats_sp_hi.as<__nv_bfloat16>(), h->feats_sp_lo.as<__nv_bfloat16>(), Dv, Mr, Dv, s, row_ext);
    if (ce != cudaSuccess) return h->cuda_fail(ce, "split features");
  } else if (row_ext) {
    ce = h->feats_perm.ensure((size_t)Mr * Dv * sizeof(float)); if (ce != cudaSuccess) return h->cuda_fail(ce, "alloc");
    Scoped sc(h, s, TAG_OTHER);
    ce = launch_gather_rows(feats, Dv, h->feats_perm.as<float>(), Dv, Mr, Dv, row_ext, s);
    if (ce != cudaSuccess) return h->cuda_fail(ce, "gather features");
    feats = h->feats_perm.as<float>();
  }
  LinearOut o; o.f32 = h->P.as<float>(); o.ldc = Hm;
  ce = run_linear(h, h->hoist, feats, Dv, h->feats_sp_hi.as<__nv_bfloat16>(), h->feats_sp_lo.as<__nv_bfloat16>(), Dv, Mr,
                  ACT_NONE, o, TAG_HOIST_GEMM, s);
  if (ce != cudaSuccess) return h->cuda_fail(ce, "hoist gemm");
  h->feats_set = true;
  return DA_OK;
}

int da_forward(da_handle* h, const float* x, const int64_t* t, float* out, float* alpha_last, void* stream) {
  if (!h) return DA_ERR_INVALID;
  if (!x || !t || !out) return h->fail(DA_ERR_INVALID, "null x / t / out");
  cudaSetDevice(h->cfg.device);
  return forward_impl(h, x, t, 0, out, alpha_last, STEP_NONE, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int da_forward_attn(da_handle* h, const float* x, const int64_t* t, float* out, float* alpha_all, void* stream) {
  if (!h) return DA_ERR_INVALID;
  if (!x || !t || !out || !alpha_all) return h->fail(DA_ERR_INVALID, "null x / t / out / alpha");
  if (h->cfg.attn_mode != DA_ATTN_CSR) return h->fail(DA_ERR_UNSUPPORTED, "attention weights are only produced with attn_mode = DA_ATTN_CSR");
  cudaSetDevice(h->cfg.device);
  return forward_impl(h, x, t, 0, out, nullptr, STEP_NONE, nullptr, nullptr, nullptr, (cudaStream_t)stream, nullptr, alpha_all);
}

int da_ddpm_step(da_handle* h, const float* x_in, float* x_out, const da_step_coef* c, const float* noise, void* stream) {
  if (!h) return DA_ERR_INVALID;
  if (!x_in || !x_out || !c) return h->fail(DA_ERR_INVALID, "null argument");
  if (h->cfg.head_kind != DA_HEAD_2D) return h->fail(DA_ERR_UNSUPPORTED, "DDPM step is defined for the 2D head only (the 3D module binds DDIM only)");
  if (c->t_index != 0 && !noise) return h->fail(DA_ERR_INVALID, "DDPM step with t_index > 0 needs noise");
  cudaSetDevice(h->cfg.device);
  return forward_impl(h, x_in, nullptr, c->t, x_out, nullptr, STEP_DDPM, c, x_in, noise, (cudaStream_t)stream);
}

int da_ddim_step(da_handle* h, const float* x_in, float* x_out, const da_step_coef* c, const float* noise, void* stream) {
  if (!h) return DA_ERR_INVALID;
  if (!x_in || !x_out || !c) return h->fail(DA_ERR_INVALID, "null argument");
  if (c->eta > 0.f && !noise) return h->fail(DA_ERR_INVALID, "DDIM step with eta > 0 needs noise");
  if (c->eta > 0.f && h->cfg.head_kind == DA_HEAD_SE3) return h->fail(DA_ERR_UNSUPPORTED, "SE3 sampler is eta = 0 only");
  cudaSetDevice(h->cfg.device);
  return forward_impl(h, x_in, nullptr, c->t, x_out, nullptr, STEP_DDIM, c, x_in, noise, (cudaStream_t)stream);
}

static int check_sched(da_handle* h, const da_schedule* sc) {
  if (!sc || !sc->betas || !sc->alphas_cumprod || !sc->sqrt_one_minus_alphas_cumprod || !sc->sqrt_recip_alphas || !sc->posterior_variance ||
      !sc->sqrt_recip_alphas_cumprod || !sc->sqrt_recipm1_alphas_cumprod)
    return h->fail(DA_ERR_INVALID, "null schedule table");
  if (sc->steps != h->cfg.steps || sc->inference_ratio <= 0) return h->fail(DA_ERR_INVALID, "schedule length / inference_ratio mismatch");
  return DA_OK;
}

int da_ddpm_step_t(da_handle* h, const float* x_in, float* x_out, const int64_t* t, int32_t t_index, const da_schedule* sched,
                   const float* noise, void* stream) {
  if (!h) return DA_ERR_INVALID;
  if (!x_in || !x_out || !t) return h->fail(DA_ERR_INVALID, "null argument");
  if (int rc = check_sched(h, sched)) return rc;
  if (h->cfg.head_kind != DA_HEAD_2D) return h->fail(DA_ERR_UNSUPPORTED, "DDPM step is defined for the 2D head only (the 3D module binds DDIM only)");
  if (t_index != 0 && !noise) return h->fail(DA_ERR_INVALID, "DDPM step with t_index > 0 needs noise");
  cudaSetDevice(h->cfg.device);
  da_step_coef c{};
  c.t_index = t_index; c.pred = DA_PRED_EPSILON; c.eta = 1.f;
  return forward_impl(h, x_in, t, 0, x_out, nullptr, STEP_DDPM, &c, x_in, noise, (cudaStream_t)stream, sched);
}

int da_ddim_step_t(da_handle* h, const float* x_in, float* x_out, const int64_t* t, int32_t pred, float eta, const da_schedule* sched,
                   const float* noise, void* stream) {
  if (!h) return DA_ERR_INVALID;
  if (!x_in || !x_out || !t) return h->fail(DA_ERR_INVALID, "null argument");
  if (int rc = check_sched(h, sched)) return rc;
  if (eta > 0.f && !noise) return h->fail(DA_ERR_INVALID, "DDIM step with eta > 0 needs noise");
  if (eta > 0.f && h->cfg.head_kind == DA_HEAD_SE3) return h->fail(DA_ERR_UNSUPPORTED, "SE3 sampler is eta = 0 only");
  if (pred != DA_PRED_START_X && pred != DA_PRED_EPSILON) return h->fail(DA_ERR_INVALID, "unknown pred");
  cudaSetDevice(h->cfg.device);
  da_step_coef c{};
  c.pred = pred; c.eta = eta;
  return forward_impl(h, x_in, t, 0, x_out, nullptr, STEP_DDIM, &c, x_in, noise, (cudaStream_t)stream, sched);
}

int da_ddim_update(da_handle* h, const float* x_in, const float* model_out, float* x_out, const da_step_coef* c,
                   const float* noise, void* stream) {
  if (!h) return DA_ERR_INVALID;
  if (!x_in || !model_out || !x_out || !c) return h->fail(DA_ERR_INVALID, "null argument");
  if (!h->graph_set) return h->fail(DA_ERR_INVALID, "graph not set");
  if (c->eta > 0.f && !noise) return h->fail(DA_ERR_INVALID, "eta > 0 needs noise");
  cudaSetDevice(h->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  Scoped sc(h, s, TAG_OTHER);
  cudaError_t ce = launch_sampler_update(x_in, model_out, x_out, h->num_real, h->cfg.out_channels, h->cfg.head_kind,
                                         STEP_DDIM, *c, noise, s);
  if (ce != cudaSuccess) return h->cuda_fail(ce, "sampler update");
  return DA_OK;
}

size_t da_workspace_bytes(const da_handle* h) { return h ? h->workspace_bytes() : 0; }
int64_t da_launch_count(const da_handle* h) { return h ? h->launches : 0; }

int da_graph_stats(const da_handle* h, int64_t* n_dense_edges, int64_t* n_csr_edges, int32_t* n_dense_graphs) {
  if (!h || !h->graph_set) return DA_ERR_INVALID;
  if (n_dense_edges) *n_dense_edges = h->use_plan ? h->plan.n_dense_edges : 0;
  if (n_csr_edges) *n_csr_edges = h->use_plan ? h->plan.residual.E : h->csr.E;
  if (n_dense_graphs) *n_dense_graphs = h->use_plan ? h->plan.n_dense_graphs : 0;
  return DA_OK;
}

int da_graph_plan_info(const da_handle* h, int64_t* out, int32_t n) {
  if (!h || !h->graph_set || !out || n < 8) return DA_ERR_INVALID;
  const DensePlan& p = h->plan;
  const bool on = h->use_plan;
  out[0] = on ? p.n_tiles : 0;
  out[1] = on ? p.n_blocks_total : 0;      // (tile, 64-source block) pairs of the dense tiles
  out[2] = on ? p.n_blocks_listed : 0;     // ... that hold at least one edge and are visited
  out[3] = on ? p.n_blocks_full : 0;       // ... of which every bit is set
  out[4] = on ? p.n_reordered_graphs : 0;  // graphs renumbered by the ring walk
  out[5] = on ? p.n_extra : 0;             // promoted extra sources
  out[6] = on ? p.n_fused : 0;             // rows finalised inside the dense kernel
  out[7] = on ? p.n_csr_rows : 0;          // rows served by the CSR kernels
  if (n >= 10) {
    out[8] = (on && p.n_tiles > 0 && p.real_rows_clean) ? 1 : 0;                       // no real row needs a CSR kernel
    out[9] = (h->fold_ready && on && p.n_tiles > 0 && p.real_rows_clean) ? 1 : 0;      // steps take the folded path (fold.cu)
  }
  if (n >= 11) out[10] = h->persist_launches;   // hidden-layer launches of the persistent kernel so far (attn_hidden.cu)
  return DA_OK;
}

int da_debug_trace(da_handle* h, long long* device_buf, int layer) {
  if (!h) return DA_ERR_INVALID;
  h->dbg_trace = device_buf; h->dbg_layer = layer;
  return DA_OK;
}

int da_set_profiling(da_handle* h, int32_t enable) {
  if (!h) return DA_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  if (!enable) drain_profile(h);
  h->profiling = enable != 0;
  return DA_OK;
}

int da_get_profile(da_handle* h, double* ms_out, int64_t* launches_out, int32_t n_classes, int32_t reset) {
  if (!h) return DA_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  drain_profile(h);
  for (int i = 0; i < n_classes && i < TAG_COUNT; ++i) {
    if (ms_out) ms_out[i] = h->prof_ms[i];
    if (launches_out) launches_out[i] = h->prof_n[i];
  }
  if (reset) for (int i = 0; i < TAG_COUNT; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
  return TAG_COUNT;
}

const char* da_profile_tag_name(int32_t i) { return (i >= 0 && i < TAG_COUNT) ? kTagNames[i] : ""; }

// ---- stand-alone operators ------------------------------------------------------------------
size_t da_op_linear_workspace_bytes(int32_t mode, int32_t M, int32_t N, int32_t K) {
  if (mode != DA_GEMM_BF16X3_UMMA) return 0;
  const size_t align = 256;
  auto up = [&](size_t b) { return (b + align - 1) / align * align; };
  return 2 * up((size_t)M * K * sizeof(__nv_bfloat16)) + 2 * up((size_t)N * K * sizeof(__nv_bfloat16));
}

// y = act(a @ w^T + bias) with caller-provided workspace (no allocation, no synchronisation): the form the
// training-side autograd functions call.  workspace must hold da_op_linear_workspace_bytes(...) bytes.
int da_op_linear_ws(int32_t mode, const float* a, const float* w, const float* bias, float* y, int32_t M, int32_t N,
                    int32_t K, int32_t act, void* workspace, size_t workspace_bytes, void* stream) {
  if (!a || !w || !y || M <= 0 || N <= 0 || K <= 0) return DA_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  LinearOut o; o.f32 = y; o.ldc = N;
  cudaError_t ce;
  if (mode == DA_GEMM_FP32_SIMT) {
    ce = launch_linear_simt(a, K, w, K, bias, o, M, N, K, act, s);
    return ce == cudaSuccess ? DA_OK : (ce == cudaErrorInvalidValue ? DA_ERR_UNSUPPORTED : DA_ERR_CUDA);
  }
  if (mode != DA_GEMM_BF16X3_UMMA) return DA_ERR_INVALID;
  if (K % 64 || N % 32) return DA_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < da_op_linear_workspace_bytes(mode, M, N, K)) return DA_ERR_INVALID;
  const size_t align = 256;
  auto up = [&](size_t b) { return (b + align - 1) / align * align; };
  uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
  __nv_bfloat16* ahi = reinterpret_cast<__nv_bfloat16*>(p); p += up((size_t)M * K * 2);
  __nv_bfloat16* alo = reinterpret_cast<__nv_bfloat16*>(p); p += up((size_t)M * K * 2);
  __nv_bfloat16* whi = reinterpret_cast<__nv_bfloat16*>(p); p += up((size_t)N * K * 2);
  __nv_bfloat16* wlo = reinterpret_cast<__nv_bfloat16*>(p);
  ce = launch_split_bf16(a, K, ahi, alo, K, M, K, s);
  if (ce == cudaSuccess) ce = launch_split_bf16(w, K, whi, wlo, K, N, K, s);
  if (ce == cudaSuccess) ce = launch_linear_umma(ahi, alo, K, whi, wlo, K, bias, o, M, N, K, act, s);
  if (ce != cudaSuccess) return (ce == cudaErrorInvalidValue) ? DA_ERR_UNSUPPORTED : DA_ERR_CUDA;
  return DA_OK;
}

int da_op_linear(int32_t mode, const float* a, const float* w, const float* bias, float* y, int32_t M, int32_t N,
                 int32_t K, int32_t act, void* stream) {
  const size_t bytes = da_op_linear_workspace_bytes(mode, M, N, K);
  void* ws = nullptr;
  if (bytes && cudaMalloc(&ws, bytes) != cudaSuccess) return DA_ERR_CUDA;
  int rc = da_op_linear_ws(mode, a, w, bias, y, M, N, K, act, ws, bytes, stream);
  if (bytes) {
    if (rc == DA_OK && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = DA_ERR_CUDA;
    cudaFree(ws);
  }
  return rc;
}

int da_op_graph_attention(const float* qkvs, const int64_t* edge_src, const int64_t* edge_dst, int64_t E, int32_t n,
                          int32_t H, int32_t C, float* y, float* alpha, void* stream) {
  if (!qkvs || !y || n <= 0 || H <= 0 || C <= 0 || E < 0) return DA_ERR_INVALID;
  if ((C + 31) / 32 > 13) return DA_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  CsrGraph g;
  const char* why = "";
  cudaError_t ce = build_csr(edge_src, edge_dst, E, n, &g, s, &why);
  if (ce != cudaSuccess) return ce == cudaErrorInvalidValue ? DA_ERR_INVALID : DA_ERR_CUDA;
  float *scores = nullptr, *stats = nullptr;
  int rc = DA_OK;
  if (alpha) {
    if (cudaMalloc(&scores, sizeof(float) * (size_t)(E > 0 ? E : 1) * H) != cudaSuccess ||
        cudaMalloc(&stats, sizeof(float) * (size_t)n * H * 2) != cudaSuccess) rc = DA_ERR_CUDA;
  }
  if (rc == DA_OK) {
    AttnCsrArgs a{};
    a.qkvs = qkvs; a.ld = 4 * H * C; a.rowptr = g.rowptr; a.col = g.col; a.n_targets = n; a.H = H; a.C = C;
    a.act = ACT_NONE; a.out.f32 = y; a.out.ldc = H * C; a.scores = scores; a.stats = stats;
    ce = launch_attn_csr(a, s);
    if (ce == cudaSuccess && alpha) ce = launch_alpha_normalize(scores, stats, g.rowptr, g.eid, n, H, alpha, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) rc = DA_ERR_CUDA;
  }
  cudaFree(scores); cudaFree(stats);
  free_csr(&g);
  return rc;
}

int da_greedy_cost_assignment(const float* pos1, int32_t ld1, const float* pos2, int32_t ld2, const int32_t* graph_ptr,
                              int32_t n_graphs, int32_t max_nodes, int64_t* out, void* stream) {
  if (!pos1 || !pos2 || !graph_ptr || !out || n_graphs < 0 || max_nodes <= 0 || ld1 < 2 || ld2 < 2) return DA_ERR_INVALID;
  cudaError_t ce = launch_greedy_assign(pos1, ld1, pos2, ld2, graph_ptr, n_graphs, max_nodes, out, (cudaStream_t)stream);
  if (ce == cudaErrorInvalidValue) return DA_ERR_UNSUPPORTED;
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_expander_edge_index(const int32_t* perm, int32_t n, int32_t degree, int32_t n_graphs, int64_t* edge_src,
                           int64_t* edge_dst, void* stream) {
  if (!perm || !edge_src || !edge_dst || n <= 0 || degree < 0 || degree >= n || n_graphs <= 0) return DA_ERR_INVALID;
  if (((long long)n * degree) % 2) return DA_ERR_INVALID;   // "nodes * degree must be even" (puzzle_dataset.py:128)
  cudaError_t ce = launch_expander_edges(perm, n, degree, n_graphs, edge_src, edge_dst, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

// ---- training-side graph object and operators (scope row N1) ------------------------------------------------
struct da_graph {
  CsrGraph by_target, by_source;
  int n = 0;
  int64_t E = 0;
  // optional dense-tile plan (da_graph_set_batch): when every edge of the batch lands in a bitmap -- the dense
  // puzzle graphs of the training configs -- the forward runs on the tensor-core kernel and only the backward
  // walks the edge lists
  DensePlan plan;
  bool dense_ok = false;
  DevBuf qimg, kimg, vimg, acc;
  int img_h = 0, img_cpad = 0;
  // per-graph descriptors {node0, n, bm_words, pad, bm_off} for the shared-memory backward
  DevBuf graphs;
  int n_graphs = 0, n_max = 0, bm_words_max = 0;
};

int da_graph_create(da_graph** out, const int64_t* edge_src, const int64_t* edge_dst, int64_t E, int32_t n, void* stream) {
  if (!out || n <= 0 || E < 0 || (E > 0 && (!edge_src || !edge_dst))) return DA_ERR_INVALID;
  *out = nullptr;
  da_graph* g = new da_graph();
  const char* why = "";
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t ce = build_csr(edge_src, edge_dst, E, n, &g->by_target, s, &why);
  if (ce == cudaSuccess) ce = build_csr(edge_dst, edge_src, E, n, &g->by_source, s, &why);   // CSC = CSR of the reversed edges
  if (ce != cudaSuccess) { free_csr(&g->by_target); free_csr(&g->by_source); delete g; return ce == cudaErrorInvalidValue ? DA_ERR_INVALID : DA_ERR_CUDA; }
  g->n = n; g->E = E;
  *out = g;
  return DA_OK;
}

int da_graph_set_batch(da_graph* g, const int64_t* edge_src, const int64_t* edge_dst, const int64_t* batch, void* stream) {
  if (!g || !batch || (g->E > 0 && (!edge_src || !edge_dst))) return DA_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const char* why = "";
  g->dense_ok = false;
  cudaError_t ce = build_dense_plan(edge_src, edge_dst, g->E, batch, g->n, g->n, &g->plan, s, &why);
  if (ce != cudaSuccess) return ce == cudaErrorInvalidValue ? DA_ERR_INVALID : DA_ERR_CUDA;
  // usable when the bitmaps hold the WHOLE multiset and every node sits in a tile: then the dense kernel's
  // (m, l) are the final softmax statistics the backward needs
  bool all_in_tiles = g->plan.n_tiles > 0 && g->plan.residual.E == 0;
  if (all_in_tiles) {
    std::vector<int32_t> slot((size_t)g->n);
    if (copy_sync(slot.data(), g->plan.node_slot, sizeof(int32_t) * slot.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess)
      return DA_ERR_CUDA;
    for (int32_t v : slot) if (v < 0) { all_in_tiles = false; break; }
  }
  g->dense_ok = all_in_tiles;
  g->n_graphs = 0;
  if (all_in_tiles) {   // one descriptor per graph, from the first tile of each
    std::vector<TileInfo> tiles((size_t)g->plan.n_tiles);
    if (copy_sync(tiles.data(), g->plan.tiles, sizeof(TileInfo) * tiles.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess)
      return DA_ERR_CUDA;
    struct GD { int32_t node0, n, bm_words, pad; int64_t bm_off; };
    std::vector<GD> gds;
    g->n_max = 0; g->bm_words_max = 0;
    for (const TileInfo& t : tiles) {
      if (t.row0 != 0) continue;
      gds.push_back(GD{t.node0, t.gn, t.bm_words, 0, t.bm_off});
      if (t.gn > g->n_max) g->n_max = t.gn;
      if (t.bm_words > g->bm_words_max) g->bm_words_max = t.bm_words;
    }
    if (g->graphs.ensure(sizeof(GD) * gds.size()) != cudaSuccess) return DA_ERR_CUDA;
    if (copy_sync(g->graphs.p, gds.data(), sizeof(GD) * gds.size(), cudaMemcpyHostToDevice, s) != cudaSuccess) return DA_ERR_CUDA;
    g->n_graphs = (int)gds.size();
  }
  return DA_OK;
}

void da_graph_destroy(da_graph* g) {
  if (!g) return;
  free_csr(&g->by_target); free_csr(&g->by_source);
  free_plan(&g->plan);
  g->qimg.release(); g->kimg.release(); g->vimg.release(); g->acc.release(); g->graphs.release();
  delete g;
}

int da_op_graph_attention_fwd(const da_graph* gc, const float* qkvs, int32_t H, int32_t C, float* y, float* stats, void* stream) {
  da_graph* g = const_cast<da_graph*>(gc);   // (lazily sized scratch images live in the graph object)
  if (!g || !qkvs || !y || !stats || H <= 0 || C <= 0) return DA_ERR_INVALID;
  if ((C + 31) / 32 > 13) return DA_ERR_UNSUPPORTED;
  const int Cpad = (C + 15) / 16 * 16;
  if (g->dense_ok && C % 8 == 0 && Cpad <= 144 && attn_csr_rows_supported(H, C)) {
    // tensor-core forward: repack -> bitmap-masked dense attention (un-fused: it leaves (acc, m, l)) -> normalise + skip.
    // The statistics land directly in `stats` in the (max, sum) convention of the CSR kernel.
    cudaStream_t s = (cudaStream_t)stream;
    const size_t img = dense_image_elems(g->plan.n_tiles, H, Cpad) * sizeof(__nv_bfloat16);
    if (g->img_h != H || g->img_cpad != Cpad || g->qimg.bytes < img) {
      if (g->qimg.ensure(img) != cudaSuccess || g->kimg.ensure(img) != cudaSuccess || g->vimg.ensure(img) != cudaSuccess)
        return DA_ERR_CUDA;
      cudaMemsetAsync(g->qimg.p, 0, g->qimg.bytes, s); cudaMemsetAsync(g->kimg.p, 0, g->kimg.bytes, s);
      cudaMemsetAsync(g->vimg.p, 0, g->vimg.bytes, s);
      g->img_h = H; g->img_cpad = Cpad;
    }
    if (g->acc.ensure((size_t)g->n * H * C * sizeof(float)) != cudaSuccess) return DA_ERR_CUDA;
    PackArgs pa{qkvs, 4 * H * C, g->plan.node_slot, g->n, H, C, Cpad, g->qimg.as<__nv_bfloat16>(), g->kimg.as<__nv_bfloat16>(),
                g->vimg.as<__nv_bfloat16>()};
    cudaError_t ce = launch_pack_images(pa, s);
    if (ce == cudaSuccess) ce = launch_gather_extra(pa, g->plan.x_src, g->plan.x_slot, g->plan.n_extra, s);
    if (ce == cudaSuccess) {
      AttnDenseArgs da_{pa.qimg, pa.kimg, pa.vimg, g->plan.tiles, g->plan.n_tiles, g->plan.bitmap, H, C, Cpad, g->acc.as<float>(),
                        stats, nullptr};
      da_.blk_list = g->plan.blk_list;
      ce = launch_attn_dense(da_, s);
    }
    if (ce == cudaSuccess) {
      AttnCsrArgs a{};
      a.qkvs = qkvs; a.ld = 4 * H * C; a.rowptr = g->plan.residual.rowptr; a.col = g->plan.residual.col;
      a.weight = g->plan.residual.weight; a.n_targets = g->n; a.H = H; a.C = C; a.act = ACT_NONE;
      a.out.f32 = y; a.out.ldc = H * C;
      a.init_acc = g->acc.as<float>(); a.init_stats = stats; a.init_slot = g->plan.node_slot;
      ce = launch_attn_csr_rows(a, s);
    }
    return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
  }
  AttnCsrArgs a{};
  a.qkvs = qkvs; a.ld = 4 * H * C; a.rowptr = g->by_target.rowptr; a.col = g->by_target.col; a.n_targets = g->n; a.H = H; a.C = C;
  a.act = ACT_NONE; a.out.f32 = y; a.out.ldc = H * C; a.stats = stats;
  cudaError_t ce = launch_attn_csr(a, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_graph_attention_bwd(const da_graph* g, const float* qkvs, const float* stats, const float* dy, int32_t H, int32_t C,
                              float* dqkvs, float* delta_ws, void* stream) {
  if (!g || !qkvs || !stats || !dy || !dqkvs || !delta_ws || H <= 0 || C <= 0) return DA_ERR_INVALID;
  if ((C + 31) / 32 > 13) return DA_ERR_UNSUPPORTED;
  cudaError_t ce;
  if (g->dense_ok && g->n_graphs > 0 && attn_backward_dense_fits(g->n_max, C, g->bm_words_max) && !getenv("DA_NO_DENSE_BWD"))
    ce = launch_attn_backward_dense(qkvs, dy, g->graphs.p, g->n_graphs, g->n_max, g->bm_words_max, g->plan.bitmap, stats, g->n, H,
                                    C, dqkvs, delta_ws, (cudaStream_t)stream);
  else
    ce = launch_attn_backward(qkvs, dy, g->by_target, g->by_source, stats, g->n, H, C, dqkvs, delta_ws, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_linear_wgrad(const float* dy, const float* x, float* dw, float* db, int32_t M, int32_t N, int32_t K, void* stream) {
  if (!dy || !x || !dw || M <= 0 || N <= 0 || K <= 0) return DA_ERR_INVALID;
  cudaError_t ce = launch_linear_wgrad(dy, x, dw, db, M, N, K, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_adafactor_step(const da_adafactor_param* params, int32_t n, float eps1, float eps2, float clip_threshold,
                      float weight_decay, void* stream) {
  if (n < 0 || (n > 0 && !params) || clip_threshold <= 0.f) return DA_ERR_INVALID;
  cudaError_t ce = launch_adafactor(params, n, eps1, eps2, clip_threshold, weight_decay, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_segment_max(const float* x, int32_t ld, const int32_t* seg_ptr, int32_t n_seg, int32_t cols, float* out, void* stream) {
  if (!x || !seg_ptr || !out || n_seg < 0 || cols <= 0 || ld < cols) return DA_ERR_INVALID;
  cudaError_t ce = launch_segment_max(x, ld, seg_ptr, n_seg, cols, out, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

// ---- scope row N4: convolution pieces of the EfficientNet-B0 patch encoder (NHWC fp32) --------------------------------
int da_op_conv2d_nhwc(const float* x, const float* w, const float* bias, float* y, int32_t N, int32_t H, int32_t W, int32_t Cin,
                      int32_t Cout, int32_t k, int32_t stride, int32_t pad, int32_t act, void* stream) {
  if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || k <= 0 || stride <= 0 || pad < 0) return DA_ERR_INVALID;
  cudaError_t ce = launch_conv2d_nhwc(x, w, bias, y, N, H, W, Cin, Cout, k, stride, pad, act, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_dwconv2d_nhwc(const float* x, const float* w, const float* bias, float* y, int32_t N, int32_t H, int32_t W, int32_t C,
                        int32_t k, int32_t stride, int32_t pad, int32_t act, void* stream) {
  if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0 || k <= 0 || stride <= 0 || pad < 0) return DA_ERR_INVALID;
  if (C % 4) return DA_ERR_UNSUPPORTED;
  cudaError_t ce = launch_dwconv2d_nhwc(x, w, bias, y, N, H, W, C, k, stride, pad, act, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_spatial_mean(const float* x, float* y, int32_t ldy, int32_t N, int32_t HW, int32_t C, void* stream) {
  if (!x || !y || N <= 0 || HW <= 0 || C <= 0 || ldy < C) return DA_ERR_INVALID;
  cudaError_t ce = launch_spatial_mean(x, y, ldy, N, HW, C, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_channel_scale(float* x, const float* gate, int32_t ldg, int32_t N, int32_t HW, int32_t C, void* stream) {
  if (!x || !gate || N <= 0 || HW <= 0 || C <= 0 || ldg < C) return DA_ERR_INVALID;
  if (C % 4 || ldg % 4) return DA_ERR_UNSUPPORTED;
  cudaError_t ce = launch_channel_scale(x, gate, ldg, N, HW, C, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_normalize_to_nhwc(const float* x, const float* mean, const float* stdv, float* y, int32_t N, int32_t C, int32_t HW, void* stream) {
  if (!x || !mean || !stdv || !y || N <= 0 || C <= 0 || HW <= 0) return DA_ERR_INVALID;
  cudaError_t ce = launch_normalize_to_nhwc(x, mean, stdv, y, N, C, HW, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_add_inplace(float* y, const float* x, int64_t n, void* stream) {
  if (!x || !y || n < 0) return DA_ERR_INVALID;
  if (n % 4) return DA_ERR_UNSUPPORTED;
  cudaError_t ce = launch_add_inplace(y, x, (size_t)n, (cudaStream_t)stream);
  return ce == cudaSuccess ? DA_OK : DA_ERR_CUDA;
}

int da_op_graph_attention_dense(const float* qkvs, const int64_t* edge_src, const int64_t* edge_dst, int64_t E,
                                const int64_t* batch, int32_t n, int32_t H, int32_t C, float* y, int64_t* n_dense_edges,
                                void* stream) {
  if (!qkvs || !y || !batch || n <= 0 || H <= 0 || C <= 0 || E < 0) return DA_ERR_INVALID;
  const int Cpad = (C + 15) / 16 * 16;
  if (C % 8 || Cpad > 144) return DA_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  DensePlan plan;
  const char* why = "";
  cudaError_t ce = build_dense_plan(edge_src, edge_dst, E, batch, n, n, &plan, s, &why);
  if (ce != cudaSuccess) return ce == cudaErrorInvalidValue ? DA_ERR_INVALID : DA_ERR_CUDA;
  if (n_dense_edges) *n_dense_edges = plan.n_dense_edges;
  __nv_bfloat16 *qi = nullptr, *ki = nullptr, *vi = nullptr;
  float *acc = nullptr, *st = nullptr;
  int rc = DA_OK;
  const bool fuse = plan.n_tiles > 0 && plan.n_fused > 0 && attn_csr_rows_supported(H, C) && attn_dense_can_fuse(C) &&
                    !(getenv("DA_NO_FUSE") && getenv("DA_NO_FUSE")[0] == '1');
  const size_t img = dense_image_elems(plan.n_tiles > 0 ? plan.n_tiles : 1, H, Cpad) * sizeof(__nv_bfloat16);
  if (cudaMalloc(&qi, img) != cudaSuccess || cudaMalloc(&ki, img) != cudaSuccess || cudaMalloc(&vi, img) != cudaSuccess ||
      cudaMalloc(&acc, sizeof(float) * (size_t)n * H * C) != cudaSuccess || cudaMalloc(&st, sizeof(float) * (size_t)n * H * 2) != cudaSuccess) {
    rc = DA_ERR_CUDA;
  } else {
    cudaMemsetAsync(qi, 0, img, s); cudaMemsetAsync(ki, 0, img, s); cudaMemsetAsync(vi, 0, img, s);
    ce = cudaSuccess;
    if (plan.n_tiles > 0) {
      PackArgs pa{qkvs, 4 * H * C, plan.node_slot, n, H, C, Cpad, qi, ki, vi};
      ce = launch_pack_images(pa, s);
      if (ce == cudaSuccess) ce = launch_gather_extra(pa, plan.x_src, plan.x_slot, plan.n_extra, s);
      if (ce == cudaSuccess) {
        AttnDenseArgs da_{qi, ki, vi, plan.tiles, plan.n_tiles, plan.bitmap, H, C, Cpad, acc, st, nullptr};
        da_.blk_list = plan.blk_list;
        if (fuse) {
          da_.row_fused = plan.row_fused; da_.qkvs = qkvs; da_.ld = 4 * H * C; da_.n_rows = n;
          da_.rowptr = plan.residual.rowptr; da_.col = plan.residual.col; da_.weight = plan.residual.weight;
          da_.act = ACT_NONE; da_.out.f32 = y; da_.out.ldc = H * C;
        }
        ce = launch_attn_dense(da_, s);
      }
    }
    if (ce == cudaSuccess) {
      AttnCsrArgs a{};
      a.qkvs = qkvs; a.ld = 4 * H * C; a.rowptr = plan.residual.rowptr; a.col = plan.residual.col; a.weight = plan.residual.weight; a.n_targets = n;
      a.H = H; a.C = C; a.act = ACT_NONE; a.out.f32 = y; a.out.ldc = H * C;
      if (plan.n_tiles > 0) { a.init_acc = acc; a.init_stats = st; a.init_slot = plan.node_slot; }
      if (attn_csr_rows_supported(H, C)) {
        AttnCsrArgs hv = a, lt = a;
        hv.node_list = plan.heavy; hv.n_targets = plan.n_heavy;
        if (fuse) { lt.node_list = plan.light_nf; lt.n_targets = plan.n_light_nf; }
        else { lt.node_list = plan.light; lt.n_targets = plan.n_light; }
        ce = launch_attn_csr_heavy(hv, s);
        if (ce == cudaSuccess) ce = launch_attn_csr_rows(lt, s);
      } else {
        ce = launch_attn_csr(a, s);
      }
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) rc = DA_ERR_CUDA;
  }
  cudaFree(qi); cudaFree(ki); cudaFree(vi); cudaFree(acc); cudaFree(st);
  free_plan(&plan);
  return rc;
}

}  // extern "C"
