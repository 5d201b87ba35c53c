// Hidden graph-transformer layers (32-channel heads) on the dense tiles: PERSISTENT form of attn_dense_kernel<32, 4>.
//
// attn_dense.cu launches one CTA per (128-target tile, head); at the benchmarked size that is 2 048 CTAs of ~20 us, two
// per SM, and the per-CTA timeline (profiles/r2_dense_hidden_cta_timeline.txt) shows what that costs: barrier / TMEM
// set-up, the Q tile parked before the first score exists, an epilogue that drains the tensor pipe, CTA turn-over -- a
// mean of 1.5 resident CTAs per SM instead of 2.  Here ONE CTA per SM runs TWO independent streams of (tile, head)
// items for the whole layer; a stream is exactly the warp set of one attn_dense CTA:
//     producer warp   K / V blocks through cp.async.bulk rings whose counters run across items (the next item's
//                     blocks are in flight while this one computes)
//     MMA warp        S = Q K^T (Q from TMEM, double-buffered per item) and O += P V, 3-pass split-bf16; the next
//                     item's first S is issued as soon as its Q is parked, under this item's epilogue
//     4 softmax warps thread = target row; same one-pass masked online softmax as attn_dense.cu (P written back over
//                     S in TMEM), then -- while the MMA warp already works on the next item -- the layer epilogue:
//                     O / l + skip -> activation -> bf16 hi / lo planes, TMA tensor store
// While one stream is in a latency-bound phase (Q park, epilogue, a barrier hand-off) the other one owns the issue
// slots, and nothing is ever re-initialised: barriers keep running phases, TMEM is allocated once (S/P 2 x 64, O 2 x 32,
// Q 2 x 32 columns per stream = all 512), skip rows and output planes have their own staging buffers.
//
// The score loop keeps no running maximum (template parameter NOMAX, the default): see the kernel's comment.  Launched
// with programmatic stream serialization (common.cuh: launch_pdl / pdl_wait), so barrier set-up and the TMEM allocation
// run under the previous kernel's tail.  What bounds it now, with the measurements: DESIGN.md section 4.
//
// Preconditions (checked by the host, api.cu): every valid tile row is finalised here (planner: no residual in-edge on
// a real row, DensePlan::real_rows_clean), tcgen05 GEMM mode (split-bf16 output planes), C = Cpad = 32.
// TransformerConv semantics: SURVEY.md section 2.3c; reference call sites Transformer_GNN.py:33-44,
// exophormer_gnn.py:203-213.
#include <cstring>
#include <cstdlib>

#include "attn_tc.cuh"

namespace da {
namespace {

constexpr int HC_ = 32;            // head dim (= padded head dim)
constexpr int HST = 4;             // K / V ring depth per stream
constexpr int NTH = 384;           // 2 streams x (producer + MMA + 4 softmax warps)
constexpr uint32_t KV_PLANE = TS * HC_ * 2;          // bytes of one bf16 plane of a K / V block (4 KB)
constexpr uint32_t KV_STAGE = 2 * KV_PLANE;          // hi + lo
constexpr uint32_t SKIP_BYTES = TM * HC_ * 4;        // 128 rows x 32 floats (4 TMA boxes of 64 rows x 16 floats)
constexpr uint32_t OUT_PLANE = TM * HC_ * 2;         // 128 rows x 32 bf16
constexpr uint32_t STREAM_BYTES = 2 * HST * KV_STAGE + SKIP_BYTES + 2 * OUT_PLANE;   // 96 KB

struct HidStream {
  uint64_t k_full[HST], k_empty[HST], v_full[HST], v_empty[HST];
  uint64_t s_full[2];    // MMA -> softmax: S of running block g is in TMEM buffer g & 1
  uint64_t p_full[2];    // softmax -> MMA: P (bf16 hi | lo) has replaced it
  uint64_t pv_done[2];   // MMA -> both: P V of the block retired (buffer free, O holds the item's blocks so far)
  uint64_t q_full[2];    // softmax -> MMA: Q tile of item it is in TMEM Q buffer it & 1
  uint64_t q_free[2];    // MMA -> softmax: every S of the item that used the buffer has retired
  uint64_t o_full;       // MMA -> softmax: every P V of the item has retired (one phase per item)
  uint64_t skip_full;    // TMA -> softmax: the item's skip rows are staged (one phase per item)
};
struct HidSmem {
  HidStream st[2];
  uint32_t tmem_base;
};

// DBG: clock64 trace of CTA 0 (development aid, scripts/trace_hidden.py): a.dbg[((stream * 8 + item) * 16 + block) * 8 + k],
// k = 0..3 softmax (wait start, S ready, S in registers, P published), 4..6 MMA (S_{j+1} issued, p_full seen, P V issued);
// per item at a.dbg[2048 + (stream * 8 + item) * 4 + k]: Q park start / done, O complete, epilogue done
// NOMAX: no running maximum in the score loop.  The reference point of a row is the masked maximum of the first block in
// which the row has an edge and only moves when a block's row sum leaves [0, 2^60) (then the block's maximum is taken from
// the scores still in TMEM and the block is redone): any reference point cancels in O / l, P keeps its relative
// precision as split bf16 at any magnitude, and l >= 1 keeps the reference's 1e-16 in the denominator negligible.
template <bool DBG, bool NOMAX>
__global__ void __launch_bounds__(NTH, 1)
attn_hidden_persist_kernel(const __grid_constant__ CUtensorMap map_skip, const __grid_constant__ CUtensorMap map_ohi,
                           const __grid_constant__ CUtensorMap map_olo, AttnDenseArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  HidSmem* sh = reinterpret_cast<HidSmem*>(smem + 2 * STREAM_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warps 0 / 1: producer + MMA of stream 0, warps 2 / 3: of stream 1, warps 4-7: softmax of stream 0, 8-11: of stream 1
  const int sidx = warp < 4 ? (warp >> 1) : ((warp - 4) >> 2);
  const int role = warp < 4 ? (warp & 1) : 2;   // 0 producer, 1 MMA, 2 softmax
  HidStream* bs = &sh->st[sidx];
  uint8_t* sbase = smem + (size_t)sidx * STREAM_BYTES;
  uint8_t* k_sm = sbase;
  uint8_t* v_sm = k_sm + HST * KV_STAGE;
  uint8_t* skip_sm = v_sm + HST * KV_STAGE;
  uint8_t* ohi_sm = skip_sm + SKIP_BYTES;
  uint8_t* olo_sm = ohi_sm + OUT_PLANE;

  const int n_items = a.n_tiles * a.H;
  const int w0 = 2 * (int)blockIdx.x + sidx, wstep = 2 * (int)gridDim.x;   // this stream's items: w0, w0 + wstep, ...

  if (threadIdx.x == 0) {
    for (int s_ = 0; s_ < 2; ++s_) {
      HidStream* b = &sh->st[s_];
      for (int i = 0; i < HST; ++i) {
        mbar_init(&b->k_full[i], 1); mbar_init(&b->k_empty[i], 1);
        mbar_init(&b->v_full[i], 1); mbar_init(&b->v_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&b->s_full[i], 1); mbar_init(&b->p_full[i], 4); mbar_init(&b->pv_done[i], 1);
        mbar_init(&b->q_full[i], 4); mbar_init(&b->q_free[i], 1);
      }
      mbar_init(&b->o_full, 1); mbar_init(&b->skip_full, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  const uint32_t tmem_s = tmem_base + (uint32_t)(sidx * 256);   // 2 x TS columns
  const uint32_t tmem_o = tmem_s + 2 * TS;                       // 2 x 32 columns: O_a = P_hi V_hi + P_lo V_hi | O_b = P_hi V_lo
  const uint32_t tmem_q0 = tmem_o + 2 * HC_;                     // 2 x 32 columns: Q as packed bf16 pairs, hi plane then lo plane
  pdl_wait();   // barriers and TMEM are set up under the previous kernel's tail; the operand images are read from here on

  if (role == 0) {  // ===== bulk-copy producer: running block counter kc over all items of the stream =====
    int kc = 0;
    for (int w = w0; w < n_items; w += wstep) {
      const int tile = w / a.H, head = w - tile * a.H;
      const TileInfo ti = a.tiles[tile];
      const uint16_t* __restrict__ blist = a.blk_list + ti.list_off;
      for (int j = 0; j < ti.n_list; ++j, ++kc) {
        const int st = kc % HST;
        const uint32_t par = (uint32_t)((kc / HST) - 1) & 1u;
        const size_t blk = (size_t)(ti.gblock0 + (int)(__ldg(blist + j) >> 1)) * a.H + head;
        if (kc >= HST) mbar_wait(&bs->k_empty[st], par);
        if (elect_one()) {
          mbar_expect_tx(&bs->k_full[st], KV_STAGE);
          bulk_load(k_sm + st * KV_STAGE, a.kimg + blk * kv_block_elems(HC_), KV_STAGE, &bs->k_full[st]);
        }
        __syncwarp();
        if (kc >= HST) mbar_wait(&bs->v_empty[st], par);
        if (elect_one()) {
          mbar_expect_tx(&bs->v_full[st], KV_STAGE);
          bulk_load(v_sm + st * KV_STAGE, a.vimg + blk * kv_block_elems(HC_), KV_STAGE, &bs->v_full[st]);
        }
        __syncwarp();
      }
    }
  } else if (role == 1) {  // ===== MMA issuer =====
    // A tcgen05.mma of M = 128, K = 16 costs ~45 cycles whatever N <= 64 is (profiles/r2_mma_issue_rate_b200.txt), and with
    // 18 of them per block the MMA count, not the softmax, paced this kernel.  The lo plane of a V block lies exactly
    // four 8-channel chunks behind its hi plane, so ONE N = 64 instruction with the hi plane's descriptor multiplies
    // P_hi with [V_hi | V_lo] into two accumulators (summed in the epilogue): 8 instead of 12 MMAs per P V.
    const uint32_t idesc_s = make_idesc(TM, TS);
    const uint32_t idesc_o2 = make_idesc(TM, 2 * HC_) | (1u << 16), idesc_o1 = make_idesc(TM, HC_) | (1u << 16);   // bit 16: B is MN-major
    const uint64_t dk0 = make_desc_nosw(smem_u32(k_sm), TS * 16, 128);
    const uint64_t dv0 = make_desc_nosw(smem_u32(v_sm), 128, TS * 16);
    constexpr uint32_t stage_u = KV_STAGE >> 4, plane_u = KV_PLANE >> 4;
    // S of running block gs from Q buffer qb; last_of_item: also tell the softmax warps that the Q buffer is free
    auto issue_s = [&](int gs, int qb, bool last_of_item) {
      mbar_wait(&bs->k_full[gs % HST], (uint32_t)(gs / HST) & 1u);
      if (gs >= 2) mbar_wait(&bs->pv_done[gs & 1], (uint32_t)((gs >> 1) - 1) & 1u);   // P of block gs - 2 consumed
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tq_hi = tmem_q0 + (uint32_t)(qb * HC_), tq_lo = tq_hi + HC_ / 2;
        const uint64_t dk_hi = dk0 + (uint32_t)(gs % HST) * stage_u, dk_lo = dk_hi + plane_u;
        const uint32_t d = tmem_s + (uint32_t)((gs & 1) * TS);
#pragma unroll
        for (int kk = 0; kk < HC_ / 16; ++kk) {
          const uint32_t ko = (uint32_t)kk * ((2 * TS * 16) >> 4);
          tc_mma_bf16_ts(d, tq_hi + kk * 8, dk_hi + ko, idesc_s, kk ? 1u : 0u);
          tc_mma_bf16_ts(d, tq_hi + kk * 8, dk_lo + ko, idesc_s, 1u);
          tc_mma_bf16_ts(d, tq_lo + kk * 8, dk_hi + ko, idesc_s, 1u);
        }
        tc_commit(&bs->s_full[gs & 1]);
        tc_commit(&bs->k_empty[gs % HST]);
        if (last_of_item) tc_commit(&bs->q_free[qb]);
      }
      __syncwarp();
    };
    if (w0 < n_items) {
      int g = 0, it = 0, w = w0;
      int nblk_cur = a.tiles[w / a.H].n_list;
      mbar_wait(&bs->q_full[0], 0);
      tc_fence_after();
      issue_s(0, 0, nblk_cur == 1);
      while (true) {
        const int w_next = w + wstep;
        const bool has_next = w_next < n_items;
        const int nblk_next = has_next ? a.tiles[w_next / a.H].n_list : 0;
        for (int j = 0; j < nblk_cur; ++j, ++g) {
          if (j + 1 < nblk_cur) issue_s(g + 1, it & 1, j + 2 == nblk_cur);   // S_{j+1} goes in front of P_j V_j
          const bool tr = DBG && blockIdx.x == 0 && lane == 0 && it < 8 && j < 16;
          long long* trp = DBG ? a.dbg + ((sidx * 8 + (it & 7)) * 16 + (j & 15)) * 8 : nullptr;
          if (tr) trp[4] = clock64();
          const int b = g & 1, vs = g % HST;
          mbar_wait(&bs->p_full[b], (uint32_t)(g >> 1) & 1u);
          if (tr) trp[5] = clock64();
          mbar_wait(&bs->v_full[vs], (uint32_t)(g / HST) & 1u);
          tc_fence_after();
          // (O is single-buffered: P_0 of this item was only published after the softmax warps had read the previous
          // item's O in their epilogue)
          if (elect_one()) {
            const uint32_t p_hi = tmem_s + (uint32_t)(b * TS), p_lo = p_hi + TS / 2;
            const uint64_t dv_hi = dv0 + (uint32_t)vs * stage_u;
            static_assert(KV_PLANE == 4 * TS * 16, "the lo plane must continue the hi plane's chunk sequence");
#pragma unroll
            for (int kk = 0; kk < TS / 16; ++kk) {
              tc_mma_bf16_ts(tmem_o, p_hi + kk * 8, dv_hi + kk * 16, idesc_o2, (j | kk) ? 1u : 0u);   // [O_a | O_b] += P_hi [V_hi | V_lo]
              tc_mma_bf16_ts(tmem_o, p_lo + kk * 8, dv_hi + kk * 16, idesc_o1, 1u);                    // O_a += P_lo V_hi
            }
            tc_commit(&bs->pv_done[b]);
            tc_commit(&bs->v_empty[vs]);
            if (j == nblk_cur - 1) tc_commit(&bs->o_full);
          }
          __syncwarp();
          if (tr) trp[6] = clock64();
        }
        if (!has_next) break;
        // the next item's first S as soon as its Q is parked (the softmax warps do that before their epilogue)
        mbar_wait(&bs->q_full[(it + 1) & 1], (uint32_t)((it + 1) >> 1) & 1u);
        tc_fence_after();
        issue_s(g, (it + 1) & 1, nblk_next == 1);
        ++it; w = w_next; nblk_cur = nblk_next;
      }
    }
  } else {  // ===== softmax warps =====
    const int q = warp & 3;
    const int r = q * 32 + lane;           // row in tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool leader = (q == 0) && (lane == 0);   // the thread that owns this stream's TMA traffic
    const uint32_t bar_id = 1 + (uint32_t)sidx;    // named barrier of the stream's 128 softmax threads
    const float c_log2 = 1.4426950408889634f / sqrtf((float)a.C);  // scale * log2(e)
    const float tau_raw = LAZY_LOG2 / c_log2;
    const int HC = a.H * a.C;
    // skip rows of an item: 4 boxes of 64 rows x 16 floats (64-byte swizzle); rows 0..63 then rows 64..127
    auto request_skip = [&](int wn) {
      const int tile = wn / a.H, head = wn - tile * a.H;
      const int node0 = a.tiles[tile].node0;
      mbar_expect_tx(&bs->skip_full, SKIP_BYTES);
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
          tma_load_2d(&map_skip, &bs->skip_full, skip_sm + (size_t)(half * 2 + bx) * 4096, 3 * HC + head * a.C + bx * 16, node0 + half * 64);
    };
    auto park = [&](int itn, int wn) {   // Q tile of item itn (work index wn) -> TMEM Q buffer itn & 1
      const int qb = itn & 1;
      if (itn >= 2) { mbar_wait(&bs->q_free[qb], (uint32_t)((itn >> 1) - 1) & 1u); tc_fence_after(); }
      const int tile = wn / a.H, head = wn - tile * a.H;
      const uint4* qsrc = reinterpret_cast<const uint4*>(a.qimg + ((size_t)tile * a.H + head) * q_block_elems(HC_));
      const uint32_t tq = tmem_q0 + (uint32_t)(qb * HC_) + lane_off;
      uint32_t qv[32];
#pragma unroll
      for (int pl_ = 0; pl_ < 2; ++pl_)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 t = __ldg(qsrc + (size_t)pl_ * (TM * HC_ / 8) + (size_t)u * TM + r);
          qv[16 * pl_ + 4 * u] = t.x; qv[16 * pl_ + 4 * u + 1] = t.y; qv[16 * pl_ + 4 * u + 2] = t.z; qv[16 * pl_ + 4 * u + 3] = t.w;
        }
      tmem_st16(tq, qv);
      tmem_st16(tq + 16, qv + 16);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bs->q_full[qb]);
    };
    if (w0 < n_items) {
      // Both streams would otherwise run in lock-step (scripts/trace_hidden.py: identical time stamps), fighting for the
      // ALU pipe in their score loops and leaving it idle together in every hand-off, Q park and epilogue.  Stream 1
      // starts late, so that each stream's latency-bound phases fall under the other one's arithmetic.
      if (sidx == 1 && a.stagger_ns > 0) __nanosleep((unsigned)a.stagger_ns);
      if (leader) request_skip(w0);
      park(0, w0);
    }
    int g = 0, it = 0;
    for (int w = w0; w < n_items; w += wstep, ++it) {
      const int tile = w / a.H, head = w - tile * a.H;
      const TileInfo ti = a.tiles[tile];
      const int nblk = ti.n_list;
      const uint16_t* __restrict__ blist = a.blk_list + ti.list_off;
      const bool row_valid = r < ti.rows;
      const uint32_t* bm_row = a.bitmap + ti.bm_off + (size_t)(ti.row0 + r) * ti.bm_words;
      const uint32_t o_t = tmem_o + lane_off;
      float m = -INFINITY, l = 0.f;  // m: reference point in raw-score units (>= true max - tau_raw)
      uint2 bits_next = row_valid ? *reinterpret_cast<const uint2*>(bm_row + (int)(__ldg(blist) >> 1) * 2) : make_uint2(0u, 0u);
      for (int j = 0; j < nblk; ++j, ++g) {
        const int b = g & 1;
        const uint2 bits = bits_next;
        if (j + 1 < nblk && row_valid) bits_next = *reinterpret_cast<const uint2*>(bm_row + (int)(__ldg(blist + j + 1) >> 1) * 2);
        const bool tr = DBG && blockIdx.x == 0 && q == 0 && lane == 0 && it < 8 && j < 16;
        long long* trp = DBG ? a.dbg + ((sidx * 8 + (it & 7)) * 16 + (j & 15)) * 8 : nullptr;
        if (tr) trp[0] = clock64();
        mbar_wait(&bs->s_full[b], (uint32_t)(g >> 1) & 1u);
        if (tr) trp[1] = clock64();
        tc_fence_after();
        const uint32_t s_addr = tmem_s + lane_off + (uint32_t)(b * TS);
        if (j == 0) {  // first block: take its masked max as the reference point
          uint32_t v[TS];
#pragma unroll
          for (int c0 = 0; c0 < TS; c0 += 16) tmem_ld16(s_addr + c0, v + c0);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < TS; ++e) {
            const uint32_t wd = (e < 32) ? bits.x : bits.y;
            if ((wd >> (e & 31)) & 1u) m = fmaxf(m, __uint_as_float(v[e]));
          }
        }
        bool prev_done = (j == 0);            // has pv_done of the previous block been observed?
        uint32_t ph[TS / 2], pl[TS / 2];      // P_j as packed bf16 pairs (hi and lo planes)
        float lsum, bmax;
        while (true) {
          const float m_sub = (m == -INFINITY) ? 0.f : m * c_log2;
          lsum = 0.f; bmax = -INFINITY;
          uint32_t v[TS];
#pragma unroll
          for (int c0 = 0; c0 < TS; c0 += 16) tmem_ld16(s_addr + c0, v + c0);
          tmem_ld_wait();
          if (tr) trp[2] = clock64();
#pragma unroll
          for (int e = 0; e < TS; e += 2) {
            const uint32_t wd = (e < 32) ? bits.x : bits.y;
            // masked scores become -inf once: max ignores them and ex2(-inf) = +0 exactly
            const float s0 = ((wd >> (e & 31)) & 1u) ? __uint_as_float(v[e]) : -INFINITY;
            const float s1 = ((wd >> ((e + 1) & 31)) & 1u) ? __uint_as_float(v[e + 1]) : -INFINITY;
            if constexpr (!NOMAX) bmax = fmaxf(bmax, fmaxf(s0, s1));
            const float p0 = ex2_approx(fmaf(s0, c_log2, -m_sub));
            const float p1 = ex2_approx(fmaf(s1, c_log2, -m_sub));
            lsum += p0 + p1;
            const uint32_t h2 = pack_bf16x2(p0, p1);
            ph[e >> 1] = h2;
            pl[e >> 1] = pack_bf16x2(p0 - __uint_as_float(h2 << 16), p1 - __uint_as_float(h2 & 0xffff0000u));
          }
          bool exceeded;
          if constexpr (NOMAX) exceeded = !(lsum < 0x1p60f) || (m == -INFINITY && (bits.x | bits.y) != 0u);
          else exceeded = bmax > m + tau_raw;   // also true when m == -inf and the block has an edge
          if (!__any_sync(0xffffffffu, exceeded)) break;
          if constexpr (NOMAX) {   // the block's masked maximum, from the scores still in TMEM (P is only written after the loop)
            bmax = m;
#pragma unroll 1
            for (int c0 = 0; c0 < TS; c0 += 16) {
              uint32_t v2[16];
              tmem_ld16(s_addr + c0, v2);
              tmem_ld_wait();
              const uint32_t wd = (c0 < 32) ? bits.x : bits.y;
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if ((wd >> ((c0 + e) & 31)) & 1u) bmax = fmaxf(bmax, __uint_as_float(v2[e]));
            }
          }
          // rare: raise the reference point, rescale the history (l and O in TMEM), redo this block
          const float m_new = exceeded ? bmax : m;
          const float alpha = (m == -INFINITY) ? 0.f : ex2_approx((m - m_new) * c_log2);
          if (!prev_done) {
            mbar_wait(&bs->pv_done[(g - 1) & 1], (uint32_t)((g - 1) >> 1) & 1u);   // O holds every earlier block of this item
            tc_fence_after();
            prev_done = true;
          }
          if (j > 0) {
#pragma unroll
            for (int c0 = 0; c0 < 2 * HC_; c0 += 16) {
              uint32_t o[16];
              tmem_ld16(o_t + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st16(o_t + c0, o);
            }
            tmem_st_wait();
          }
          l *= alpha;
          m = m_new;
        }
        l += lsum;
        // P_j replaces S_j in place (every S value of this row is already in registers): hi | lo planes
        tmem_st16(s_addr, ph);
        tmem_st16(s_addr + 16, ph + 16);
        tmem_st16(s_addr + 32, pl);
        tmem_st16(s_addr + 48, pl + 16);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bs->p_full[b]);
        if (tr) trp[3] = clock64();
      }
      const bool tri = DBG && blockIdx.x == 0 && q == 0 && lane == 0 && it < 8;
      long long* trq = DBG ? a.dbg + 2048 + (sidx * 8 + (it & 7)) * 4 : nullptr;
      if (tri) trq[0] = clock64();
      // ---- between items: the next item's Q tile first (the MMA warp issues its first S under our epilogue) ----
      const int w_next = w + wstep;
      const bool has_next = w_next < n_items;
      if (has_next) park(it + 1, w_next);
      if (tri) trq[1] = clock64();
      // ---- epilogue: O / l + skip -> activation -> bf16 hi / lo planes ----
      // the previous item's tensor stores must have read the staging planes before anyone overwrites them
      if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      mbar_wait(&bs->o_full, (uint32_t)it & 1u);
      if (tri) trq[2] = clock64();
      tc_fence_after();
      uint32_t o[HC_];
      {
        uint32_t ob[HC_];
        tmem_ld16(o_t, o);
        tmem_ld16(o_t + 16, o + 16);
        tmem_ld16(o_t + 32, ob);
        tmem_ld16(o_t + 48, ob + 16);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < HC_; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) + __uint_as_float(ob[e]));
      }
      mbar_wait(&bs->skip_full, (uint32_t)it & 1u);
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // (orders the leader's wait_group before the staging writes)
      const float inv = 1.f / (l + 1e-16f);   // rows without an in-edge: O = 0 exactly
      const bool tma_out = ti.rows == TM;
      const int node = ti.node0 + r;
      const int rr = r & 63;
      const uint8_t* skip_row = skip_sm + (size_t)(r >> 6) * 8192 + (size_t)rr * 64;
      const uint32_t swz = (uint32_t)((rr >> 1) & 3);
#pragma unroll
      for (int c0 = 0; c0 < HC_; c0 += 16) {
        float y[16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 t = *reinterpret_cast<const float4*>(skip_row + (size_t)(c0 >> 4) * 4096 + (((uint32_t)u ^ swz) << 4));
          y[4 * u] = t.x; y[4 * u + 1] = t.y; y[4 * u + 2] = t.z; y[4 * u + 3] = t.w;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) y[e] = fmaf(__uint_as_float(o[c0 + e]), inv, y[e]);
        if (a.act != ACT_NONE) {
#pragma unroll
          for (int e = 0; e < 16; ++e) y[e] = apply_act_rt(y[e], a.act);
        }
        uint32_t hh[8], ll[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const uint32_t h2 = pack_bf16x2(y[e], y[e + 1]);
          hh[e >> 1] = h2;
          ll[e >> 1] = pack_bf16x2(y[e] - __uint_as_float(h2 << 16), y[e + 1] - __uint_as_float(h2 & 0xffff0000u));
        }
        if (tma_out) {
          uint4* dh = reinterpret_cast<uint4*>(ohi_sm + (size_t)r * (HC_ * 2) + (size_t)c0 * 2);
          uint4* dl = reinterpret_cast<uint4*>(olo_sm + (size_t)r * (HC_ * 2) + (size_t)c0 * 2);
          dh[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]); dh[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
          dl[0] = make_uint4(ll[0], ll[1], ll[2], ll[3]); dl[1] = make_uint4(ll[4], ll[5], ll[6], ll[7]);
        } else if (row_valid) {   // the last tile of a graph: valid rows straight to global memory
          __nv_bfloat16* dh = a.out.hi + (size_t)node * a.out.ld_split + head * a.C + c0;
          __nv_bfloat16* dl = a.out.lo + (size_t)node * a.out.ld_split + head * a.C + c0;
          *reinterpret_cast<uint4*>(dh) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<uint4*>(dh + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
          *reinterpret_cast<uint4*>(dl) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          *reinterpret_cast<uint4*>(dl + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
        }
      }
      // generic-proxy writes -> async proxy; after the barrier every thread has also finished reading the skip rows
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (leader) {
        if (tma_out) {
          tma_store_2d(&map_ohi, ohi_sm, head * a.C, ti.node0);
          tma_store_2d(&map_olo, olo_sm, head * a.C, ti.node0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (has_next) request_skip(w_next);
      }
      if (tri) trq[3] = clock64();
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA's smem goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

}  // namespace

bool attn_hidden_persist_supported(const AttnDenseArgs& a) {
  return a.C == HC_ && a.Cpad == HC_ && a.H > 0 && a.row_fused != nullptr && a.resid == nullptr && a.out.hi != nullptr &&
         a.out.lo != nullptr && a.out.f32 == nullptr && a.out.ld_split % 8 == 0 && a.ld % 4 == 0 && a.blk_list != nullptr;
}

cudaError_t launch_attn_hidden_persist(const AttnDenseArgs& a, cudaStream_t s) {
  if (a.n_tiles <= 0) return cudaSuccess;
  if (!attn_hidden_persist_supported(a)) return cudaErrorInvalidValue;
  CUtensorMap map_skip, map_ohi, map_olo;
  if (!get_tensor_map_2d(a.qkvs, 4, a.n_rows, a.ld, a.ld, 64, 16, 1, &map_skip) ||
      !get_tensor_map_2d(a.out.hi, 2, a.n_rows_out, a.out.ld_split, a.out.ld_split, 128, a.C, 0, &map_ohi) ||
      !get_tensor_map_2d(a.out.lo, 2, a.n_rows_out, a.out.ld_split, a.out.ld_split, 128, a.C, 0, &map_olo))
    return cudaErrorInvalidValue;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int items = a.n_tiles * a.H;
  const unsigned grid = (unsigned)((items + 1) / 2 < sms ? (items + 1) / 2 : sms);
  const size_t smem_bytes = 2 * (size_t)STREAM_BYTES + sizeof(HidSmem) + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_hidden_persist_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_hidden_persist_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_hidden_persist_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  AttnDenseArgs b = a;
  if (items < 8 * (int)grid) b.stagger_ns = 0;   // short streams (small batches): the start delay would cost more than it hides
  {  // A/B aid: DA_HIDDEN_TRACE_VARIANT=1 runs the tracing instantiation (into a scratch buffer) on every launch
    static int force = -1;
    static long long* scratch = nullptr;
    if (force < 0) {
      const char* e = getenv("DA_HIDDEN_TRACE_VARIANT");
      force = (e != nullptr && e[0] == '1') ? 1 : 0;
      if (force && cudaMalloc(&scratch, 4096 * sizeof(long long)) != cudaSuccess) force = 0;
    }
    if (force && !b.dbg) b.dbg = scratch;
  }
  // Score loop without a running maximum unless DA_HIDDEN_NOMAX=0 (read per launch).  A/B on B200, three hidden launches of
  // the c3 step: 0.446 ms with the running maximum, 0.421 ms without; a first version that kept the 64 scores in registers
  // for the rare path spilled and took 0.522 ms.
  const char* e_nomax = getenv("DA_HIDDEN_NOMAX");
  const bool nomax = !(e_nomax != nullptr && e_nomax[0] == '0');
  if (b.dbg) return launch_pdl(attn_hidden_persist_kernel<true, false>, dim3(grid), dim3(NTH), smem_bytes, s, map_skip, map_ohi, map_olo, b);
  if (nomax) return launch_pdl(attn_hidden_persist_kernel<false, true>, dim3(grid), dim3(NTH), smem_bytes, s, map_skip, map_ohi, map_olo, b);
  return launch_pdl(attn_hidden_persist_kernel<false, false>, dim3(grid), dim3(NTH), smem_bytes, s, map_skip, map_ohi, map_olo, b);
}

}  // namespace da
