// Masked dense graph attention on tcgen05 tensor cores (attn_mode = DA_ATTN_AUTO), with the layer epilogue fused.
//
// For the graphs selected by the planner (plan.cu) the TransformerConv message/softmax/aggregate
// stage is evaluated as flash-style tiled attention restricted by the
// graph's adjacency bitmap:
//     S = Q K^T               128 targets x 64 sources per tile, fp32 accumulators in TMEM
//     P = exp(scale*S - m)    only where bitmap(i, j) = 1 (online running reference point / sum per target row)
//     O += P V                fp32 accumulator in TMEM, rescaled when the reference point moves
// Both products use the 3-pass split-bf16 scheme of gemm_umma.cu (hi*hi + hi*lo + lo*hi) so the
// result stays within ~1e-5 of fp32.  The operands come from "images" written by the QKVS GEMM's epilogue
// (pack_images_kernel in exact-fp32 mode): per (tile, head) contiguous blocks already in the canonical
// no-swizzle K-major core-matrix layout of tcgen05.mma (8 rows x 16 bytes per core matrix), so a K or V
// block is ONE cp.async.bulk.  The planner's promoted residual edges are bitmap bits on extra columns whose
// K / V rows gather_extra_kernel copies into the padding rows of the images.
//
// CTA = (128-row target tile, head).  Warp roles:
//   warp 0     bulk-copy producer (K / V blocks through mbarrier rings)
//   warp 1     TMEM allocation, single-thread tcgen05.mma issue (S_{j+1} is issued before P_j V_j
//              so the tensor core works while the softmax warps are busy)
//   warps 2-5  softmax: thread == target row (TMEM lane), so the row max / sum need no shuffles;
//              Q parked in TMEM once, one pass over each S tile, P written back into the S columns;
//              then the epilogue: O / l + skip (+ trunk residual) -> activation -> bf16 hi / lo split.
//              skip / residual rows arrive as swizzled TMA boxes in ring stages that have gone idle,
//              full tiles leave through TMA tensor stores.
// Rows the kernel cannot finalise (more than DA_FUSE_MAX_RESIDUAL residual in-edges) get their un-normalised O and
// (m, l) written to global memory instead; attn_csr.cu continues the same online softmax for them.
#include <cstring>
#include <cstdlib>

#include "attn_tc.cuh"

namespace da {
namespace {


__global__ void pack_images_kernel(PackArgs a) {
  // one warp = 32 consecutive nodes (lane = node) x a strided set of (part, head, 8-channel chunk)
  const int lane = threadIdx.x & 31;
  const int node = blockIdx.x * 32 + lane;
  const int HC = a.H * a.C;
  const int chunks = a.Cpad / 8;                 // per head, including zero padding up to Cpad
  const int items_per_part = a.H * chunks;
  const int warp_in_cta = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int slot = node < a.n ? a.node_slot[node] : -1;
  if (slot < 0) return;
  const int tile = slot >> 7, r = slot & 127;
  const int blk = slot >> 6, rb = slot & 63;
  const float* row = a.qkvs + (size_t)node * a.ld;
  for (int it = warp_in_cta + blockIdx.y * warps; it < 3 * items_per_part; it += warps * gridDim.y) {
    const int part = it / items_per_part, rem = it % items_per_part;
    const int h = rem / chunks, ch = rem % chunks;
    const int c = ch * 8;
    __nv_bfloat16 hi[8], lo[8];
    if (c < a.C) {  // C % 8 == 0: a chunk is entirely data or entirely padding
      const float* src = row + part * HC + h * a.C + c;
      const float4 v0 = *reinterpret_cast<const float4*>(src);
      const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        hi[e] = __float2bfloat16_rn(v[e]);
        lo[e] = __float2bfloat16_rn(v[e] - __bfloat162float(hi[e]));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) { hi[e] = __float2bfloat16_rn(0.f); lo[e] = hi[e]; }
    }
    if (part == 0) {
      __nv_bfloat16* base = a.qimg + ((size_t)tile * a.H + h) * q_block_elems(a.Cpad);
      const size_t off = (size_t)ch * (TM * 8) + (size_t)r * 8;
      *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(base + (size_t)TM * a.Cpad + off) = *reinterpret_cast<uint4*>(lo);
    } else {  // K and V share one layout; V is consumed as an MN-major B operand
      __nv_bfloat16* base = (part == 1 ? a.kimg : a.vimg) + ((size_t)blk * a.H + h) * kv_block_elems(a.Cpad);
      const size_t off = (size_t)ch * (TS * 8) + (size_t)rb * 8;
      *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
      *reinterpret_cast<uint4*>(base + (size_t)TS * a.Cpad + off) = *reinterpret_cast<uint4*>(lo);
    }
  }
}

// K / V rows of the plan's extra sources -> their (padding) image rows.  One warp per (entry, part); lanes over
// (head, 8-channel chunk) items.
__global__ void gather_extra_kernel(PackArgs a, const int32_t* __restrict__ x_src, const int32_t* __restrict__ x_slot, int n_extra) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wg = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (wg >= n_extra * 2) return;
  const int ent = wg >> 1, part = 1 + (wg & 1);   // 1 = K, 2 = V
  const int node = x_src[ent], slot = x_slot[ent];
  // folded last layer: the V' part has its own head dim (rows are [Q | K | V'], so its offset is still 2 * H * C)
  const int C_ = (part == 2 && a.Cv > 0) ? a.Cv : a.C, Cpad_ = (part == 2 && a.Cv > 0) ? a.Cvpad : a.Cpad;
  const int HC = a.H * a.C, chunks = Cpad_ / 8;
  const int blk = slot >> 6, rb = slot & 63;
  const float* row = a.qkvs + (size_t)node * a.ld + part * HC;
  for (int it = lane; it < a.H * chunks; it += 32) {
    const int h = it / chunks, ch = it % chunks, c = ch * 8;
    __nv_bfloat16 hi[8], lo[8];
    if (c < C_) {
      const float4 v0 = *reinterpret_cast<const float4*>(row + h * C_ + c);
      const float4 v1 = *reinterpret_cast<const float4*>(row + h * C_ + c + 4);
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        hi[e] = __float2bfloat16_rn(v[e]);
        lo[e] = __float2bfloat16_rn(v[e] - __bfloat162float(hi[e]));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) { hi[e] = __float2bfloat16_rn(0.f); lo[e] = hi[e]; }
    }
    __nv_bfloat16* base = (part == 1 ? a.kimg : a.vimg) + ((size_t)blk * a.H + h) * kv_block_elems(Cpad_);
    const size_t off = (size_t)ch * (TS * 8) + (size_t)rb * 8;
    *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(base + (size_t)TS * Cpad_ + off) = *reinterpret_cast<uint4*>(lo);
  }
}

struct DenseSmem {
  uint64_t q_full;
  uint64_t k_full[MAXST], k_empty[MAXST];
  uint64_t v_full[MAXST], v_empty[MAXST];
  uint64_t s_full[2];    // MMA -> softmax: S_j is in TMEM buffer j & 1
  uint64_t p_full[2];    // softmax -> MMA: P_j (bf16 hi/lo) has replaced S_j in the same TMEM columns
  uint64_t pv_done[2];   // MMA -> both:   P_j V_j retired (buffer free again, O holds blocks <= j)
  uint64_t epi_full;     // fused epilogue: every row's skip values have landed in the (by then idle) K ring
  uint32_t tmem_base;
};


// CPAD_T: padded head dim as a compile-time constant (0 = take it from the arguments); ST: K / V ring
// depth (power of two).  The single MMA-issuing thread is on the critical path of every block, so its
// loops are fully unrolled and every operand descriptor is "base + immediate".
template <int CPAD_T, int ST>
__global__ void __launch_bounds__(NT)
attn_dense_kernel(const __grid_constant__ CUtensorMap map_skip, const __grid_constant__ CUtensorMap map_resid,
                  const __grid_constant__ CUtensorMap map_ohi, const __grid_constant__ CUtensorMap map_olo,
                  AttnDenseArgs a, int tmem_cols, int stage_flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int KST = ST, VST = ST;
  const int Cpad = CPAD_T ? CPAD_T : a.Cpad;
  const uint32_t kv_plane = TS * Cpad * 2;                 // bytes
  uint8_t* k_sm = smem;                                    // KST stages x 2 planes
  uint8_t* v_sm = k_sm + KST * 2 * kv_plane;              // VST stages x 2 planes
  DenseSmem* sh = reinterpret_cast<DenseSmem*>(v_sm + VST * 2 * kv_plane);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / a.H, head = blockIdx.x % a.H;
  const TileInfo ti = a.tiles[tile];
  // the source blocks this tile visits: the planner's list of blocks that hold at least one edge of the tile (in the
  // planner's internal node order the adjacency of an Exphander graph is a band, so ~1/4 of the blocks drop out at 60 %
  // density and ~2/3 at 20 %).  Every role walks the same list; ring stages / TMEM buffers are indexed by the POSITION j.
  const int nblk = ti.n_list;
  const uint16_t* __restrict__ blist = a.blk_list + ti.list_off;
  // development aid: per-CTA (start, end, SM id) records behind the 128-entry trace of CTA 0 when dbg[127] says there is room
  const bool dbg_rec = a.dbg != nullptr && threadIdx.x == 0 && a.dbg[127] > (long long)blockIdx.x;
  if (dbg_rec) {
    unsigned long long gt; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    a.dbg[128 + 8 * blockIdx.x + 0] = (long long)gt; a.dbg[128 + 8 * blockIdx.x + 2] = smid;
    a.dbg[128 + 8 * blockIdx.x + 3] = clock64();
  }
  const bool dbg_sm = a.dbg != nullptr && threadIdx.x == 64 && a.dbg[127] > (long long)blockIdx.x;   // first softmax thread

  if (threadIdx.x == 0) {
    mbar_init(&sh->q_full, 4);   // the four softmax warps have parked Q in TMEM
    for (int i = 0; i < MAXST; ++i) {
      mbar_init(&sh->k_full[i], 1); mbar_init(&sh->k_empty[i], 1);
      mbar_init(&sh->v_full[i], 1); mbar_init(&sh->v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh->s_full[i], 1); mbar_init(&sh->p_full[i], 4); mbar_init(&sh->pv_done[i], 1); }
    mbar_init(&sh->epi_full, 1);     // the thread that issues the staging copies
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"((uint32_t)tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  if (dbg_sm) a.dbg[128 + 8 * blockIdx.x + 4] = clock64();   // setup (barriers, TMEM allocation) done
  const uint32_t tmem_s = tmem_base;             // 2 x TS columns: S_j (fp32), later P_j (bf16 hi | lo)
  const uint32_t tmem_o = tmem_base + 2 * TS;    // Cpad columns
  const uint32_t tmem_q = tmem_o + Cpad;         // Cpad columns: Q as packed bf16 pairs, hi plane then lo plane

  if (warp == 0) {  // ===== bulk-copy producer (warp-uniform): K and V blocks run ST ahead =====
    auto blk_off = [&](int j) { return ((size_t)(ti.gblock0 + (int)(__ldg(blist + j) >> 1)) * a.H + head) * kv_block_elems(Cpad); };
    auto load_k = [&](int j) {
      if (elect_one()) {
        const int st = j % KST;
        mbar_expect_tx(&sh->k_full[st], 2 * kv_plane);
        bulk_load(k_sm + st * 2 * kv_plane, a.kimg + blk_off(j), 2 * kv_plane, &sh->k_full[st]);
      }
      __syncwarp();
    };
    auto load_v = [&](int j) {
      if (elect_one()) {
        const int st = j % VST;
        mbar_expect_tx(&sh->v_full[st], 2 * kv_plane);
        bulk_load(v_sm + st * 2 * kv_plane, a.vimg + blk_off(j), 2 * kv_plane, &sh->v_full[st]);
      }
      __syncwarp();
    };
    for (int j = 0; j < KST && j < nblk; ++j) load_k(j);
    for (int j = 0; j < VST && j < nblk; ++j) load_v(j);
    for (int j = 0; j < nblk; ++j) {
      if (j + KST < nblk) {  // slot j % KST is free once S_j (its (j / KST)-th user) has retired
        mbar_wait(&sh->k_empty[j % KST], (j / KST) & 1);
        load_k(j + KST);
      }
      if (j + VST < nblk) {
        mbar_wait(&sh->v_empty[j % VST], (j / VST) & 1);
        load_v(j + VST);
      }
    }
  } else if (warp == 1) {  // ===== MMA issuer (warp-uniform; one elected lane issues) =====
    const uint32_t idesc_s = make_idesc(TM, TS), idesc_o = make_idesc(TM, Cpad) | (1u << 16);  // bit 16: B is MN-major
    const int ksteps = Cpad / 16;
    const uint32_t tq_hi = tmem_q, tq_lo = tmem_q + Cpad / 2;   // A operand from TMEM: 8 columns per k-step
    // descriptor bases; per-MMA descriptors are base + (byte offset >> 4) in the 14-bit address field
    const uint64_t dk0 = make_desc_nosw(smem_u32(k_sm), TS * 16, 128);
    const uint64_t dv0 = make_desc_nosw(smem_u32(v_sm), 128, TS * 16);   // MN-major B: LBO = next 8 sources, SBO = next 8 channels
    const uint32_t stage_u = (2 * kv_plane) >> 4, plane_u = kv_plane >> 4;
    auto issue_s = [&](int j) {
      if (elect_one()) {
        const uint64_t dk_hi = dk0 + (uint32_t)(j % KST) * stage_u, dk_lo = dk_hi + plane_u;
        const uint32_t d = tmem_s + (uint32_t)((j & 1) * TS);
#pragma unroll
        for (int kk = 0; kk < (CPAD_T ? CPAD_T / 16 : 16); ++kk) {
          if (!CPAD_T && kk >= ksteps) break;
          // one k-step = 16 channels = 2 K chunks (stride TS*16 B) = 8 TMEM columns of Q
          const uint32_t ko = (uint32_t)kk * ((2 * TS * 16) >> 4);
          tc_mma_bf16_ts(d, tq_hi + kk * 8, dk_hi + ko, idesc_s, kk ? 1u : 0u);
          tc_mma_bf16_ts(d, tq_hi + kk * 8, dk_lo + ko, idesc_s, 1u);
          tc_mma_bf16_ts(d, tq_lo + kk * 8, dk_hi + ko, idesc_s, 1u);
        }
        tc_commit(&sh->s_full[j & 1]);
        tc_commit(&sh->k_empty[j % KST]);
      }
      __syncwarp();
    };
    mbar_wait(&sh->q_full, 0);   // Q is in TMEM
    mbar_wait(&sh->k_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nblk; ++j) {
      if (j + 1 < nblk) {
        const int jn = j + 1;
        mbar_wait(&sh->k_full[jn % KST], (jn / KST) & 1);
        if (jn >= 2) mbar_wait(&sh->pv_done[jn & 1], ((jn >> 1) - 1) & 1);  // P_{jn-2} consumed: buffer free
        tc_fence_after();
        if (a.dbg && blockIdx.x == 0 && lane == 0 && j < 15) a.dbg[0 * 64 + j * 4 + 0] = clock64();   // S_{j+1} issue
        issue_s(jn);
      }
      const int b = j & 1, vs = j % VST;
      mbar_wait(&sh->p_full[b], (j >> 1) & 1);   // P_j in TMEM (and O corrected if needed)
      if (a.dbg && blockIdx.x == 0 && lane == 0 && j < 15) a.dbg[0 * 64 + j * 4 + 1] = clock64();     // p_full observed
      mbar_wait(&sh->v_full[vs], (j / VST) & 1);
      if (a.dbg && blockIdx.x == 0 && lane == 0 && j < 15) a.dbg[0 * 64 + j * 4 + 2] = clock64();     // v_full observed
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_hi = tmem_s + (uint32_t)(b * TS), p_lo = p_hi + TS / 2;   // packed bf16: 2 sources per column
        const uint64_t dv_hi = dv0 + (uint32_t)vs * stage_u, dv_lo = dv_hi + plane_u;
#pragma unroll
        for (int kk = 0; kk < TS / 16; ++kk) {
          // one k-step = 16 sources = 256 B of the V block = 8 TMEM columns of P
          tc_mma_bf16_ts(tmem_o, p_hi + kk * 8, dv_hi + kk * 16, idesc_o, (j | kk) ? 1u : 0u);
          tc_mma_bf16_ts(tmem_o, p_hi + kk * 8, dv_lo + kk * 16, idesc_o, 1u);
          tc_mma_bf16_ts(tmem_o, p_lo + kk * 8, dv_hi + kk * 16, idesc_o, 1u);
        }
        tc_commit(&sh->pv_done[b]);
        tc_commit(&sh->v_empty[vs]);
      }
      __syncwarp();
    }
  } else {  // ===== softmax warps =====
    const int q = warp & 3;
    const int r = q * 32 + lane;           // row in tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool row_valid = r < ti.rows;
    const float c_log2 = 1.4426950408889634f / sqrtf((float)a.C);  // scale * log2(e)
    const float tau_raw = LAZY_LOG2 / c_log2;
    const uint32_t* bm_row = a.bitmap + ti.bm_off + (size_t)(ti.row0 + r) * ti.bm_words;
    {  // park this row's Q (split-bf16 image, [plane][chunk][row][8]) in TMEM: A operand of every S = Q K^T
      const uint4* qsrc = reinterpret_cast<const uint4*>(a.qimg + ((size_t)tile * a.H + head) * q_block_elems(Cpad));
      for (int pl_ = 0; pl_ < 2; ++pl_) {
        for (int ch0 = 0; ch0 < Cpad / 8; ch0 += 4) {   // 4 chunks = 32 channels = 16 TMEM columns per store
          uint32_t qv[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (ch0 + u < Cpad / 8) t = __ldg(qsrc + (size_t)pl_ * (TM * Cpad / 8) + (size_t)(ch0 + u) * TM + r);
            qv[4 * u] = t.x; qv[4 * u + 1] = t.y; qv[4 * u + 2] = t.z; qv[4 * u + 3] = t.w;
          }
          const uint32_t dst = tmem_q + lane_off + (uint32_t)(pl_ * (Cpad / 2) + ch0 * 4);
          if (ch0 + 4 <= Cpad / 8) tmem_st16(dst, qv);
          else {  // tail of 8 columns (Cpad = 16 mod 32)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(dst), "r"(qv[0]),
                         "r"(qv[1]), "r"(qv[2]), "r"(qv[3]), "r"(qv[4]), "r"(qv[5]), "r"(qv[6]), "r"(qv[7]) : "memory");
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->q_full);
      if (dbg_sm) a.dbg[128 + 8 * blockIdx.x + 5] = clock64();   // Q parked
    }
    const int node = ti.node0 + r;
    const int HC = a.H * a.C;
    // fused finalisation (see the epilogue): this row's skip values are staged in the K ring, one row per
    // thread, padded by 16 bytes so that the per-row float4 reads are bank-conflict free
    const bool fused = a.row_fused != nullptr && row_valid && a.row_fused[node] != 0;
    // Staging (stage_flags bit 0: trunk residual staged, bit 1: outputs leave through TMA stores).  Skip / residual
    // values arrive as TMA boxes of 64 rows x 16 floats with the 64-byte swizzle, so that the row-per-thread float4
    // reads below are bank-conflict free; a half tile (64 rows) of Cpad/16 boxes is exactly one K / V ring stage:
    //   skip rows 0..63 -> K stage 0, skip rows 64..127 -> K stage 1, residual rows 0..63 -> K stage 2,
    //   residual rows 64..127 -> V stage n % ST (idle during the last two blocks).
    // Each region is requested as soon as the last block that uses its stage has retired its S = Q K^T.
    const bool stage_resid = a.resid != nullptr && (stage_flags & 1);
    const uint32_t stage_b = 2 * kv_plane;                  // bytes of one ring stage = (Cpad / 16) boxes of 4 KB
    const int nbox = Cpad / 16;
    const int rr = r & 63;
    const uint32_t row_off = (uint32_t)rr * 64u;
    const uint32_t swz = (uint32_t)((rr >> 1) & 3);
    const uint8_t* skip_reg = k_sm + (size_t)(r >> 6) * stage_b;
    const uint8_t* resid_reg = (r < 64) ? k_sm + (size_t)2 * stage_b : v_sm + (size_t)(nblk % VST) * stage_b;
    auto staged4 = [&](const uint8_t* reg, int box, int u) -> float4 {   // float4 u of this row's chunk `box`
      return *reinterpret_cast<const float4*>(reg + (size_t)box * 4096 + row_off + (((uint32_t)u ^ swz) << 4));
    };
    auto issue_region = [&](const CUtensorMap* map, uint8_t* dst, int col0, int row0) {
      for (int bx = 0; bx < nbox; ++bx) tma_load_2d(map, &sh->epi_full, dst + (size_t)bx * 4096, col0 + bx * 16, row0);
    };
    auto last_user = [&](int st_) { return st_ < nblk ? (nblk - 1) - ((nblk - 1 - st_) % KST) : 0; };
    // residual in-edges of a fused row (<= DA_FUSE_MAX_RESIDUAL; none at all once the planner has promoted them)
    int res_beg = 0, n_e = 0;
    if (fused) { res_beg = a.rowptr[node]; n_e = a.rowptr[node + 1] - res_beg; }
    float m = -INFINITY, l = 0.f;  // m: reference point in raw-score units (>= true max - tau_raw)
    uint2 bits_next = row_valid ? *reinterpret_cast<const uint2*>(bm_row + (int)(__ldg(blist) >> 1) * 2) : make_uint2(0u, 0u);
    for (int j = 0; j < nblk; ++j) {
      const int b = j & 1;
      const uint2 bits = bits_next;
      if (j + 1 < nblk && row_valid) bits_next = *reinterpret_cast<const uint2*>(bm_row + (int)(__ldg(blist + j + 1) >> 1) * 2);
      if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64 && j < 15) a.dbg[1 * 64 + j * 4 + 0] = clock64();   // start waiting for S_j
      mbar_wait(&sh->s_full[b], (j >> 1) & 1);
      if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64 && j < 15) a.dbg[1 * 64 + j * 4 + 1] = clock64();   // S_j ready
      tc_fence_after();
      if (a.row_fused != nullptr && warp == 2) {
        // S_j has retired: K stages whose last user is block j (or that are never used) are idle from now on
        if (elect_one()) {
          if (j == 0) mbar_expect_tx(&sh->epi_full, (stage_resid ? 4u : 2u) * (uint32_t)nbox * 4096u);   // the one arrival
          if (j == last_user(0)) issue_region(&map_skip, k_sm, 3 * HC + head * a.C, ti.node0);
          if (j == last_user(1)) issue_region(&map_skip, k_sm + stage_b, 3 * HC + head * a.C, ti.node0 + 64);
          if (stage_resid) {
            if (j == last_user(2)) issue_region(&map_resid, k_sm + 2 * stage_b, head * a.C, ti.node0);
            // (S_{n-1} was only issued after P_{n-3} V_{n-3} retired, so V stage n % ST is idle by now)
            if (j == nblk - 1) issue_region(&map_resid, v_sm + (size_t)(nblk % VST) * stage_b, head * a.C, ti.node0 + 64);
          }
        }
        __syncwarp();
      }
      const uint32_t s_addr = tmem_s + lane_off + (uint32_t)(b * TS);
      if (j == 0) {  // first block: take its masked max as the reference point
        uint32_t v[TS];
#pragma unroll
        for (int c0 = 0; c0 < TS; c0 += 16) tmem_ld16(s_addr + c0, v + c0);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < TS; ++e) {
          const uint32_t w = (e < 32) ? bits.x : bits.y;
          if ((w >> (e & 31)) & 1u) m = fmaxf(m, __uint_as_float(v[e]));
        }
      }
      bool prev_done = (j == 0);            // has pv_done of block j-1 been observed?
      uint32_t ph[TS / 2], pl[TS / 2];      // P_j as packed bf16 pairs (hi and lo planes)
      float lsum, bmax;
      while (true) {
        const float m_sub = (m == -INFINITY) ? 0.f : m * c_log2;
        lsum = 0.f; bmax = -INFINITY;
        uint32_t v[TS];
#pragma unroll
        for (int c0 = 0; c0 < TS; c0 += 16) tmem_ld16(s_addr + c0, v + c0);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < TS; e += 2) {
          const uint32_t w = (e < 32) ? bits.x : bits.y;
          // masked scores become -inf once: max ignores them and ex2(-inf) = +0 exactly
          const float s0 = ((w >> (e & 31)) & 1u) ? __uint_as_float(v[e]) : -INFINITY;
          const float s1 = ((w >> ((e + 1) & 31)) & 1u) ? __uint_as_float(v[e + 1]) : -INFINITY;
          bmax = fmaxf(bmax, fmaxf(s0, s1));
          const float p0 = ex2_approx(fmaf(s0, c_log2, -m_sub));
          const float p1 = ex2_approx(fmaf(s1, c_log2, -m_sub));
          lsum += p0 + p1;
          const uint32_t h2 = pack_bf16x2(p0, p1);
          ph[e >> 1] = h2;
          pl[e >> 1] = pack_bf16x2(p0 - __uint_as_float(h2 << 16), p1 - __uint_as_float(h2 & 0xffff0000u));
        }
        const bool exceeded = bmax > m + tau_raw;   // also true when m == -inf and the block has an edge
        if (!__any_sync(0xffffffffu, exceeded)) break;
        // rare: raise the reference point, rescale the history (l and O in TMEM), redo this block
        const float m_new = exceeded ? bmax : m;
        const float alpha = (m == -INFINITY) ? 0.f : ex2_approx((m - m_new) * c_log2);
        if (!prev_done) {
          mbar_wait(&sh->pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);   // O holds every block < j
          tc_fence_after();
          prev_done = true;
        }
        if (j > 0) {
          for (int c0 = 0; c0 < Cpad; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_o + lane_off + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st16(tmem_o + lane_off + c0, o);
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
      if (lane == 0) mbar_arrive(&sh->p_full[b]);
      if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64 && j < 15) a.dbg[1 * 64 + j * 4 + 2] = clock64();   // P_j published
    }
    // epilogue
    if (dbg_sm) a.dbg[128 + 8 * blockIdx.x + 6] = clock64();   // softmax loop done
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64) a.dbg[120] = clock64();   // last P published
    mbar_wait(&sh->pv_done[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);
    tc_fence_after();
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64) a.dbg[121] = clock64();   // last PV retired
    // ---- fused finalisation (rows flagged by the planner): continue the online softmax over this row's
    // (<= DA_FUSE_MAX_RESIDUAL) residual in-edges, normalise, + skip (+ trunk residual), activation, store the
    // layer output.  Same arithmetic as attn_csr_rows_kernel, minus the (acc, stats) round trip through HBM.
    // tcgen05.ld / wait are warp-collective, so those are executed by every lane and only the per-row work
    // is predicated.
    constexpr int RM = DA_FUSE_MAX_RESIDUAL;
    int src[RM]; float pe[RM], wgt[RM], d[RM];
    float o_scale = 0.f;
    if (fused) {
      const int beg = res_beg;
#pragma unroll
      for (int e = 0; e < RM; ++e) {
        src[e] = (e < n_e) ? a.col[beg + e] : node;
        wgt[e] = (e < n_e) ? (a.weight ? a.weight[beg + e] : 1.f) : 0.f;
        d[e] = 0.f;
      }
    }
    if (__any_sync(0xffffffffu, n_e > 0)) {
      // raw (unscaled) scores q . k_j in the units of m.  Q comes back from TMEM (this lane's row, split-bf16
      // hi + lo planes: the same operand the tensor-core scores used), K_j from the source's fp32 row.
      for (int c0 = 0; c0 < Cpad; c0 += 16) {   // 16 channels = 8 TMEM columns of packed pairs per plane
        uint32_t qh[8], ql[8];
        tmem_ld8(tmem_q + lane_off + (uint32_t)(c0 >> 1), qh);
        tmem_ld8(tmem_q + lane_off + (uint32_t)((Cpad >> 1) + (c0 >> 1)), ql);
        float4 kk[RM][4];
#pragma unroll
        for (int e = 0; e < RM; ++e) {
          if (e < n_e) {
            const float* krow = a.qkvs + (size_t)src[e] * a.ld + HC + head * a.C + c0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              kk[e][u] = (c0 + 4 * u < a.C) ? __ldg(reinterpret_cast<const float4*>(krow + 4 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tmem_ld_wait();
        float qf[16];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          qf[2 * w] = __uint_as_float(qh[w] << 16) + __uint_as_float(ql[w] << 16);
          qf[2 * w + 1] = __uint_as_float(qh[w] & 0xffff0000u) + __uint_as_float(ql[w] & 0xffff0000u);
        }
#pragma unroll
        for (int e = 0; e < RM; ++e) {
          if (e < n_e) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              d[e] = fmaf(qf[4 * u], kk[e][u].x, fmaf(qf[4 * u + 1], kk[e][u].y,
                     fmaf(qf[4 * u + 2], kk[e][u].z, fmaf(qf[4 * u + 3], kk[e][u].w, d[e]))));
          }
        }
      }
    }
    if (fused) {
      float m_fin = m;
#pragma unroll
      for (int e = 0; e < RM; ++e) if (e < n_e) m_fin = fmaxf(m_fin, d[e]);
      // m == -inf: no bitmap edge on this row (O and l are exactly 0); m_fin == -inf: no in-edge at all
      const float alpha = (m == -INFINITY) ? 0.f : ex2_approx((m - m_fin) * c_log2);
      float l_fin = l * alpha;
#pragma unroll
      for (int e = 0; e < RM; ++e) {
        pe[e] = (e < n_e) ? wgt[e] * ex2_approx((d[e] - m_fin) * c_log2) : 0.f;
        l_fin += pe[e];
      }
      const float inv = 1.f / (l_fin + 1e-16f);
      o_scale = alpha * inv;
#pragma unroll
      for (int e = 0; e < RM; ++e) pe[e] *= inv;
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64) a.dbg[122] = clock64();   // residual scores done
    if (a.row_fused != nullptr) mbar_wait(&sh->epi_full, 0);   // skip rows are in shared memory
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64) a.dbg[123] = clock64();   // staged rows landed
    const float* rsd = (a.resid && !stage_resid) ? a.resid + (size_t)node * a.ld_resid + head * a.C : nullptr;   // slow path
    // Full tiles leave through two TMA tensor stores (hi / lo planes of 128 rows x C bf16, staged in the two V stages
    // that are idle now); partial tiles (the last one of a graph) and fp32 outputs are stored by the threads.
    const bool tma_out = (stage_flags & 2) && a.out.hi != nullptr && ti.rows == TM;
    uint8_t* ohi_sm = v_sm + (size_t)((nblk + 1) % VST) * stage_b;
    uint8_t* olo_sm = v_sm + (size_t)((nblk + 2) % VST) * stage_b;
    const uint32_t orow_b = (uint32_t)a.C * 2u;   // bytes of one staged output row per plane
    // one 16-channel chunk of this row: v = the chunk of O (already in registers)
    auto chunk_body = [&](const int c0, const uint32_t (&v)[16]) {
      float4 vv[RM][4];   // V rows of the residual sources (rare once the planner has promoted them)
#pragma unroll
      for (int ee = 0; ee < RM; ++ee) {
        if (ee < n_e) {
          const float* vrow = a.qkvs + (size_t)src[ee] * a.ld + 2 * HC + head * a.C + c0;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            vv[ee][u] = (c0 + 4 * u < a.C) ? __ldg(reinterpret_cast<const float4*>(vrow + 4 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (fused) {
        float y[16];
        const int box = c0 >> 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 t = staged4(skip_reg, box, u);
          y[4 * u] = t.x; y[4 * u + 1] = t.y; y[4 * u + 2] = t.z; y[4 * u + 3] = t.w;
        }
        if (stage_resid) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 t = staged4(resid_reg, box, u);
            y[4 * u] += t.x; y[4 * u + 1] += t.y; y[4 * u + 2] += t.z; y[4 * u + 3] += t.w;
          }
        } else if (rsd) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c0 + 4 * u < a.C) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(rsd + c0 + 4 * u));
              y[4 * u] += t.x; y[4 * u + 1] += t.y; y[4 * u + 2] += t.z; y[4 * u + 3] += t.w;
            }
          }
        }
        // the reference adds (attention + skip) + resid; here the attention term joins the pre-summed rest
        if (n_e > 0) {
          float att[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) att[e] = __uint_as_float(v[e]) * o_scale;
#pragma unroll
          for (int ee = 0; ee < RM; ++ee) {
            if (ee < n_e) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                att[4 * u] = fmaf(pe[ee], vv[ee][u].x, att[4 * u]); att[4 * u + 1] = fmaf(pe[ee], vv[ee][u].y, att[4 * u + 1]);
                att[4 * u + 2] = fmaf(pe[ee], vv[ee][u].z, att[4 * u + 2]); att[4 * u + 3] = fmaf(pe[ee], vv[ee][u].w, att[4 * u + 3]);
              }
            }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) y[e] += att[e];
        } else {   // the common case after promotion: one FMA per channel
#pragma unroll
          for (int e = 0; e < 16; ++e) y[e] = fmaf(__uint_as_float(v[e]), o_scale, y[e]);
        }
        if (a.act != ACT_NONE) {
#pragma unroll
          for (int e = 0; e < 16; ++e) y[e] = apply_act_rt(y[e], a.act);
        }
        const int nval = (a.C - c0) < 16 ? (a.C - c0) : 16;   // valid channels of this chunk (> 0: Cpad - C < 16)
        if (a.out.f32) {
          float* dst = a.out.f32 + (size_t)node * a.out.ldc + head * a.C + c0;
          if (nval == 16) {
#pragma unroll
            for (int e4 = 0; e4 < 16; e4 += 4)
              *reinterpret_cast<float4*>(dst + e4) = make_float4(y[e4], y[e4 + 1], y[e4 + 2], y[e4 + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (e < nval) dst[e] = y[e];
          }
        }
        if (a.out.hi) {
          uint32_t hh[8], ll[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const uint32_t h2 = pack_bf16x2(y[e], y[e + 1]);
            hh[e >> 1] = h2;
            ll[e >> 1] = pack_bf16x2(y[e] - __uint_as_float(h2 << 16), y[e + 1] - __uint_as_float(h2 & 0xffff0000u));
          }
          if (tma_out) {   // (ti.rows == 128 implies C == Cpad chunks are whole: a.C % 16 == 0 is checked on the host)
            uint4* dh = reinterpret_cast<uint4*>(ohi_sm + (size_t)r * orow_b + (size_t)c0 * 2);
            uint4* dl = reinterpret_cast<uint4*>(olo_sm + (size_t)r * orow_b + (size_t)c0 * 2);
            dh[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]); dh[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
            dl[0] = make_uint4(ll[0], ll[1], ll[2], ll[3]); dl[1] = make_uint4(ll[4], ll[5], ll[6], ll[7]);
          } else {
            __nv_bfloat16* dh = a.out.hi + (size_t)node * a.out.ld_split + head * a.C + c0;
            __nv_bfloat16* dl = a.out.lo + (size_t)node * a.out.ld_split + head * a.C + c0;
            if (nval == 16 && (a.C & 7) == 0 && (a.out.ld_split & 7) == 0) {
              *reinterpret_cast<uint4*>(dh) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
              *reinterpret_cast<uint4*>(dh + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
              *reinterpret_cast<uint4*>(dl) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
              *reinterpret_cast<uint4*>(dl + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {   // compile-time indices: hh / ll stay in registers
                if (e < nval) {
                  const uint32_t hw = hh[e >> 1], lw = ll[e >> 1];
                  reinterpret_cast<unsigned short*>(dh)[e] = (e & 1) ? (unsigned short)(hw >> 16) : (unsigned short)(hw & 0xffffu);
                  reinterpret_cast<unsigned short*>(dl)[e] = (e & 1) ? (unsigned short)(lw >> 16) : (unsigned short)(lw & 0xffffu);
                }
              }
            }
          }
        }
      } else {
        // un-normalised O to global; attn_csr.cu continues over the residual edges
        if (row_valid) {
          float* dst = a.acc + (size_t)node * HC + head * a.C + c0;
          if (c0 + 16 <= a.C && (a.C & 3) == 0) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              *reinterpret_cast<float4*>(dst + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                 __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
          } else {
            for (int e = 0; e < 16; ++e)
              if (c0 + e < a.C) dst[e] = __uint_as_float(v[e]);
          }
        }
      }
    };
    // tcgen05.ld / wait are warp-collective: issued by every lane, outside the per-row branches of chunk_body.
    // (Software-pipelining the TMEM loads across chunks was measured slower: the unrolled bodies cost more in
    // instruction fetch and registers than the ~100-cycle TMEM latency they hide.)
    for (int c0 = 0; c0 < Cpad; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_o + lane_off + c0, v);
      tmem_ld_wait();
      chunk_body(c0, v);
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64) a.dbg[125] = clock64();   // chunk loop done
    if (tma_out) {
      // generic-proxy writes -> async proxy, all 128 epilogue threads (named barrier 1), then one thread stores both planes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2) {
        if (elect_one()) {
          tma_store_2d(&map_ohi, ohi_sm, head * a.C, ti.node0);
          tma_store_2d(&map_olo, olo_sm, head * a.C, ti.node0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the reads
        }
        __syncwarp();
      }
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 64) a.dbg[124] = clock64();   // outputs stored
    if (row_valid && !fused) {
      float* st = a.stats + ((size_t)node * a.H + head) * 2;
      st[0] = (m == -INFINITY) ? -INFINITY : m / sqrtf((float)a.C);  // natural-log units of the scaled score
      st[1] = l;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols));
  }
  if (dbg_rec) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.dbg[128 + 8 * blockIdx.x + 1] = (long long)gt;
    a.dbg[128 + 8 * blockIdx.x + 7] = clock64();
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Folded last layer (fold.cu): S = Q K^T over the C-channel images as above, but the aggregated values are the
// 32-channel V' = x (W_a^h W_v^h)^T of final_mlp[0] folded into lin_value, and there is no skip / residual / activation:
// the epilogue is "normalise and store 32 floats per (row, head)".  P V' costs 12 MMAs of N = 32 per block instead of
// 12 of N = 144, the V ring shrinks to 8 KB per stage (ring depth 4 at C = 144), O to 32 TMEM columns.
// Same roles, barriers and softmax loop as attn_dense_kernel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FV = 32;   // channels of V' per head = width of final_mlp[0]

template <int CQ_T, int ST>
__global__ void __launch_bounds__(NT)
attn_dense_fold_kernel(AttnFoldArgs a, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int Cq = CQ_T ? CQ_T : a.Cpad;
  const uint32_t k_plane = TS * Cq * 2, v_plane = TS * FV * 2;   // bytes
  uint8_t* k_sm = smem;
  uint8_t* v_sm = k_sm + ST * 2 * k_plane;
  DenseSmem* sh = reinterpret_cast<DenseSmem*>(v_sm + ST * 2 * v_plane);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / a.H, head = blockIdx.x % a.H;
  const TileInfo ti = a.tiles[tile];
  const int nblk = ti.n_list;
  const uint16_t* __restrict__ blist = a.blk_list + ti.list_off;

  if (threadIdx.x == 0) {
    mbar_init(&sh->q_full, 4);
    for (int i = 0; i < MAXST; ++i) {
      mbar_init(&sh->k_full[i], 1); mbar_init(&sh->k_empty[i], 1);
      mbar_init(&sh->v_full[i], 1); mbar_init(&sh->v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh->s_full[i], 1); mbar_init(&sh->p_full[i], 4); mbar_init(&sh->pv_done[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"((uint32_t)tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  const uint32_t tmem_s = tmem_base;             // 2 x TS columns: S_j (fp32), later P_j (bf16 hi | lo)
  const uint32_t tmem_o = tmem_base + 2 * TS;    // FV columns
  const uint32_t tmem_q = tmem_o + FV;           // Cq columns: Q as packed bf16 pairs, hi plane then lo plane

  if (warp == 0) {  // ===== bulk-copy producer =====
    auto blk_idx = [&](int j) { return (size_t)(ti.gblock0 + (int)(__ldg(blist + j) >> 1)) * a.H + head; };
    auto load_k = [&](int j) {
      if (elect_one()) {
        const int st = j % ST;
        mbar_expect_tx(&sh->k_full[st], 2 * k_plane);
        bulk_load(k_sm + st * 2 * k_plane, a.kimg + blk_idx(j) * kv_block_elems(Cq), 2 * k_plane, &sh->k_full[st]);
      }
      __syncwarp();
    };
    auto load_v = [&](int j) {
      if (elect_one()) {
        const int st = j % ST;
        mbar_expect_tx(&sh->v_full[st], 2 * v_plane);
        bulk_load(v_sm + st * 2 * v_plane, a.vimg + blk_idx(j) * kv_block_elems(FV), 2 * v_plane, &sh->v_full[st]);
      }
      __syncwarp();
    };
    for (int j = 0; j < ST && j < nblk; ++j) load_k(j);
    for (int j = 0; j < ST && j < nblk; ++j) load_v(j);
    for (int j = 0; j < nblk; ++j) {
      if (j + ST < nblk) {
        mbar_wait(&sh->k_empty[j % ST], (j / ST) & 1);
        load_k(j + ST);
        mbar_wait(&sh->v_empty[j % ST], (j / ST) & 1);
        load_v(j + ST);
      }
    }
  } else if (warp == 1) {  // ===== MMA issuer =====
    const uint32_t idesc_s = make_idesc(TM, TS), idesc_o = make_idesc(TM, FV) | (1u << 16);  // bit 16: B is MN-major
    const int ksteps = Cq / 16;
    const uint32_t tq_hi = tmem_q, tq_lo = tmem_q + Cq / 2;
    const uint64_t dk0 = make_desc_nosw(smem_u32(k_sm), TS * 16, 128);
    const uint64_t dv0 = make_desc_nosw(smem_u32(v_sm), 128, TS * 16);
    const uint32_t kstage_u = (2 * k_plane) >> 4, kplane_u = k_plane >> 4;
    const uint32_t vstage_u = (2 * v_plane) >> 4, vplane_u = v_plane >> 4;
    auto issue_s = [&](int j) {
      if (elect_one()) {
        const uint64_t dk_hi = dk0 + (uint32_t)(j % ST) * kstage_u, dk_lo = dk_hi + kplane_u;
        const uint32_t d = tmem_s + (uint32_t)((j & 1) * TS);
#pragma unroll
        for (int kk = 0; kk < (CQ_T ? CQ_T / 16 : 16); ++kk) {
          if (!CQ_T && kk >= ksteps) break;
          const uint32_t ko = (uint32_t)kk * ((2 * TS * 16) >> 4);
          tc_mma_bf16_ts(d, tq_hi + kk * 8, dk_hi + ko, idesc_s, kk ? 1u : 0u);
          tc_mma_bf16_ts(d, tq_hi + kk * 8, dk_lo + ko, idesc_s, 1u);
          tc_mma_bf16_ts(d, tq_lo + kk * 8, dk_hi + ko, idesc_s, 1u);
        }
        tc_commit(&sh->s_full[j & 1]);
        tc_commit(&sh->k_empty[j % ST]);
      }
      __syncwarp();
    };
    mbar_wait(&sh->q_full, 0);
    mbar_wait(&sh->k_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nblk; ++j) {
      if (j + 1 < nblk) {
        const int jn = j + 1;
        mbar_wait(&sh->k_full[jn % ST], (jn / ST) & 1);
        if (jn >= 2) mbar_wait(&sh->pv_done[jn & 1], ((jn >> 1) - 1) & 1);
        tc_fence_after();
        issue_s(jn);
      }
      const int b = j & 1, vs = j % ST;
      mbar_wait(&sh->p_full[b], (j >> 1) & 1);
      mbar_wait(&sh->v_full[vs], (j / ST) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_hi = tmem_s + (uint32_t)(b * TS), p_lo = p_hi + TS / 2;
        const uint64_t dv_hi = dv0 + (uint32_t)vs * vstage_u, dv_lo = dv_hi + vplane_u;
#pragma unroll
        for (int kk = 0; kk < TS / 16; ++kk) {
          tc_mma_bf16_ts(tmem_o, p_hi + kk * 8, dv_hi + kk * 16, idesc_o, (j | kk) ? 1u : 0u);
          tc_mma_bf16_ts(tmem_o, p_hi + kk * 8, dv_lo + kk * 16, idesc_o, 1u);
          tc_mma_bf16_ts(tmem_o, p_lo + kk * 8, dv_hi + kk * 16, idesc_o, 1u);
        }
        tc_commit(&sh->pv_done[b]);
        tc_commit(&sh->v_empty[vs]);
      }
      __syncwarp();
    }
  } else {  // ===== softmax warps =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool row_valid = r < ti.rows;
    const float c_log2 = 1.4426950408889634f / sqrtf((float)a.C);
    const float tau_raw = LAZY_LOG2 / c_log2;
    const uint32_t* bm_row = a.bitmap + ti.bm_off + (size_t)(ti.row0 + r) * ti.bm_words;
    {  // park this row's Q in TMEM
      const uint4* qsrc = reinterpret_cast<const uint4*>(a.qimg + ((size_t)tile * a.H + head) * q_block_elems(Cq));
      for (int pl_ = 0; pl_ < 2; ++pl_) {
        for (int ch0 = 0; ch0 < Cq / 8; ch0 += 4) {
          uint32_t qv[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (ch0 + u < Cq / 8) t = __ldg(qsrc + (size_t)pl_ * (TM * Cq / 8) + (size_t)(ch0 + u) * TM + r);
            qv[4 * u] = t.x; qv[4 * u + 1] = t.y; qv[4 * u + 2] = t.z; qv[4 * u + 3] = t.w;
          }
          const uint32_t dst = tmem_q + lane_off + (uint32_t)(pl_ * (Cq / 2) + ch0 * 4);
          if (ch0 + 4 <= Cq / 8) tmem_st16(dst, qv);
          else {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(dst), "r"(qv[0]),
                         "r"(qv[1]), "r"(qv[2]), "r"(qv[3]), "r"(qv[4]), "r"(qv[5]), "r"(qv[6]), "r"(qv[7]) : "memory");
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->q_full);
    }
    float m = -INFINITY, l = 0.f;
    uint2 bits_next = row_valid ? *reinterpret_cast<const uint2*>(bm_row + (int)(__ldg(blist) >> 1) * 2) : make_uint2(0u, 0u);
    for (int j = 0; j < nblk; ++j) {
      const int b = j & 1;
      const uint2 bits = bits_next;
      if (j + 1 < nblk && row_valid) bits_next = *reinterpret_cast<const uint2*>(bm_row + (int)(__ldg(blist + j + 1) >> 1) * 2);
      mbar_wait(&sh->s_full[b], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_s + lane_off + (uint32_t)(b * TS);
      if (j == 0) {
        uint32_t v[TS];
#pragma unroll
        for (int c0 = 0; c0 < TS; c0 += 16) tmem_ld16(s_addr + c0, v + c0);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < TS; ++e) {
          const uint32_t w = (e < 32) ? bits.x : bits.y;
          if ((w >> (e & 31)) & 1u) m = fmaxf(m, __uint_as_float(v[e]));
        }
      }
      bool prev_done = (j == 0);
      uint32_t ph[TS / 2], pl[TS / 2];
      float lsum, bmax;
      while (true) {
        const float m_sub = (m == -INFINITY) ? 0.f : m * c_log2;
        lsum = 0.f; bmax = -INFINITY;
        uint32_t v[TS];
#pragma unroll
        for (int c0 = 0; c0 < TS; c0 += 16) tmem_ld16(s_addr + c0, v + c0);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < TS; e += 2) {
          const uint32_t w = (e < 32) ? bits.x : bits.y;
          const float s0 = ((w >> (e & 31)) & 1u) ? __uint_as_float(v[e]) : -INFINITY;
          const float s1 = ((w >> ((e + 1) & 31)) & 1u) ? __uint_as_float(v[e + 1]) : -INFINITY;
          bmax = fmaxf(bmax, fmaxf(s0, s1));
          const float p0 = ex2_approx(fmaf(s0, c_log2, -m_sub));
          const float p1 = ex2_approx(fmaf(s1, c_log2, -m_sub));
          lsum += p0 + p1;
          const uint32_t h2 = pack_bf16x2(p0, p1);
          ph[e >> 1] = h2;
          pl[e >> 1] = pack_bf16x2(p0 - __uint_as_float(h2 << 16), p1 - __uint_as_float(h2 & 0xffff0000u));
        }
        const bool exceeded = bmax > m + tau_raw;
        if (!__any_sync(0xffffffffu, exceeded)) break;
        const float m_new = exceeded ? bmax : m;
        const float alpha = (m == -INFINITY) ? 0.f : ex2_approx((m - m_new) * c_log2);
        if (!prev_done) {
          mbar_wait(&sh->pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          prev_done = true;
        }
        if (j > 0) {
#pragma unroll
          for (int c0 = 0; c0 < FV; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_o + lane_off + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st16(tmem_o + lane_off + c0, o);
          }
          tmem_st_wait();
        }
        l *= alpha;
        m = m_new;
      }
      l += lsum;
      tmem_st16(s_addr, ph);
      tmem_st16(s_addr + 16, ph + 16);
      tmem_st16(s_addr + 32, pl);
      tmem_st16(s_addr + 48, pl + 16);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->p_full[b]);
    }
    mbar_wait(&sh->pv_done[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);
    tc_fence_after();
    // epilogue: partial[node, head, :] = O / (l + 1e-16)   (rows without an in-edge: O = 0 exactly)
    uint32_t o[FV];
    tmem_ld16(tmem_o + lane_off, o);
    tmem_ld16(tmem_o + lane_off + 16, o + 16);
    tmem_ld_wait();
    if (row_valid) {
      const float inv = 1.f / (l + 1e-16f);
      float4* dst = reinterpret_cast<float4*>(a.partial + ((size_t)(ti.node0 + r) * a.H + head) * FV);
#pragma unroll
      for (int e = 0; e < FV; e += 4)
        dst[e >> 2] = make_float4(__uint_as_float(o[e]) * inv, __uint_as_float(o[e + 1]) * inv, __uint_as_float(o[e + 2]) * inv,
                                  __uint_as_float(o[e + 3]) * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent form of the folded last layer: one CTA per SM walks (tile, head) items w = blockIdx.x, + gridDim.x, ...
// With a 73 KB Q tile and a 180 KB K / V ring only one CTA fits an SM, so in the one-item-per-CTA kernel above nothing
// hides an item's fixed cost (TMEM allocation, barrier set-up, the Q tile's trip from global memory, the first K block,
// the epilogue): ~1/3 of its run time.  Here the pipeline never drains:
//   warp 0      K / V producer, ring counters run across items (the next item's blocks are requested while this one computes)
//   warp 1      MMA issue; S of the NEXT item's first block is issued before the last P V' of the current one
//   warps 2-5   softmax (as above), hands the row sums l to the epilogue warps through shared memory
//   warps 6-9   park the next item's Q tile in the second TMEM Q buffer, then finalise the current item:
//               O (double-buffered in TMEM) / l -> partial[node, head, :]
// TMEM: S/P 2 x 64 | O 2 x 32 (hi / lo halves of ONE buffer) | Q 2 x Cq columns (480 of 512 at Cq = 144).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NTP = 320;

struct FoldSmem {
  uint64_t k_full[MAXST], k_empty[MAXST], v_full[MAXST], v_empty[MAXST];
  uint64_t s_full[2], p_full[2], pv_done[2];
  uint64_t q_full[2];    // Q-park warps -> MMA: item's Q tile is in TMEM buffer (item & 1)
  uint64_t q_free[2];    // MMA -> Q-park warps: every S of the item that used the buffer has retired
  uint64_t o_full[2];    // MMA -> epilogue: every P V' of the item has retired (O buffer item & 1)
  uint64_t o_empty[2];   // epilogue -> MMA / softmax: O buffer and l buffer (item & 1) have been read
  uint64_t l_full[2];    // softmax -> epilogue: the item's row sums are in lbuf[item & 1]
  uint32_t tmem_base;
  float lbuf[2][TM];
};

template <int CQ_T, int ST>
__global__ void __launch_bounds__(NTP, 1)
attn_fold_persist_kernel(AttnFoldArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  const int Cq = CQ_T ? CQ_T : a.Cpad;
  const uint32_t k_plane = TS * Cq * 2, v_plane = TS * FV * 2;   // bytes
  uint8_t* k_sm = smem;
  uint8_t* v_sm = k_sm + ST * 2 * k_plane;
  FoldSmem* sh = reinterpret_cast<FoldSmem*>(v_sm + ST * 2 * v_plane);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = a.n_tiles * a.H;
  const int w0 = blockIdx.x, wstep = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAXST; ++i) {
      mbar_init(&sh->k_full[i], 1); mbar_init(&sh->k_empty[i], 1);
      mbar_init(&sh->v_full[i], 1); mbar_init(&sh->v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh->s_full[i], 1); mbar_init(&sh->p_full[i], 4); mbar_init(&sh->pv_done[i], 1);
      mbar_init(&sh->q_full[i], 4); mbar_init(&sh->q_free[i], 1);
      mbar_init(&sh->o_full[i], 1); mbar_init(&sh->o_empty[i], 4); mbar_init(&sh->l_full[i], 4);
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
  const uint32_t tmem_s = tmem_base;                  // 2 x TS columns
  const uint32_t tmem_o0 = tmem_base + 2 * TS;        // 2 x FV columns: O_a = P_hi V'_hi + P_lo V'_hi | O_b = P_hi V'_lo (one buffer)
  const uint32_t tmem_q0 = tmem_o0 + 2 * FV;          // 2 x Cq columns
  pdl_wait();   // set-up above runs under the previous kernel's tail

  if (warp == 0) {  // ===== bulk-copy producer: running block counter kc over all items =====
    int kc = 0;
    for (int w = w0; w < n_items; w += wstep) {
      const int tile = w / a.H, head = w - tile * a.H;
      const TileInfo ti = a.tiles[tile];
      const uint16_t* __restrict__ blist = a.blk_list + ti.list_off;
      for (int j = 0; j < ti.n_list; ++j, ++kc) {
        const int st = kc % ST;
        const uint32_t par = (uint32_t)((kc / ST) - 1) & 1u;
        const size_t blk = (size_t)(ti.gblock0 + (int)(__ldg(blist + j) >> 1)) * a.H + head;
        if (kc >= ST) mbar_wait(&sh->k_empty[st], par);
        if (elect_one()) {
          mbar_expect_tx(&sh->k_full[st], 2 * k_plane);
          bulk_load(k_sm + st * 2 * k_plane, a.kimg + blk * kv_block_elems(Cq), 2 * k_plane, &sh->k_full[st]);
        }
        __syncwarp();
        if (kc >= ST) mbar_wait(&sh->v_empty[st], par);
        if (elect_one()) {
          mbar_expect_tx(&sh->v_full[st], 2 * v_plane);
          bulk_load(v_sm + st * 2 * v_plane, a.vimg + blk * kv_block_elems(FV), 2 * v_plane, &sh->v_full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {  // ===== MMA issuer =====
    // P_hi [V'_hi | V'_lo] is ONE N = 64 instruction (the lo plane continues the hi plane's chunk sequence in the stage, and
    // an MMA of this shape costs ~45 cycles for any N <= 64: profiles/r2_mma_issue_rate_b200.txt): 8 instead of 12 per block
    const uint32_t idesc_s = make_idesc(TM, TS);
    const uint32_t idesc_o2 = make_idesc(TM, 2 * FV) | (1u << 16), idesc_o1 = make_idesc(TM, FV) | (1u << 16);
    static_assert(TS * FV * 2 == 4 * TS * 16, "the lo plane must continue the hi plane's chunk sequence");
    const int ksteps = Cq / 16;
    const uint64_t dk0 = make_desc_nosw(smem_u32(k_sm), TS * 16, 128);
    const uint64_t dv0 = make_desc_nosw(smem_u32(v_sm), 128, TS * 16);
    const uint32_t kstage_u = (2 * k_plane) >> 4, kplane_u = k_plane >> 4;
    const uint32_t vstage_u = (2 * v_plane) >> 4, vplane_u = v_plane >> 4;
    // S of running block gs from the Q buffer qb; last_of_item: also tell the Q-park warps that the buffer is free
    auto issue_s = [&](int gs, int qb, bool last_of_item) {
      mbar_wait(&sh->k_full[gs % ST], (uint32_t)(gs / ST) & 1u);
      if (gs >= 2) mbar_wait(&sh->pv_done[gs & 1], (uint32_t)((gs >> 1) - 1) & 1u);   // P of block gs - 2 consumed
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tq_hi = tmem_q0 + (uint32_t)(qb * Cq), tq_lo = tq_hi + Cq / 2;
        const uint64_t dk_hi = dk0 + (uint32_t)(gs % ST) * kstage_u, dk_lo = dk_hi + kplane_u;
        const uint32_t d = tmem_s + (uint32_t)((gs & 1) * TS);
#pragma unroll
        for (int kk = 0; kk < (CQ_T ? CQ_T / 16 : 16); ++kk) {
          if (!CQ_T && kk >= ksteps) break;
          const uint32_t ko = (uint32_t)kk * ((2 * TS * 16) >> 4);
          tc_mma_bf16_ts(d, tq_hi + kk * 8, dk_hi + ko, idesc_s, kk ? 1u : 0u);
          tc_mma_bf16_ts(d, tq_hi + kk * 8, dk_lo + ko, idesc_s, 1u);
          tc_mma_bf16_ts(d, tq_lo + kk * 8, dk_hi + ko, idesc_s, 1u);
        }
        tc_commit(&sh->s_full[gs & 1]);
        tc_commit(&sh->k_empty[gs % ST]);
        if (last_of_item) tc_commit(&sh->q_free[qb]);
      }
      __syncwarp();
    };
    if (w0 < n_items) {
      int g = 0, it = 0, w = w0;
      int nblk_cur = a.tiles[w / a.H].n_list;
      mbar_wait(&sh->q_full[0], 0);
      tc_fence_after();
      issue_s(0, 0, nblk_cur == 1);
      while (true) {
        const int w_next = w + wstep;
        const bool has_next = w_next < n_items;
        const int nblk_next = has_next ? a.tiles[w_next / a.H].n_list : 0;
        const int ob = it & 1;
        for (int j = 0; j < nblk_cur; ++j, ++g) {
          if (j + 1 < nblk_cur) {
            issue_s(g + 1, it & 1, j + 2 == nblk_cur);
          } else if (has_next) {   // the next item's first S goes in front of this item's last P V'
            mbar_wait(&sh->q_full[(it + 1) & 1], (uint32_t)((it + 1) >> 1) & 1u);
            tc_fence_after();
            issue_s(g + 1, (it + 1) & 1, nblk_next == 1);
          }
          const int b = g & 1, vs = g % ST;
          mbar_wait(&sh->p_full[b], (uint32_t)(g >> 1) & 1u);
          mbar_wait(&sh->v_full[vs], (uint32_t)(g / ST) & 1u);
          if (j == 0 && it >= 1) mbar_wait(&sh->o_empty[(it - 1) & 1], (uint32_t)((it - 1) >> 1) & 1u);   // epilogue of item it - 1 has read O
          tc_fence_after();
          if (elect_one()) {
            const uint32_t p_hi = tmem_s + (uint32_t)(b * TS), p_lo = p_hi + TS / 2;
            const uint32_t o_t = tmem_o0;
            const uint64_t dv_hi = dv0 + (uint32_t)vs * vstage_u;
#pragma unroll
            for (int kk = 0; kk < TS / 16; ++kk) {
              tc_mma_bf16_ts(o_t, p_hi + kk * 8, dv_hi + kk * 16, idesc_o2, (j | kk) ? 1u : 0u);
              tc_mma_bf16_ts(o_t, p_lo + kk * 8, dv_hi + kk * 16, idesc_o1, 1u);
            }
            tc_commit(&sh->pv_done[b]);
            tc_commit(&sh->v_empty[vs]);
            if (j == nblk_cur - 1) tc_commit(&sh->o_full[ob]);
          }
          __syncwarp();
        }
        if (!has_next) break;
        ++it; w = w_next; nblk_cur = nblk_next;
      }
    }
  } else if (warp < 6) {  // ===== softmax warps =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c_log2 = 1.4426950408889634f / sqrtf((float)a.C);
    const float tau_raw = LAZY_LOG2 / c_log2;
    int g = 0, it = 0;
    for (int w = w0; w < n_items; w += wstep, ++it) {
      const int tile = w / a.H;
      const TileInfo ti = a.tiles[tile];
      const int nblk = ti.n_list;
      const uint16_t* __restrict__ blist = a.blk_list + ti.list_off;
      const bool row_valid = r < ti.rows;
      const uint32_t* bm_row = a.bitmap + ti.bm_off + (size_t)(ti.row0 + r) * ti.bm_words;
      const uint32_t o_t = tmem_o0 + lane_off;
      float m = -INFINITY, l = 0.f;
      uint2 bits_next = row_valid ? *reinterpret_cast<const uint2*>(bm_row + (int)(__ldg(blist) >> 1) * 2) : make_uint2(0u, 0u);
      for (int j = 0; j < nblk; ++j, ++g) {
        const int b = g & 1;
        const uint2 bits = bits_next;
        if (j + 1 < nblk && row_valid) bits_next = *reinterpret_cast<const uint2*>(bm_row + (int)(__ldg(blist + j + 1) >> 1) * 2);
        mbar_wait(&sh->s_full[b], (uint32_t)(g >> 1) & 1u);
        tc_fence_after();
        const uint32_t s_addr = tmem_s + lane_off + (uint32_t)(b * TS);
        if (j == 0) {
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
        bool prev_done = (j == 0);
        uint32_t ph[TS / 2], pl[TS / 2];
        float lsum, bmax;
        while (true) {
          const float m_sub = (m == -INFINITY) ? 0.f : m * c_log2;
          lsum = 0.f; bmax = -INFINITY;
          uint32_t v[TS];
#pragma unroll
          for (int c0 = 0; c0 < TS; c0 += 16) tmem_ld16(s_addr + c0, v + c0);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < TS; e += 2) {
            const uint32_t wd = (e < 32) ? bits.x : bits.y;
            const float s0 = ((wd >> (e & 31)) & 1u) ? __uint_as_float(v[e]) : -INFINITY;
            const float s1 = ((wd >> ((e + 1) & 31)) & 1u) ? __uint_as_float(v[e + 1]) : -INFINITY;
            bmax = fmaxf(bmax, fmaxf(s0, s1));
            const float p0 = ex2_approx(fmaf(s0, c_log2, -m_sub));
            const float p1 = ex2_approx(fmaf(s1, c_log2, -m_sub));
            lsum += p0 + p1;
            const uint32_t h2 = pack_bf16x2(p0, p1);
            ph[e >> 1] = h2;
            pl[e >> 1] = pack_bf16x2(p0 - __uint_as_float(h2 << 16), p1 - __uint_as_float(h2 & 0xffff0000u));
          }
          const bool exceeded = bmax > m + tau_raw;
          if (!__any_sync(0xffffffffu, exceeded)) break;
          const float m_new = exceeded ? bmax : m;
          const float alpha = (m == -INFINITY) ? 0.f : ex2_approx((m - m_new) * c_log2);
          if (!prev_done) {
            mbar_wait(&sh->pv_done[(g - 1) & 1], (uint32_t)((g - 1) >> 1) & 1u);   // O holds every earlier block of this item
            tc_fence_after();
            prev_done = true;
          }
          if (j > 0) {
#pragma unroll
            for (int c0 = 0; c0 < 2 * FV; c0 += 16) {
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
        tmem_st16(s_addr, ph);
        tmem_st16(s_addr + 16, ph + 16);
        tmem_st16(s_addr + 32, pl);
        tmem_st16(s_addr + 48, pl + 16);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->p_full[b]);
      }
      // hand the row sums to the epilogue warps
      if (it >= 2) mbar_wait(&sh->o_empty[it & 1], (uint32_t)((it >> 1) - 1) & 1u);   // lbuf[it & 1] of item it - 2 has been read
      sh->lbuf[it & 1][r] = l;
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->l_full[it & 1]);
    }
  } else {  // ===== Q-park + epilogue warps =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    auto park = [&](int itn, int wn) {   // Q tile of item itn (work index wn) -> TMEM Q buffer itn & 1
      const int qb = itn & 1;
      if (itn >= 2) { mbar_wait(&sh->q_free[qb], (uint32_t)((itn >> 1) - 1) & 1u); tc_fence_after(); }
      const int tile = wn / a.H, head = wn - tile * a.H;
      const uint4* qsrc = reinterpret_cast<const uint4*>(a.qimg + ((size_t)tile * a.H + head) * q_block_elems(Cq));
      const uint32_t tq = tmem_q0 + (uint32_t)(qb * Cq) + lane_off;
      for (int pl_ = 0; pl_ < 2; ++pl_) {
        for (int ch0 = 0; ch0 < Cq / 8; ch0 += 4) {
          uint32_t qv[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (ch0 + u < Cq / 8) t = __ldg(qsrc + (size_t)pl_ * (TM * Cq / 8) + (size_t)(ch0 + u) * TM + r);
            qv[4 * u] = t.x; qv[4 * u + 1] = t.y; qv[4 * u + 2] = t.z; qv[4 * u + 3] = t.w;
          }
          const uint32_t dst = tq + (uint32_t)(pl_ * (Cq / 2) + ch0 * 4);
          if (ch0 + 4 <= Cq / 8) tmem_st16(dst, qv);
          else {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(dst), "r"(qv[0]),
                         "r"(qv[1]), "r"(qv[2]), "r"(qv[3]), "r"(qv[4]), "r"(qv[5]), "r"(qv[6]), "r"(qv[7]) : "memory");
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->q_full[qb]);
    };
    if (w0 < n_items) park(0, w0);
    int it = 0;
    for (int w = w0; w < n_items; w += wstep, ++it) {
      if (w + wstep < n_items) park(it + 1, w + wstep);
      const int tile = w / a.H, head = w - tile * a.H;
      const TileInfo ti = a.tiles[tile];
      const int ob = it & 1;
      mbar_wait(&sh->l_full[ob], (uint32_t)(it >> 1) & 1u);
      mbar_wait(&sh->o_full[ob], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      uint32_t o[FV], o2[FV];
      tmem_ld16(tmem_o0 + lane_off, o);
      tmem_ld16(tmem_o0 + lane_off + 16, o + 16);
      tmem_ld16(tmem_o0 + lane_off + 32, o2);
      tmem_ld16(tmem_o0 + lane_off + 48, o2 + 16);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < FV; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) + __uint_as_float(o2[e]));
      const float l = sh->lbuf[ob][r];
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->o_empty[ob]);
      if (r < ti.rows) {   // partial[node, head, :] = O / (l + 1e-16)   (rows without an in-edge: O = 0 exactly)
        const float inv = 1.f / (l + 1e-16f);
        float4* dst = reinterpret_cast<float4*>(a.partial + ((size_t)(ti.node0 + r) * a.H + head) * FV);
#pragma unroll
        for (int e = 0; e < FV; e += 4)
          dst[e >> 2] = make_float4(__uint_as_float(o[e]) * inv, __uint_as_float(o[e + 1]) * inv, __uint_as_float(o[e + 2]) * inv,
                                    __uint_as_float(o[e + 3]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

}  // namespace

cudaError_t launch_pack_images(const PackArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  if ((a.C % 8) || (a.ld % 4)) return cudaErrorInvalidValue;
  if (a.Cpad % 16 || a.Cpad < a.C) return cudaErrorInvalidValue;
  const int groups = 3 * a.H * a.Cpad / 8;
  int gy = (groups + 8 * 8 - 1) / (8 * 8);  // ~8 items per warp
  if (gy < 1) gy = 1;
  dim3 grid((a.n + 31) / 32, gy);
  pack_images_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_gather_extra(const PackArgs& a, const int32_t* x_src, const int32_t* x_slot, int n_extra, cudaStream_t s) {
  if (n_extra <= 0) return cudaSuccess;
  if ((a.C % 8) || (a.ld % 4) || a.Cpad % 16 || a.Cpad < a.C) return cudaErrorInvalidValue;
  if (a.Cv > 0 && ((a.Cv % 8) || a.Cvpad % 16 || a.Cvpad < a.Cv)) return cudaErrorInvalidValue;
  const long long threads = (long long)n_extra * 2 * 32;
  return launch_pdl(gather_extra_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, s, a, x_src, x_slot, n_extra);
}

namespace {
// TMEM columns, K / V ring depth and dynamic shared memory of the kernel for a padded head dim
bool dense_config(int Cpad, int* cols_out, int* st_out, size_t* smem_out) {
  if (Cpad % 16 || Cpad > 256 || Cpad < 16) return false;
  int need = 2 * TS + 2 * Cpad, cols = 32;   // S/P double buffer + O + Q
  while (cols < need) cols <<= 1;
  if (cols > 512) return false;
  auto smem_for = [&](int st) { return (size_t)(2 * st) * 2 * TS * Cpad * 2 + sizeof(DenseSmem) + 128; };
  const size_t limit = 227 * 1024;
  // K / V ring depth: 4 when two CTAs still fit per SM (small head dims), else 3, else 2
  int st = 4;
  if (cols <= 256 ? (2 * smem_for(4) + 2048 > limit) : (smem_for(4) > limit)) st = (smem_for(3) <= limit) ? 3 : 2;
  if (smem_for(st) > limit) return false;
  *cols_out = cols; *st_out = st; *smem_out = smem_for(st);
  return true;
}
}  // namespace

bool attn_dense_can_fuse(int C) {
  const int Cpad = (C + 15) / 16 * 16;
  int cols, st; size_t smem;
  if ((C & 3) || !dense_config(Cpad, &cols, &st, &smem)) return false;
  // 128 skip rows + 64 trunk-residual rows of C floats in the K ring (3 stages of 2 planes of 64 x Cpad bf16 hold
  // exactly that when C == Cpad), the other 64 residual rows in an idle V stage
  return st >= 3;
}

cudaError_t launch_attn_dense(const AttnDenseArgs& a, cudaStream_t s) {
  if (a.n_tiles <= 0) return cudaSuccess;
  const int Cpad = a.Cpad;
  int cols, st; size_t smem;
  if (!dense_config(Cpad, &cols, &st, &smem)) return cudaErrorInvalidValue;
  if (a.row_fused != nullptr && !attn_dense_can_fuse(a.C)) return cudaErrorInvalidValue;
  // skip (and trunk-residual) values are staged by TMA boxes of 64 rows x 16 floats (64-byte swizzle); full tiles
  // leave through TMA stores of 128 rows x C bf16 per plane
  CUtensorMap map_skip, map_resid, map_ohi, map_olo;
  memset(&map_skip, 0, sizeof(map_skip)); memset(&map_resid, 0, sizeof(map_resid));
  memset(&map_ohi, 0, sizeof(map_ohi)); memset(&map_olo, 0, sizeof(map_olo));
  int stage_flags = 0;
  if (a.row_fused != nullptr) {
    if (!get_tensor_map_2d(a.qkvs, 4, a.n_rows, a.ld, a.ld, 64, 16, 1, &map_skip)) return cudaErrorInvalidValue;
    if (a.resid != nullptr && get_tensor_map_2d(a.resid, 4, a.n_rows_resid, a.ld_resid, a.ld_resid, 64, 16, 1, &map_resid))
      stage_flags |= 1;
    if (a.out.hi != nullptr && a.C % 16 == 0 && a.C <= 256 && a.out.ld_split % 8 == 0 &&
        get_tensor_map_2d(a.out.hi, 2, a.n_rows_out, a.out.ld_split, a.out.ld_split, 128, a.C, 0, &map_ohi) &&
        get_tensor_map_2d(a.out.lo, 2, a.n_rows_out, a.out.ld_split, a.out.ld_split, 128, a.C, 0, &map_olo))
      stage_flags |= 2;
  }
  const unsigned grid = a.n_tiles * a.H;
#define DA_LAUNCH(CP, ST_)                                                                                          \
  do {                                                                                                              \
    static size_t smem_set = 0;                                                                                     \
    if (smem > smem_set) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(attn_dense_kernel<CP, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                               \
      smem_set = smem;                                                                                              \
    }                                                                                                               \
    attn_dense_kernel<CP, ST_><<<grid, NT, smem, s>>>(map_skip, map_resid, map_ohi, map_olo, a, cols, stage_flags);                                                     \
  } while (0)
  if (Cpad == 32 && st == 4) DA_LAUNCH(32, 4);
  else if (Cpad == 144 && st == 3) DA_LAUNCH(144, 3);
  else if (st == 4) DA_LAUNCH(0, 4);
  else if (st == 3) DA_LAUNCH(0, 3);
  else DA_LAUNCH(0, 2);
#undef DA_LAUNCH
  return cudaGetLastError();
}

namespace {
bool fold_config(int Cq, int* cols_out, int* st_out, size_t* smem_out) {
  if (Cq % 16 || Cq > 256 || Cq < 16) return false;
  int need = 2 * TS + FV + Cq, cols = 32;   // S/P double buffer + O + Q
  while (cols < need) cols <<= 1;
  if (cols > 512) return false;
  auto smem_for = [&](int st) { return (size_t)st * 2 * TS * (Cq + FV) * 2 + sizeof(DenseSmem) + 128; };
  const size_t limit = 227 * 1024;
  int st = 4;
  if (cols <= 256 ? (2 * smem_for(4) + 2048 > limit) : (smem_for(4) > limit)) st = (smem_for(3) <= limit) ? 3 : 2;
  if (smem_for(st) > limit) return false;
  *cols_out = cols; *st_out = st; *smem_out = smem_for(st);
  return true;
}
}  // namespace

bool attn_dense_fold_supported(int Cpad) {
  int cols, st; size_t smem;
  return fold_config(Cpad, &cols, &st, &smem);
}

cudaError_t launch_attn_dense_fold(const AttnFoldArgs& a, cudaStream_t s) {
  if (a.n_tiles <= 0) return cudaSuccess;
  int cols, st; size_t smem;
  if (!fold_config(a.Cpad, &cols, &st, &smem)) return cudaErrorInvalidValue;
  // persistent form (one CTA per SM over the (tile, head) items, Q / O double-buffered in TMEM): whenever both Q tiles fit
  if (a.persistent && 2 * TS + 2 * FV + 2 * a.Cpad <= 512) {
    auto smem_p = [&](int st_) { return (size_t)st_ * 2 * TS * (a.Cpad + FV) * 2 + sizeof(FoldSmem) + 128; };
    const int stp = smem_p(4) <= 227 * 1024 ? 4 : (smem_p(3) <= 227 * 1024 ? 3 : 2);
    const size_t smp = smem_p(stp);
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
    }
    const int items = a.n_tiles * a.H;
    const unsigned gridp = items < sms ? items : sms;
#define DA_LAUNCH_P(CP, ST_)                                                                                        \
    do {                                                                                                            \
      static size_t smem_set = 0;                                                                                   \
      if (smp > smem_set) {                                                                                         \
        cudaError_t e = cudaFuncSetAttribute(attn_fold_persist_kernel<CP, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smp); \
        if (e != cudaSuccess) return e;                                                                             \
        smem_set = smp;                                                                                             \
      }                                                                                                             \
      { cudaError_t e_ = launch_pdl(attn_fold_persist_kernel<CP, ST_>, dim3(gridp), dim3(NTP), smp, s, a); if (e_ != cudaSuccess) return e_; } \
    } while (0)
    if (a.Cpad == 144 && stp == 4) DA_LAUNCH_P(144, 4);
    else if (stp == 4) DA_LAUNCH_P(0, 4);
    else if (stp == 3) DA_LAUNCH_P(0, 3);
    else DA_LAUNCH_P(0, 2);
#undef DA_LAUNCH_P
    return cudaGetLastError();
  }
  const unsigned grid = a.n_tiles * a.H;
#define DA_LAUNCH_F(CP, ST_)                                                                                        \
  do {                                                                                                              \
    static size_t smem_set = 0;                                                                                     \
    if (smem > smem_set) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(attn_dense_fold_kernel<CP, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                               \
      smem_set = smem;                                                                                              \
    }                                                                                                               \
    attn_dense_fold_kernel<CP, ST_><<<grid, NT, smem, s>>>(a, cols);                                                \
  } while (0)
  if (a.Cpad == 144 && st == 4) DA_LAUNCH_F(144, 4);
  else if (st == 4) DA_LAUNCH_F(0, 4);
  else if (st == 3) DA_LAUNCH_F(0, 3);
  else DA_LAUNCH_F(0, 2);
#undef DA_LAUNCH_F
  return cudaGetLastError();
}

}  // namespace da
