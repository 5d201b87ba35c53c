// Tensor-core linear layer for sm_100a: y = act(x @ w^T + bias) with fp32-faithful results.
//
// Every nn.Linear on the hot path (efficient_gat.py:88-102, the four TransformerConv projections
// fused as one [Q|K|V|skip] GEMM per layer) runs here.  The reference computes them in fp32
// (torch 1.12, TF32 off); a single bf16 or TF32 pass misses the 1e-4 parity bar (2e-3 / 3e-4
// end to end), so each fp32 operand is carried as two bf16 planes (hi = bf16(x), lo = bf16(x - hi))
// and the product is evaluated as three tcgen05 MMAs per k-step
//        a_hi*w_hi + a_hi*w_lo + a_lo*w_hi          (fp32 accumulation in TMEM)
// which leaves a ~2^-17 relative error per product (5e-6 end to end, measured).
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor of the four operand planes into a 128B-swizzled
//               smem ring (full/empty mbarriers)
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit to the barriers
//   warps 2-5   epilogue: tcgen05.ld of the 128 x BN fp32 accumulator (double-buffered in TMEM so
//               the next tile's MMAs overlap it), + bias, activation, fp32 and/or split-bf16 stores
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstring>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "umma.cuh"

namespace da {
namespace {

constexpr int BM = 128;       // rows per tile = UMMA M
constexpr int BK = 64;        // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int EPI_WARPS = 8;
constexpr int EPI_PITCH = 32;   // floats per staged row; 16-byte chunks are XOR-swizzled by (row & 7) instead of padded

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in HW instead of spinning
      : "memory");
}
// One lane of a converged warp, chosen by the hardware.  Role warps keep warp-uniform control flow and
// only predicate the uniform-datapath instructions (TMA, tcgen05.mma / commit) with this: inside a plain
// `if (lane == 0)` region the compiler wraps EVERY such instruction in an elect/branch "waterfall" loop
// (~50 cycles per MMA issue), which starves the tensor pipe.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) | [32,46) SBO >> 4 =
//   1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::f16: c_format F32 (bit 4), a/b format BF16 (bits 7, 10),
// K-major A and B (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
struct Cfg {
  static constexpr int STAGE_BYTES = 2 * BM * BK * 2 + 2 * BN * BK * 2;   // a_hi, a_lo, w_hi, w_lo
  static constexpr int STAGES = (BN >= 256) ? 2 : ((BN >= 128) ? 3 : 4);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;            // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * 32 * EPI_PITCH * 4 /*epilogue staging*/;
};

struct EpiParams {
  const float* bias;
  float* cf; int ldc;
  __nv_bfloat16* chi; __nv_bfloat16* clo; int ldsp;
  int M, N, K, act;
  // fused operand-image output (see LinearOut)
  const int32_t* node_slot; __nv_bfloat16* qimg; __nv_bfloat16* kimg; __nv_bfloat16* vimg;
  int iH, iC, iCpad, irows;
  const uint8_t* f32_tile_flags;
};
// Mode-specific arguments travel as a separate kernel parameter: growing EpiParams itself changes ptxas' register
// allocation of the plain epilogue (168 -> 142 registers) and costs the store-bound GEMMs 30-50 %.
struct EpiNone { const void* unused; };
template <int MODE> struct EpiExtraT { using type = EpiNone; };
template <> struct EpiExtraT<2> { using type = HeadFinalArgs; };
struct EpiFold { int Cv, Cvpad; };   // MODE 3: [Q | K | V'] columns of the folded last layer (LinearOut::img_Cv)
template <> struct EpiExtraT<3> { using type = EpiFold; };

// MODE 0: plain epilogue; 2: fused 2-D head (BN == 32); 3: as 0 with the third image part V' in its own geometry.  Compile-time so that the plain, store-bound epilogue keeps its
// code shape.  (Measured and dropped: MODE 1, the per-layer gather of the promoted extra sources folded into this
// epilogue -- the extra code moved ptxas to a 144-register allocation of the whole epilogue and cost the QKVS GEMMs
// 15-50 %, far more than the four 7 us gather launches it saved.)
template <int BN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_umma_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                   const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                   EpiParams p, const typename EpiExtraT<MODE>::type ex) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2; // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = p.N / BN;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_ahi); tma_prefetch_desc(&map_alo); tma_prefetch_desc(&map_whi); tma_prefetch_desc(&map_wlo);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // whole warp: allocate TMEM columns, publish the base address through smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"((uint32_t)C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();   // barriers, TMEM and tensor maps are set up under the previous kernel's tail; its outputs are read from here on

  if (warp == 0) {  // ===== TMA producer (warp-uniform; one elected lane issues) =====
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + stage * C::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_2d(&map_ahi, &full_bar[stage], st, kb * BK, m0);
          tma_load_2d(&map_alo, &full_bar[stage], st + BM * BK * 2, kb * BK, m0);
          tma_load_2d(&map_whi, &full_bar[stage], st + 2 * BM * BK * 2, kb * BK, n0);
          tma_load_2d(&map_wlo, &full_bar[stage], st + 2 * BM * BK * 2 + BN * BK * 2, kb * BK, n0);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {  // ===== MMA issuer (warp-uniform; one elected lane issues) =====
    constexpr uint32_t idesc = make_idesc(BM, BN);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);          // TMA bytes have landed
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t da_hi = make_desc_sw128(st), da_lo = make_desc_sw128(st + BM * BK * 2);
          const uint64_t db_hi = make_desc_sw128(st + 2 * BM * BK * 2), db_lo = make_desc_sw128(st + 2 * BM * BK * 2 + BN * BK * 2);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t koff = (k * UMMA_K * 2) >> 4;  // 16-byte units along K inside the 128-byte swizzle row
            tc_mma_bf16(d_tmem, da_hi + koff, db_hi + koff, idesc, (kb | k) ? 1u : 0u);
            tc_mma_bf16(d_tmem, da_hi + koff, db_lo + koff, idesc, 1u);
            tc_mma_bf16(d_tmem, da_lo + koff, db_hi + koff, idesc, 1u);
          }
          tc_commit(&empty_bar[stage]);                // frees the smem slot when these MMAs retire
          if (kb == num_kb - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {  // ===== epilogue warps: TMEM -> registers -> smem (transpose) -> coalesced global stores =====
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // which half of the tile's columns this warp drains
    float* stage = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256) + (warp - 2) * (32 * EPI_PITCH);
    const int sub_row = lane >> 3;             // read-back mapping: 4 rows x 8 float4 per instruction
    const int sub_col = (lane & 7) * 4;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const int row_base = m0 + q * 32;
      int slots[4] = {-1, -1, -1, -1};
      int my_slot = -1;   // image row of THIS thread's accumulator row (direct path below)
      if (p.node_slot != nullptr) {   // issued before the accumulator wait so the latency is hidden
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = row_base + i * 8 + (lane >> 2);
          if (row < p.irows) slots[i] = __ldg(p.node_slot + row);
        }
        if (row_base + lane < p.irows) my_slot = __ldg(p.node_slot + row_base + lane);
      }
      const uint32_t tflags = p.f32_tile_flags ? (uint32_t)__ldg(p.f32_tile_flags + (m0 >> 7)) : 3u;   // BM == 128
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c0 = (BN >= 64 ? half * (BN / 2) : 0); c0 < (BN >= 64 ? (half + 1) * (BN / 2) : (half == 0 ? BN : 0)); c0 += 32) {
        const int col = n0 + c0 + sub_col;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        // operand-image path (below): its bias words and its (part, head, channel) split are requested / computed here, in
        // front of the accumulator load -- the ncu source view had the epilogue warps waiting on exactly these loads
        const int ig_col = n0 + c0 + (lane & 3) * 8;
        float4 ibb0 = make_float4(0.f, 0.f, 0.f, 0.f), ibb1 = ibb0;
        int ipart = 3, ih = 0, ic = 0, iC_ = p.iC, iCpad_ = p.iCpad;
        if (p.node_slot != nullptr) {
          const int HC = p.iH * p.iC;
          ipart = ig_col / HC;
          if (ipart < 3) {
            const int within = ig_col - ipart * HC;
            if constexpr (MODE == 3) {
              if (ipart == 2) { iC_ = ex.Cv; iCpad_ = ex.Cvpad; }
            }
            ih = within / iC_; ic = within - ih * iC_;
            if (p.bias) { ibb0 = __ldg(reinterpret_cast<const float4*>(p.bias + ig_col)); ibb1 = __ldg(reinterpret_cast<const float4*>(p.bias + ig_col + 4)); }
          }
        }
        uint32_t r[32];
        tmem_ld32(t_addr + c0, r);
        tmem_ld_wait();
        if constexpr (BN == 32 && MODE == 2) {
          // fused 2-D head: this thread holds the whole 32-wide row: activation, final_mlp[2], sampler update, store
          const HeadFinalArgs& hd = ex;
          const int row = row_base + lane;
          if (row < p.M) {
            float u[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) u[k] = apply_act_rt(__uint_as_float(r[k]) + (p.bias ? __ldg(p.bias + n0 + k) : 0.f), p.act);
            const int ext = hd.row_ext ? __ldg(hd.row_ext + row) : row;
            da_step_coef cf_ = hd.coef;
            if (hd.tabs.t != nullptr && hd.step_mode != STEP_NONE) cf_ = node_coef(hd.coef, hd.tabs, ext);
            for (int c = 0; c < hd.C_out; ++c) {
              float sacc = __ldg(hd.b_b + c);
#pragma unroll
              for (int k = 0; k < 32; ++k) sacc = fmaf(__ldg(hd.w_b + c * 32 + k), u[k], sacc);
              float x = 0.f, nz = 0.f;
              const int eidx = ext * hd.C_out + c;
              if (hd.step_mode != STEP_NONE) {
                x = hd.x_in[eidx];
                if (hd.noise) nz = hd.noise[eidx];
              }
              hd.out[eidx] = step_update(hd.step_mode, x, sacc, nz, cf_);
            }
          }
          continue;
        }
        if constexpr (MODE == 3) {
          // (folded last layer only: measured on the hidden-layer GEMMs, whose skip quarter keeps the staging tile in use,
          // the same path was 9 % SLOWER than the staged one -- 0.162 vs 0.149 ms for the two K = 256 launches -- and 6 %
          // faster here.)
          // Q / K / V chunks that nobody reads in fp32 go straight from the accumulator registers to the operand images:
          // thread == row, and within an 8-channel chunk the image rows of consecutive nodes are 16 bytes apart, so the
          // 32 lanes of a store instruction already write 512 contiguous bytes -- no transpose through shared memory
          const int HC = p.iH * p.iC;
          const int part = (n0 + c0) / HC;   // a 32-column chunk lies in one part (HC % 32 == 0)
          const bool f32_needed = p.cf != nullptr && (tflags == 3u || (part == 0 ? (tflags & 1u) : (tflags & 2u)) != 0u);
          if (p.node_slot != nullptr && p.chi == nullptr && part < 3 && !f32_needed) {
            if (my_slot >= 0) {
              int iC = p.iC, iCpad = p.iCpad;
              if (part == 2) { iC = ex.Cv; iCpad = ex.Cvpad; }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int g_col = n0 + c0 + u * 8;
                const int within = g_col - part * HC;
                const int h = within / iC, c = within - h * iC;
                float4 bb0 = make_float4(0.f, 0.f, 0.f, 0.f), bb1 = bb0;
                if (p.bias) { bb0 = __ldg(reinterpret_cast<const float4*>(p.bias + g_col)); bb1 = __ldg(reinterpret_cast<const float4*>(p.bias + g_col + 4)); }
                const float vv[8] = {__uint_as_float(r[8 * u]) + bb0.x, __uint_as_float(r[8 * u + 1]) + bb0.y,
                                     __uint_as_float(r[8 * u + 2]) + bb0.z, __uint_as_float(r[8 * u + 3]) + bb0.w,
                                     __uint_as_float(r[8 * u + 4]) + bb1.x, __uint_as_float(r[8 * u + 5]) + bb1.y,
                                     __uint_as_float(r[8 * u + 6]) + bb1.z, __uint_as_float(r[8 * u + 7]) + bb1.w};
                __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  hi[e] = __float2bfloat16_rn(vv[e]);
                  lo[e] = __float2bfloat16_rn(vv[e] - __bfloat162float(hi[e]));
                }
                if (part == 0) {
                  const int tile_i = my_slot >> 7, r_ = my_slot & 127;
                  __nv_bfloat16* base = p.qimg + ((size_t)tile_i * p.iH + h) * ((size_t)2 * 128 * iCpad);
                  const size_t off = (size_t)(c >> 3) * (128 * 8) + (size_t)r_ * 8;
                  *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
                  *reinterpret_cast<uint4*>(base + (size_t)128 * iCpad + off) = *reinterpret_cast<uint4*>(lo);
                } else {
                  const int blk = my_slot >> 6, rb = my_slot & 63;
                  __nv_bfloat16* base = (part == 1 ? p.kimg : p.vimg) + ((size_t)blk * p.iH + h) * ((size_t)2 * 64 * iCpad);
                  const size_t off = (size_t)(c >> 3) * (64 * 8) + (size_t)rb * 8;
                  *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
                  *reinterpret_cast<uint4*>(base + (size_t)64 * iCpad + off) = *reinterpret_cast<uint4*>(lo);
                }
              }
            }
            continue;
          }
        }
        // thread == row: park the 32 columns of this row in the staging tile
        float4* srow = reinterpret_cast<float4*>(stage + lane * EPI_PITCH);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          srow[j ^ (lane & 7)] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        // [Q | K | V | skip] outputs: a 32-column chunk lies in one part; tiles nobody reads in fp32 skip the store
        bool f32_on = p.cf != nullptr;
        if (tflags != 3u) {
          const int part = (n0 + c0) / (p.iH * p.iC);
          f32_on = f32_on && (part == 3 || (part == 0 ? (tflags & 1u) : (tflags & 2u)) != 0u);
        }
        if (f32_on || p.chi)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + sub_row;
          const int row = row_base + rr;
          float4 v = *reinterpret_cast<const float4*>(stage + rr * EPI_PITCH + (((lane & 7) ^ (rr & 7)) << 2));
          v.x = apply_act_rt(v.x + b4.x, p.act); v.y = apply_act_rt(v.y + b4.y, p.act);
          v.z = apply_act_rt(v.z + b4.z, p.act); v.w = apply_act_rt(v.w + b4.w, p.act);
          if (row < p.M) {
            if (f32_on) *reinterpret_cast<float4*>(p.cf + (size_t)row * p.ldc + col) = v;
            if (p.chi) {
              __nv_bfloat16 h[4], l[4];
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h[e] = __float2bfloat16_rn(vv[e]);
                l[e] = __float2bfloat16_rn(vv[e] - __bfloat162float(h[e]));
              }
              *reinterpret_cast<uint2*>(p.chi + (size_t)row * p.ldsp + col) = *reinterpret_cast<uint2*>(h);
              *reinterpret_cast<uint2*>(p.clo + (size_t)row * p.ldsp + col) = *reinterpret_cast<uint2*>(l);
            }
          }
        }
        if (p.node_slot != nullptr) {
          // operand images: 8 consecutive channels (16 bytes of bf16) per lane, 8 rows per instruction
          const int part = ipart, h = ih, c = ic, iCpad = iCpad_;
          const float4 bb0 = ibb0, bb1 = ibb1;
          if (part < 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = i * 8 + (lane >> 2);
              const int row = row_base + rr;
              const int slot = slots[i];
              (void)row;
              const int ch0 = (lane & 3) * 2;
              const float4 v0 = *reinterpret_cast<const float4*>(stage + rr * EPI_PITCH + (((ch0) ^ (rr & 7)) << 2));
              const float4 v1 = *reinterpret_cast<const float4*>(stage + rr * EPI_PITCH + (((ch0 + 1) ^ (rr & 7)) << 2));
              if (slot < 0) continue;
              const float vv[8] = {v0.x + bb0.x, v0.y + bb0.y, v0.z + bb0.z, v0.w + bb0.w,
                                   v1.x + bb1.x, v1.y + bb1.y, v1.z + bb1.z, v1.w + bb1.w};
              __nv_bfloat16 hi[8], lo[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                hi[e] = __float2bfloat16_rn(vv[e]);
                lo[e] = __float2bfloat16_rn(vv[e] - __bfloat162float(hi[e]));
              }
              if (part == 0) {
                const int tile_i = slot >> 7, r_ = slot & 127;
                __nv_bfloat16* base = p.qimg + ((size_t)tile_i * p.iH + h) * ((size_t)2 * 128 * iCpad);
                const size_t off = (size_t)(c >> 3) * (128 * 8) + (size_t)r_ * 8;
                *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
                *reinterpret_cast<uint4*>(base + (size_t)128 * iCpad + off) = *reinterpret_cast<uint4*>(lo);
              } else {
                const int blk = slot >> 6, rb = slot & 63;
                __nv_bfloat16* base = (part == 1 ? p.kimg : p.vimg) + ((size_t)blk * p.iH + h) * ((size_t)2 * 64 * iCpad);
                const size_t off = (size_t)(c >> 3) * (64 * 8) + (size_t)rb * 8;   // K and V share one layout
                *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<uint4*>(hi);
                *reinterpret_cast<uint4*>(base + (size_t)64 * iCpad + off) = *reinterpret_cast<uint4*>(lo);
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS));
  }
}

// ---- host side ------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int rows, cols, ld, box_rows;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    for (int v : {k.rows, k.cols, k.ld, k.box_rows}) h = h * 1000003u ^ std::hash<int>()(v);
    return h;
  }
};

// 2-D bf16 tensor [rows, cols] with row stride ld (elements); box = [box_rows, 64], 128-byte swizzle.
bool get_tensor_map(const __nv_bfloat16* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{ptr, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  auto enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

}  // namespace

bool get_tensor_map_2d(const void* ptr, int elem_bytes, int rows, int cols, int ld, int box_rows, int box_cols, int swizzle,
                       void* out) {
  struct Key { const void* p; int e, r, c, l, br, bc, sw; bool operator==(const Key& o) const { return p == o.p && e == o.e && r == o.r && c == o.c && l == o.l && br == o.br && bc == o.bc && sw == o.sw; } };
  struct KeyHash { size_t operator()(const Key& k) const { size_t h = std::hash<const void*>()(k.p); for (int v : {k.e, k.r, k.c, k.l, k.br, k.bc, k.sw}) h = h * 1000003u ^ std::hash<int>()(v); return h; } };
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  static std::mutex mu;
  Key key{ptr, elem_bytes, rows, cols, ld, box_rows, box_cols, swizzle};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { memcpy(out, &it->second, sizeof(CUtensorMap)); return true; }
  auto enc = get_encode();
  if (!enc || box_cols > 256 || box_rows > 256 || (elem_bytes != 4 && elem_bytes != 2)) return false;
  if (((size_t)box_cols * elem_bytes) % 16 || ((size_t)ld * elem_bytes) % 16 || (reinterpret_cast<uintptr_t>(ptr) & 15)) return false;
  if (swizzle == 1 && box_cols * elem_bytes != 64) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  memcpy(out, &m, sizeof(CUtensorMap));
  return true;
}

namespace {

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int MODE>
cudaError_t launch_bn(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, int lda, const __nv_bfloat16* w_hi,
                      const __nv_bfloat16* w_lo, int ldw, const EpiParams& p, const typename EpiExtraT<MODE>::type& ex, cudaStream_t s) {
  using C = Cfg<BN>;
  CUtensorMap mah, mal, mwh, mwl;
  if (!get_tensor_map(a_hi, p.M, p.K, lda, BM, &mah) || !get_tensor_map(a_lo, p.M, p.K, lda, BM, &mal) ||
      !get_tensor_map(w_hi, p.N, p.K, ldw, BN, &mwh) || !get_tensor_map(w_lo, p.N, p.K, ldw, BN, &mwl))
    return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linear_umma_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  return launch_pdl(linear_umma_kernel<BN, MODE>, dim3(grid), dim3(NUM_THREADS), (size_t)C::SMEM_BYTES, s, mah, mal, mwh, mwl, p, ex);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_linear_umma(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, int lda,
                               const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, int ldw, const float* bias,
                               const LinearOut& out, int M, int N, int K, int act, cudaStream_t s) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (K % BK || N % 32 || (lda % 8) || (ldw % 8)) return cudaErrorInvalidValue;
  if ((out.f32 && out.ldc % 4) || (out.hi && out.ld_split % 8)) return cudaErrorInvalidValue;
  EpiParams p{bias, out.f32, out.ldc, out.hi, out.lo, out.ld_split, M, N, K, act,
              out.img_node_slot, out.qimg, out.kimg, out.vimg, out.img_H, out.img_C, out.img_Cpad, out.img_rows,
              out.img_node_slot ? out.f32_tile_flags : nullptr};
  const EpiNone none{nullptr};
  if (out.head != nullptr) {
    if (N != 32 || out.head->Nh != 32 || out.head->head_kind != DA_HEAD_2D) return cudaErrorInvalidValue;
    return launch_bn<32, 2>(a_hi, a_lo, lda, w_hi, w_lo, ldw, p, *out.head, s);
  }
  if (out.img_node_slot && out.img_Cv > 0) {   // folded last layer: [Q | K | V']
    if (act != ACT_NONE || out.img_C % 8 || out.img_Cv % 8 || (out.img_H * out.img_C) % 32 ||
        N != 2 * out.img_H * out.img_C + out.img_H * out.img_Cv || N % 128)
      return cudaErrorInvalidValue;
    const EpiFold fold{out.img_Cv, out.img_Cvpad};
    if (N % 256 == 0 && K <= 256 && M >= 4096) return launch_bn<256, 3>(a_hi, a_lo, lda, w_hi, w_lo, ldw, p, fold, s);
    return launch_bn<128, 3>(a_hi, a_lo, lda, w_hi, w_lo, ldw, p, fold, s);
  }
  if (out.img_node_slot && (act != ACT_NONE || out.img_C % 8 || N != 4 * out.img_H * out.img_C)) return cudaErrorInvalidValue;
  // short-K GEMMs (K <= 256) are bound by the L2 -> SM operand traffic (a 128 x 128 tile loads 256 KB of split-bf16
  // operands for 2.7 us of MMAs): 256-wide tiles re-use the A block twice as often
  static const bool wide = !(getenv("DA_GEMM_BN256") && getenv("DA_GEMM_BN256")[0] == '0');
  if (wide && N % 256 == 0 && K <= 256 && M >= 4096) return launch_bn<256, 0>(a_hi, a_lo, lda, w_hi, w_lo, ldw, p, none, s);
  if (N % 128 == 0) return launch_bn<128, 0>(a_hi, a_lo, lda, w_hi, w_lo, ldw, p, none, s);
  if (N % 64 == 0) return launch_bn<64, 0>(a_hi, a_lo, lda, w_hi, w_lo, ldw, p, none, s);
  return launch_bn<32, 0>(a_hi, a_lo, lda, w_hi, w_lo, ldw, p, none, s);
}

}  // namespace da
