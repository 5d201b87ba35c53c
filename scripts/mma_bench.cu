// Development aid: issue rate of tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, of where A lives (shared
// memory "SS" / TMEM "TS") and of the B layout (K-major / MN-major, no swizzle), measured as a back-to-back chain of
// accumulating MMAs issued by one thread per CTA, one CTA per SM on every SM (and on a single SM for comparison).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mma_bench scripts/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// TS: A from TMEM; BMN: B is MN-major; NACC: number of distinct accumulators the chain cycles through
template <int N, bool TS, bool BMN, int NACC>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N) | (BMN ? (1u << 16) : 0u);
    const uint64_t da = make_desc_nosw(smem_u32(sm), 128 * 16, 128);                     // A: 128 rows x 16 k, K-major
    const uint64_t db = BMN ? make_desc_nosw(smem_u32(sm + 8192), 128, 64 * 16)          // B MN-major (the V layout of attn_dense.cu)
                            : make_desc_nosw(smem_u32(sm + 8192), N * 16, 128);          // B K-major (the K layout)
    const uint32_t d0 = tbase, a_t = tbase + 448;   // accumulators from column 0, A operand (8 columns) at 448
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const uint32_t d = d0 + (uint32_t)((u % NACC) * N);
        if (TS) mma_ts(d, a_t, db, idesc, 1u);
        else mma_ss(d, da, db, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u));
  }
}

template <int N, bool TS, bool BMN, int NACC>
void run(const char* name, int sms) {
  long long* d; cudaMalloc(&d, sizeof(long long) * 256);
  auto kern = k<N, TS, BMN, NACC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 256;
  for (int grid : {1, sms}) {
    kern<<<grid, 128, 64 * 1024>>>(8, d);   // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<grid, 128, 64 * 1024>>>(iters, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
    const double per = mean / (iters * 16.0);
    printf("%-44s grid=%3d  cycles/MMA = %7.1f  (peak-rate floor %5.1f)  ns/MMA = %6.1f  executed TFLOP/s (all CTAs) = %7.1f  %s\n", name, grid, per,
           128.0 * N / 256.0, ms * 1e6 / (iters * 16.0), 2.0 * 128 * N * 16 * iters * 16.0 * grid / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(d);
}

int main() {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<32, true, true, 1>("TS  N=32  B MN-major (P V' of the fold)", sms);
  run<32, true, true, 2>("TS  N=32  B MN-major, 2 accumulators", sms);
  run<64, true, false, 1>("TS  N=64  B K-major  (S = Q K^T)", sms);
  run<64, true, false, 2>("TS  N=64  B K-major, 2 accumulators", sms);
  run<128, true, false, 1>("TS  N=128 B K-major", sms);
  run<128, true, false, 2>("TS  N=128 B K-major, 2 accumulators", sms);
  run<256, true, false, 1>("TS  N=256 B K-major", sms);
  run<144, true, true, 1>("TS  N=144 B MN-major (old last-layer P V)", sms);
  run<64, false, false, 1>("SS  N=64  B K-major", sms);
  run<64, false, false, 2>("SS  N=64  B K-major, 2 accumulators", sms);
  run<128, false, false, 1>("SS  N=128 B K-major", sms);
  run<256, false, false, 1>("SS  N=256 B K-major", sms);
  run<32, false, true, 1>("SS  N=32  B MN-major", sms);
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
