// Development aid: issue throughput of the instructions the dense-attention softmax loop is built from,
// measured on the box's B200 (cycles per warp-instruction per SM sub-partition, at 1 / 2 / 4 warps per sub-partition).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_bench scripts/pipe_bench.cu && /tmp/pipe_bench
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 256
#define UNR 8

enum Op { EX2, CVT_BF16X2, PRMT, LOP3, SELP, FFMA, FADD, FFMA2, FADD2, FMAX3, MIX_OLD, MIX_NEW, LOP3P_SEL, NOPS };
const char* names[] = {"ex2.approx.f32", "cvt.rn.bf16x2.f32", "prmt.b32", "lop3.b32", "selp.f32", "fma.f32", "add.f32",
                       "fma.f32x2", "add.f32x2", "max3.f32", "mix_old(score)", "mix_new(score)", "bit-test+selp"};

template <int OP>
__global__ void k(float* out, int n_iter, long long* cyc) {
  float a[UNR], b[UNR];
  unsigned u[UNR];
#pragma unroll
  for (int i = 0; i < UNR; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = 1.0f + i * 0.01f; u[i] = threadIdx.x * 2654435761u + i; }
  unsigned w = threadIdx.x * 40503u + 12345u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < n_iter; ++it) {
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == CVT_BF16X2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i]));
      if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) % UNR]));
      if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % UNR]), "r"(w));
      if (OP == SELP) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.f32 %0, %0, %1, p;}" : "+f"(a[i]) : "f"(b[i]), "r"(w));
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) % UNR]));
      if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      if (OP == FMAX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) % UNR]));
      if (OP == LOP3P_SEL) {   // the current mask idiom: bit test -> predicate -> select
        asm volatile("{.reg .pred p; .reg .b32 t; and.b32 t, %2, %3; setp.ne.u32 p, t, 0; selp.f32 %0, %0, %1, p;}"
                     : "+f"(a[i]) : "f"(b[i]), "r"(w), "r"(1u << i));
      }
    }
    if (OP == FFMA2 || OP == FADD2) {
#pragma unroll
      for (int i = 0; i < UNR; i += 2) {
        unsigned long long x, y, z;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b[i]), "f"(b[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(b[(i + 2) % UNR]), "f"(b[(i + 3) % UNR]));
        // 4 packed ops per pair so that the op count per iteration matches UNR * ... (see host scaling)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
          else asm volatile("add.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
        }
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
      }
    }
    if (OP == MIX_OLD) {   // per score: bit test, select, fma, ex2, add(sum), max; per pair: cvt hi, 2 extract, 2 sub, cvt lo
#pragma unroll
      for (int i = 0; i < UNR; i += 2) {
        float s0 = a[i], s1 = a[i + 1];
        float x0 = ((w >> i) & 1u) ? s0 : -INFINITY, x1 = ((w >> (i + 1)) & 1u) ? s1 : -INFINITY;
        b[0] = fmaxf(b[0], fmaxf(x0, x1));
        float p0, p1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(x0, 1.25f, -3.f)));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(x1, 1.25f, -3.f)));
        b[1] += p0 + p1;
        unsigned h2, l2;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(p1), "f"(p0));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(p1 - __uint_as_float(h2 & 0xffff0000u)), "f"(p0 - __uint_as_float(h2 << 16)));
        u[i] ^= h2; u[i + 1] ^= l2;
        a[i] = s0 + 0.125f; a[i + 1] = s1 - 0.125f;
      }
    }
    if (OP == MIX_NEW) {   // truncating hi (prmt), lo by prmt too, OR-based overflow detection, packed fma / add where possible
#pragma unroll
      for (int i = 0; i < UNR; i += 2) {
        float s0 = a[i], s1 = a[i + 1];
        float x0 = ((w >> i) & 1u) ? s0 : -INFINITY, x1 = ((w >> (i + 1)) & 1u) ? s1 : -INFINITY;
        float p0, p1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(x0, 1.25f, -3.f)));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(x1, 1.25f, -3.f)));
        unsigned h2, l2;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h2) : "r"(__float_as_uint(p0)), "r"(__float_as_uint(p1)));
        float l0 = p0 - __uint_as_float(__float_as_uint(p0) & 0xffff0000u), l1 = p1 - __uint_as_float(__float_as_uint(p1) & 0xffff0000u);
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(l2) : "r"(__float_as_uint(l0)), "r"(__float_as_uint(l1)));
        b[1] += p0 + p1;
        u[0] |= h2;
        u[i] ^= h2; u[i + 1] ^= l2;
        a[i] = s0 + 0.125f; a[i + 1] = s1 - 0.125f;
      }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < UNR; ++i) acc += a[i] + b[i] + __uint_as_float(u[i] & 0x3f800000u);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(float* out, long long* cyc) {
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp;
    k<OP><<<148, threads>>>(out, ITER, cyc);   // warm
    k<OP><<<148, threads>>>(out, ITER, cyc);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double ops = (double)ITER * UNR;   // "ops" per warp: one per array element per iteration (mix: one score)
    printf("%-20s warps/SMSP=%d  cycles=%8lld  cycles per warp-op = %.3f   per SMSP-op = %.3f\n", names[OP], warps_per_smsp, c,
           c / ops, c / ops / warps_per_smsp);
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  run<EX2>(out, cyc); run<CVT_BF16X2>(out, cyc); run<PRMT>(out, cyc); run<LOP3>(out, cyc); run<SELP>(out, cyc);
  run<FFMA>(out, cyc); run<FADD>(out, cyc); run<FFMA2>(out, cyc); run<FADD2>(out, cyc); run<FMAX3>(out, cyc);
  run<LOP3P_SEL>(out, cyc); run<MIX_OLD>(out, cyc); run<MIX_NEW>(out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
