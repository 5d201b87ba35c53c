// Development aid: candidate inner loops of the dense-attention softmax (attn_dense.cu), timed in isolation on the B200.
// One "block" = 64 scores per thread (one row of a 128 x 64 tile): mask by a 64-bit bitmap row, p = exp2(s * c - m),
// split p into bf16 hi / lo pairs, row sum, overflow detection.  Scores come from shared memory and the packed results
// go back to shared memory (stand-ins for tcgen05.ld / tcgen05.st).  Reports cycles per score per SM sub-partition at
// 1 / 2 / 4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/softmax_bench scripts/softmax_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define TS 64

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r; }
__device__ __forceinline__ unsigned long long pk2(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

struct Out { uint32_t ph[TS / 2], pl[TS / 2]; float lsum; uint32_t flag; };

// V0: the round-1 loop
__device__ __forceinline__ void v0(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  float lsum = 0.f, bmax = -INFINITY;
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const uint32_t w = (e < 32) ? bits.x : bits.y;
    const float s0 = ((w >> (e & 31)) & 1u) ? v[e] : -INFINITY;
    const float s1 = ((w >> ((e + 1) & 31)) & 1u) ? v[e + 1] : -INFINITY;
    bmax = fmaxf(bmax, fmaxf(s0, s1));
    const float p0 = ex2_approx(fmaf(s0, c, -msub)), p1 = ex2_approx(fmaf(s1, c, -msub));
    lsum += p0 + p1;
    const uint32_t h2 = pack_bf16x2(p0, p1);
    o.ph[e >> 1] = h2;
    o.pl[e >> 1] = pack_bf16x2(p0 - __uint_as_float(h2 << 16), p1 - __uint_as_float(h2 & 0xffff0000u));
  }
  o.lsum = lsum; o.flag = bmax > 100.f;
}
// V1: V0 without the per-score max: overflow detection by OR-ing the packed hi words (exp2 argument biased by -7 so
// that "p >= 2" <=> bit 14 of the bf16 exponent set)
__device__ __forceinline__ void v1(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  float lsum = 0.f; uint32_t orr = 0;
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const uint32_t w = (e < 32) ? bits.x : bits.y;
    const float s0 = ((w >> (e & 31)) & 1u) ? v[e] : -INFINITY;
    const float s1 = ((w >> ((e + 1) & 31)) & 1u) ? v[e + 1] : -INFINITY;
    const float p0 = ex2_approx(fmaf(s0, c, -msub)), p1 = ex2_approx(fmaf(s1, c, -msub));
    lsum += p0 + p1;
    const uint32_t h2 = pack_bf16x2(p0, p1);
    orr |= h2;
    o.ph[e >> 1] = h2;
    o.pl[e >> 1] = pack_bf16x2(p0 - __uint_as_float(h2 << 16), p1 - __uint_as_float(h2 & 0xffff0000u));
  }
  o.lsum = lsum; o.flag = orr & 0x40004000u;
}
// V2: truncating split: hi = top 16 bits (prmt packs a pair), lo = p - hi exactly, truncated again (prmt)
__device__ __forceinline__ void v2(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  float lsum = 0.f; uint32_t orr = 0;
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const uint32_t w = (e < 32) ? bits.x : bits.y;
    const float s0 = ((w >> (e & 31)) & 1u) ? v[e] : -INFINITY;
    const float s1 = ((w >> ((e + 1) & 31)) & 1u) ? v[e + 1] : -INFINITY;
    const float p0 = ex2_approx(fmaf(s0, c, -msub)), p1 = ex2_approx(fmaf(s1, c, -msub));
    lsum += p0 + p1;
    const uint32_t h2 = prmt(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
    orr |= h2;
    const float l0 = p0 - __uint_as_float(__float_as_uint(p0) & 0xffff0000u), l1 = p1 - __uint_as_float(__float_as_uint(p1) & 0xffff0000u);
    o.ph[e >> 1] = h2;
    o.pl[e >> 1] = prmt(__float_as_uint(l0), __float_as_uint(l1), 0x7632);
  }
  o.lsum = lsum; o.flag = orr & 0x40004000u;
}
// V3: Veltkamp split on the FMA pipe with packed f32x2 arithmetic: hi = RN_8bit(p) exactly (low 16 bits zero), lo = p - hi
__device__ __forceinline__ void v3(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  unsigned long long lsum2 = pk2(0.f, 0.f); uint32_t orr = 0;
  const unsigned long long c2 = pk2(c, c), nm2 = pk2(-msub, -msub), k2 = pk2(65537.f, 65537.f);
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const uint32_t w = (e < 32) ? bits.x : bits.y;
    const float s0 = ((w >> (e & 31)) & 1u) ? v[e] : -INFINITY;
    const float s1 = ((w >> ((e + 1) & 31)) & 1u) ? v[e + 1] : -INFINITY;
    float x0, x1;
    upk2(fma2(pk2(s0, s1), c2, nm2), x0, x1);
    const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
    const unsigned long long p2 = pk2(p0, p1);
    lsum2 = add2(lsum2, p2);
    const unsigned long long t2 = mul2(p2, k2);
    const unsigned long long hi2 = sub2(t2, sub2(t2, p2));
    const unsigned long long lo2 = sub2(p2, hi2);
    float h0, h1, l0, l1;
    upk2(hi2, h0, h1); upk2(lo2, l0, l1);
    const uint32_t h2 = prmt(__float_as_uint(h0), __float_as_uint(h1), 0x7632);
    orr |= h2;
    o.ph[e >> 1] = h2;
    o.pl[e >> 1] = prmt(__float_as_uint(l0), __float_as_uint(l1), 0x7632);
  }
  float a, b; upk2(lsum2, a, b);
  o.lsum = a + b; o.flag = orr & 0x40004000u;
}
// V4: V2 with the lo plane rounded by cvt.rn (one F2FP per pair instead of two)
__device__ __forceinline__ void v4(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  float lsum = 0.f; uint32_t orr = 0;
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const uint32_t w = (e < 32) ? bits.x : bits.y;
    const float s0 = ((w >> (e & 31)) & 1u) ? v[e] : -INFINITY;
    const float s1 = ((w >> ((e + 1) & 31)) & 1u) ? v[e + 1] : -INFINITY;
    const float p0 = ex2_approx(fmaf(s0, c, -msub)), p1 = ex2_approx(fmaf(s1, c, -msub));
    lsum += p0 + p1;
    const uint32_t h2 = prmt(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
    orr |= h2;
    const float l0 = p0 - __uint_as_float(__float_as_uint(p0) & 0xffff0000u), l1 = p1 - __uint_as_float(__float_as_uint(p1) & 0xffff0000u);
    o.ph[e >> 1] = h2;
    o.pl[e >> 1] = pack_bf16x2(l0, l1);
  }
  o.lsum = lsum; o.flag = orr & 0x40004000u;
}
// V5: V2 with packed scale / sum / subtract (fewer issue slots on the FMA pipe)
__device__ __forceinline__ void v5(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  unsigned long long lsum2 = pk2(0.f, 0.f); uint32_t orr = 0;
  const unsigned long long c2 = pk2(c, c), nm2 = pk2(-msub, -msub);
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const uint32_t w = (e < 32) ? bits.x : bits.y;
    const float s0 = ((w >> (e & 31)) & 1u) ? v[e] : -INFINITY;
    const float s1 = ((w >> ((e + 1) & 31)) & 1u) ? v[e + 1] : -INFINITY;
    float x0, x1;
    upk2(fma2(pk2(s0, s1), c2, nm2), x0, x1);
    const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
    const unsigned long long p2 = pk2(p0, p1);
    lsum2 = add2(lsum2, p2);
    const uint32_t h2 = prmt(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
    orr |= h2;
    float l0, l1;
    upk2(sub2(p2, pk2(__uint_as_float(__float_as_uint(p0) & 0xffff0000u), __uint_as_float(__float_as_uint(p1) & 0xffff0000u))), l0, l1);
    o.ph[e >> 1] = h2;
    o.pl[e >> 1] = prmt(__float_as_uint(l0), __float_as_uint(l1), 0x7632);
  }
  float a, b; upk2(lsum2, a, b);
  o.lsum = a + b; o.flag = orr & 0x40004000u;
}
// V6: no mask at all (what the select costs), otherwise V2
__device__ __forceinline__ void v6(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  float lsum = 0.f; uint32_t orr = 0;
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const float p0 = ex2_approx(fmaf(v[e], c, -msub)), p1 = ex2_approx(fmaf(v[e + 1], c, -msub));
    lsum += p0 + p1;
    const uint32_t h2 = prmt(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
    orr |= h2;
    const float l0 = p0 - __uint_as_float(__float_as_uint(p0) & 0xffff0000u), l1 = p1 - __uint_as_float(__float_as_uint(p1) & 0xffff0000u);
    o.ph[e >> 1] = h2;
    o.pl[e >> 1] = prmt(__float_as_uint(l0), __float_as_uint(l1), 0x7632);
  }
  o.lsum = lsum; o.flag = (orr & 0x40004000u) + bits.x;
}
// V7: only exp2 (the SFU floor): mask + fma + ex2 + sum
__device__ __forceinline__ void v7(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  float lsum = 0.f;
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const float p0 = ex2_approx(fmaf(v[e], c, -msub)), p1 = ex2_approx(fmaf(v[e + 1], c, -msub));
    lsum += p0 + p1;
    o.ph[e >> 1] = __float_as_uint(p0); o.pl[e >> 1] = __float_as_uint(p1);
  }
  o.lsum = lsum; o.flag = bits.x;
}
// V8: V2 with the mask applied as an integer AND on the sign-extended bit (shift pair) -- alternative mask idiom
__device__ __forceinline__ void v8(const float (&v)[TS], uint2 bits, float c, float msub, Out& o) {
  float lsum = 0.f; uint32_t orr = 0;
#pragma unroll
  for (int e = 0; e < TS; e += 2) {
    const uint32_t w = (e < 32) ? bits.x : bits.y;
    const float q0 = ex2_approx(fmaf(v[e], c, -msub)), q1 = ex2_approx(fmaf(v[e + 1], c, -msub));
    const float p0 = __uint_as_float(__float_as_uint(q0) & (uint32_t)((int32_t)(w << (31 - (e & 31))) >> 31));
    const float p1 = __uint_as_float(__float_as_uint(q1) & (uint32_t)((int32_t)(w << (31 - ((e + 1) & 31))) >> 31));
    lsum += p0 + p1;
    const uint32_t h2 = prmt(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
    orr |= h2;
    const float l0 = p0 - __uint_as_float(__float_as_uint(p0) & 0xffff0000u), l1 = p1 - __uint_as_float(__float_as_uint(p1) & 0xffff0000u);
    o.ph[e >> 1] = h2;
    o.pl[e >> 1] = prmt(__float_as_uint(l0), __float_as_uint(l1), 0x7632);
  }
  o.lsum = lsum; o.flag = orr & 0x40004000u;
}

template <int V, int nt>
__global__ void __launch_bounds__(nt) bench(const float* in, float* out, int iters, long long* cyc) {
  extern __shared__ float sm[];   // [TS][nt] scores, then [TS][nt] packed outputs
  const int tid = threadIdx.x;
  for (int e = 0; e < TS; ++e) sm[e * nt + tid] = in[(e * 977 + tid * 31 + blockIdx.x) & 4095];
  uint32_t* so = reinterpret_cast<uint32_t*>(sm + TS * nt);
  uint2 bits = make_uint2(0x9e3779b9u * (tid + 1), 0x85ebca6bu * (tid + 7));
  float msub = 3.0f, lacc = 0.f; uint32_t facc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float v[TS];
#pragma unroll
    for (int e = 0; e < TS; ++e) v[e] = sm[e * nt + tid];
    Out o;
    if (V == 0) v0(v, bits, 0.255f, msub, o);
    if (V == 1) v1(v, bits, 0.255f, msub, o);
    if (V == 2) v2(v, bits, 0.255f, msub, o);
    if (V == 3) v3(v, bits, 0.255f, msub, o);
    if (V == 4) v4(v, bits, 0.255f, msub, o);
    if (V == 5) v5(v, bits, 0.255f, msub, o);
    if (V == 6) v6(v, bits, 0.255f, msub, o);
    if (V == 7) v7(v, bits, 0.255f, msub, o);
    if (V == 8) v8(v, bits, 0.255f, msub, o);
#pragma unroll
    for (int e = 0; e < TS / 2; ++e) { so[e * nt + tid] = o.ph[e]; so[(TS / 2 + e) * nt + tid] = o.pl[e]; }
    lacc += o.lsum; facc |= o.flag;
    bits.x = (bits.x << 1) | (bits.x >> 31); bits.y ^= bits.x;
    msub += 0.001f;
  }
  const long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * nt + tid] = lacc + (float)facc + __uint_as_float(so[tid]);
  atomicMax((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
}

template <int V, int WPS>
void run1(const float* in, float* out, long long* cyc, const char* name) {
  const int iters = 200;
  constexpr int threads = 128 * WPS;
  const size_t smem = (size_t)2 * TS * threads * sizeof(float);
  cudaFuncSetAttribute(bench<V, threads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(cyc, 0, sizeof(long long));
    bench<V, threads><<<148, threads, smem>>>(in, out, iters, cyc);
  }
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("%-44s warps/SMSP=%d  cycles/score/warp = %6.2f   cycles/score/SMSP = %6.2f\n", name, WPS, (double)c / iters / TS,
         (double)c / iters / TS / WPS);
}
template <int V>
void run(const float* in, float* out, long long* cyc, const char* name) {
  run1<V, 1>(in, out, cyc, name); run1<V, 2>(in, out, cyc, name); run1<V, 3>(in, out, cyc, name);
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * sizeof(float)); cudaMalloc(&out, 148 * 512 * sizeof(float)); cudaMalloc(&cyc, sizeof(long long));
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = (float)((i * 2654435761u) % 1000) / 100.f - 5.f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>(in, out, cyc, "V0 round-1 loop");
  run<1>(in, out, cyc, "V1 V0 - max + OR detect");
  run<2>(in, out, cyc, "V2 truncating split (prmt/lop/fadd/prmt)");
  run<3>(in, out, cyc, "V3 Veltkamp split, packed f32x2");
  run<4>(in, out, cyc, "V4 V2 with cvt.rn lo plane");
  run<5>(in, out, cyc, "V5 V2 with packed fma/add/sub");
  run<6>(in, out, cyc, "V6 V2 without the mask");
  run<7>(in, out, cyc, "V7 fma + ex2 + sum only (SFU floor)");
  run<8>(in, out, cyc, "V8 V2, mask by AND after exp2");
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
