// tcgen05 (UMMA) split-bf16 linear layer: declaration.  Implementation in gemm_umma.cu.
#pragma once
#include "common.cuh"

namespace da {

// y = act((a_hi + a_lo) @ (w_hi + w_lo)^T + bias), evaluated as the three tensor-core products
// a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation in TMEM (the a_lo*w_lo term, ~2^-18
// relative, is dropped).  a_*: [M, lda] bf16, w_*: [N, ldw] bf16, K-contiguous; K % 64 == 0, N % 16 == 0.
cudaError_t launch_linear_umma(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, int lda,
                               const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, int ldw, const float* bias,
                               const LinearOut& out, int M, int N, int K, int act, cudaStream_t s);

}  // namespace da
