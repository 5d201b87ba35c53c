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

// Row-major 2-D tensor map [rows, cols] of 4-byte (fp32) or 2-byte (bf16) elements, row stride ld elements,
// box = [box_rows, box_cols], swizzle 0 = none / 1 = 64-byte.  Used by the dense attention epilogue (staging of
// skip / trunk-residual values, tensor stores of the layer output).  `out` receives a 128-byte CUtensorMap
// (cached per argument tuple).
bool get_tensor_map_2d(const void* ptr, int elem_bytes, int rows, int cols, int ld, int box_rows, int box_cols, int swizzle,
                       void* out);

}  // namespace da
