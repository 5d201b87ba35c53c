// Greedy minimum-cost assignment of predicted piece positions to grid cells -- the metric step that
// immediately follows the sampling loop (scope row N2).
//
// Reference: greedy_cost_assignment, puzzle_diff/model/spatial_diffusion.py:179-216 -- a TorchScript while
// loop that, n times per puzzle, takes the global minimum of the still-unassigned part of the n x n distance
// matrix (boolean-mask indexing + .item() host sync + nonzero()), i.e. O(n) host round trips of O(n^2) work.
//
// Here: one CTA per puzzle, no host sync.  Every live row keeps its current best live column (value, column);
// an iteration is (1) block-wide argmin over the row minima with the reference's row-major first-occurrence
// tie break (smallest value, then smallest row, then smallest column), (2) retire that row and column,
// (3) re-scan only the rows whose cached best column was just retired (one warp per row).  Distances are
// recomputed on the fly as sqrt(dx*dx + dy*dy) in separately rounded fp32 operations (what torch.norm on the
// broadcast difference evaluates), so ties resolve exactly as in the reference.
#include "common.cuh"

namespace da {
namespace {

constexpr int ASSIGN_NT = 1024;

__device__ __forceinline__ float dist2d(float x1, float y1, float x2, float y2) {
  const float dx = __fsub_rn(x1, x2), dy = __fsub_rn(y1, y2);
  return sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

// lexicographic (value, index) minimum
__device__ __forceinline__ void argmin_pair(float& v, int& i, float ov, int oi) {
  if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__global__ void __launch_bounds__(ASSIGN_NT)
greedy_assign_kernel(const float* __restrict__ pos1, int ld1, const float* __restrict__ pos2, int ld2,
                     const int32_t* __restrict__ graph_ptr, int64_t* __restrict__ out) {
  extern __shared__ float sm[];
  const int g = blockIdx.x;
  const int n0 = graph_ptr[g], n = graph_ptr[g + 1] - n0;
  float* x1 = sm;            float* y1 = x1 + n;
  float* x2 = y1 + n;        float* y2 = x2 + n;
  float* rmin_v = y2 + n;                                   // best live column value per row
  int* rmin_j = reinterpret_cast<int*>(rmin_v + n);          // ... and its column
  int* col_alive = rmin_j + n;
  int* row_alive = col_alive + n;
  int* redo = row_alive + n;                                 // rows to re-scan this iteration
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ int n_redo, best_i, best_j;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;

  for (int i = tid; i < n; i += blockDim.x) {
    x1[i] = pos1[(size_t)(n0 + i) * ld1]; y1[i] = pos1[(size_t)(n0 + i) * ld1 + 1];
    x2[i] = pos2[(size_t)(n0 + i) * ld2]; y2[i] = pos2[(size_t)(n0 + i) * ld2 + 1];
    col_alive[i] = 1; row_alive[i] = 1;
  }
  __syncthreads();
  auto scan_row = [&](int r) {   // executed by one whole warp
    float bv = INFINITY; int bj = 0x7fffffff;
    const float ax = x1[r], ay = y1[r];
    for (int j = lane; j < n; j += 32)
      if (col_alive[j]) argmin_pair(bv, bj, dist2d(ax, ay, x2[j], y2[j]), j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      argmin_pair(bv, bj, ov, oj);
    }
    if (lane == 0) { rmin_v[r] = bv; rmin_j[r] = bj; }
  };
  for (int r = warp; r < n; r += nwarp) scan_row(r);
  __syncthreads();

  for (int it = 0; it < n; ++it) {
    // (1) global argmin over the live rows' cached minima: smallest value, then smallest row
    float bv = INFINITY; int bi = 0x7fffffff;
    for (int r = tid; r < n; r += blockDim.x)
      if (row_alive[r]) argmin_pair(bv, bi, rmin_v[r], r);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      argmin_pair(bv, bi, ov, oi);
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = lane < nwarp ? red_v[lane] : INFINITY;
      bi = lane < nwarp ? red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmin_pair(bv, bi, ov, oi);
      }
      if (lane == 0) {
        const int bj = rmin_j[bi];
        best_i = bi; best_j = bj; n_redo = 0;
        int64_t* o3 = out + (size_t)(n0 + it) * 3;
        o3[0] = bi; o3[1] = bj; o3[2] = (int64_t)bv;   // the reference stores min_val into an int64 tensor
        row_alive[bi] = 0; col_alive[bj] = 0;
      }
    }
    __syncthreads();
    // (2) rows whose cached best column was just retired must be re-scanned
    const int bj = best_j;
    for (int r = tid; r < n; r += blockDim.x)
      if (row_alive[r] && rmin_j[r] == bj) redo[atomicAdd(&n_redo, 1)] = r;
    __syncthreads();
    const int nr = n_redo;
    for (int k = warp; k < nr; k += nwarp) scan_row(redo[k]);
    __syncthreads();
  }
}

}  // namespace

cudaError_t launch_greedy_assign(const float* pos1, int ld1, const float* pos2, int ld2, const int32_t* graph_ptr,
                                 int n_graphs, int max_n, int64_t* out, cudaStream_t s) {
  if (n_graphs <= 0) return cudaSuccess;
  const size_t smem = (size_t)max_n * (5 * sizeof(float) + 4 * sizeof(int));
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(greedy_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  greedy_assign_kernel<<<n_graphs, ASSIGN_NT, smem, s>>>(pos1, ld1, pos2, ld2, graph_ptr, out);
  return cudaGetLastError();
}

}  // namespace da
