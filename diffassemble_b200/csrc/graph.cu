// Device-side graph preparation: edge list (PyG edge_index, int64) -> CSR by target.
// Done once per batch (da_set_graph), never inside the step loop.  A stable radix sort on the
// target id keeps the caller's edge order inside each segment, so the per-node summation order
// is deterministic run to run.
#include <cub/cub.cuh>

#include "common.cuh"

namespace da {
namespace {

__global__ void narrow_edges_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                    int n, int32_t* __restrict__ key, int32_t* __restrict__ val,
                                    int32_t* __restrict__ bad) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = src[e], d = dst[e];
  if (s < 0 || s >= n || d < 0 || d >= n) { atomicExch(bad, 1); d = 0; }
  key[e] = (int32_t)d;
  val[e] = (int32_t)e;
}

__global__ void gather_cols_kernel(const int64_t* __restrict__ src, const int32_t* __restrict__ eid, int64_t E,
                                   int n, int32_t* __restrict__ col) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= E) return;
  int64_t s = src[eid[p]];
  col[p] = (int32_t)((s < 0 || s >= n) ? 0 : s);
}

// rowptr[i] = first CSR slot whose (sorted) target is >= i
__global__ void rowptr_kernel(const int32_t* __restrict__ sorted_dst, int64_t E, int n, int32_t* __restrict__ rowptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int64_t lo = 0, hi = E;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_dst[mid] < i) lo = mid + 1; else hi = mid;
  }
  rowptr[i] = (int32_t)lo;
}

}  // namespace

void free_csr(CsrGraph* g, cudaStream_t s) {
  if (!g) return;
  tmp_free(g->rowptr, s); tmp_free(g->col, s); tmp_free(g->eid, s); tmp_free(g->weight, s);
  *g = CsrGraph();
}

namespace {
__global__ void pack_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E, int n,
                                 unsigned long long* __restrict__ key, int32_t* __restrict__ bad) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = src[e], d = dst[e];
  if (s < 0 || s >= n || d < 0 || d >= n) { atomicExch(bad, 1); s = 0; d = 0; }
  key[e] = ((unsigned long long)d << 32) | (unsigned long long)s;
}
__global__ void unpack_runs_kernel(const unsigned long long* __restrict__ ukey, const int32_t* __restrict__ cnt,
                                   const int32_t* __restrict__ nruns, int32_t* __restrict__ col,
                                   float* __restrict__ weight, int32_t* __restrict__ udst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *nruns) return;
  col[i] = (int32_t)(ukey[i] & 0xffffffffull);
  udst[i] = (int32_t)(ukey[i] >> 32);
  weight[i] = (float)cnt[i];
}
}  // namespace

// CSR by target with duplicate edges collapsed into (source, multiplicity) pairs: a duplicate edge
// contributes a second identical term to the segment softmax, i.e. weight * exp(score).
cudaError_t build_csr_compressed(const int64_t* src, const int64_t* dst, int64_t E, int n, CsrGraph* g, cudaStream_t s,
                                 const char** err) {
  *err = "";
  free_csr(g, s);
  if (E >= (int64_t)1 << 31) { *err = "edge count exceeds int32 range"; return cudaErrorInvalidValue; }
  cudaError_t ce;
#define DA_TRY(x) do { ce = (x); if (ce != cudaSuccess) { *err = #x; goto fail; } } while (0)
  unsigned long long *key = nullptr, *key_sorted = nullptr, *ukey = nullptr;
  int32_t *cnt = nullptr, *nruns = nullptr, *bad = nullptr, *udst = nullptr;
  void* tmp = nullptr;
  size_t tb1 = 0, tb2 = 0;
  int32_t bad_h = 0, nruns_h = 0;
  int bits = 1;
  g->n = n;
  DA_TRY(tmp_alloc(&g->rowptr, sizeof(int32_t) * (size_t)(n + 1), s));
  if (E == 0) {
    g->E = 0;
    DA_TRY(tmp_alloc(&g->col, sizeof(int32_t), s));
    DA_TRY(tmp_alloc(&g->weight, sizeof(float), s));
    DA_TRY(cudaMemsetAsync(g->rowptr, 0, sizeof(int32_t) * (size_t)(n + 1), s));
    return cudaSuccess;
  }
  DA_TRY(tmp_alloc(&key, sizeof(unsigned long long) * (size_t)E, s));
  DA_TRY(tmp_alloc(&key_sorted, sizeof(unsigned long long) * (size_t)E, s));
  DA_TRY(tmp_alloc(&ukey, sizeof(unsigned long long) * (size_t)E, s));
  DA_TRY(tmp_alloc(&cnt, sizeof(int32_t) * (size_t)E, s));
  DA_TRY(tmp_alloc(&udst, sizeof(int32_t) * (size_t)E, s));
  DA_TRY(tmp_alloc(&nruns, sizeof(int32_t), s));
  DA_TRY(tmp_alloc(&bad, sizeof(int32_t), s));
  DA_TRY(cudaMemsetAsync(bad, 0, sizeof(int32_t), s));
  pack_keys_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(src, dst, E, n, key, bad);
  DA_TRY(cudaGetLastError());
  while ((1ll << bits) < (long long)n + 1 && bits < 31) ++bits;
  DA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tb1, key, key_sorted, (int)E, 0, 32 + bits, s));
  DA_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, tb2, key_sorted, ukey, cnt, nruns, (int)E, s));
  if (tb2 > tb1) tb1 = tb2;
  DA_TRY(tmp_alloc(&tmp, tb1, s));
  DA_TRY(cub::DeviceRadixSort::SortKeys(tmp, tb1, key, key_sorted, (int)E, 0, 32 + bits, s));
  DA_TRY(cub::DeviceRunLengthEncode::Encode(tmp, tb1, key_sorted, ukey, cnt, nruns, (int)E, s));
  DA_TRY(cudaMemcpyAsync(&nruns_h, nruns, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  DA_TRY(cudaMemcpyAsync(&bad_h, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  DA_TRY(cudaStreamSynchronize(s));
  if (bad_h) { *err = "edge_index entry outside [0, num_total)"; ce = cudaErrorInvalidValue; goto fail; }
  g->E = nruns_h;
  DA_TRY(tmp_alloc(&g->col, sizeof(int32_t) * (size_t)nruns_h, s));
  DA_TRY(tmp_alloc(&g->weight, sizeof(float) * (size_t)nruns_h, s));
  unpack_runs_kernel<<<(nruns_h + 255) / 256, 256, 0, s>>>(ukey, cnt, nruns, g->col, g->weight, udst);
  DA_TRY(cudaGetLastError());
  rowptr_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>(udst, nruns_h, n, g->rowptr);
  DA_TRY(cudaGetLastError());
  DA_TRY(cudaStreamSynchronize(s));
  tmp_free(key, s); tmp_free(key_sorted, s); tmp_free(ukey, s); tmp_free(cnt, s); tmp_free(udst, s); tmp_free(nruns, s); tmp_free(bad, s); tmp_free(tmp, s);
  return cudaSuccess;
fail:
  tmp_free(key, s); tmp_free(key_sorted, s); tmp_free(ukey, s); tmp_free(cnt, s); tmp_free(udst, s); tmp_free(nruns, s); tmp_free(bad, s); tmp_free(tmp, s);
  free_csr(g, s);
  return ce;
#undef DA_TRY
}

cudaError_t build_csr(const int64_t* src, const int64_t* dst, int64_t E, int n, CsrGraph* g, cudaStream_t s,
                      const char** err) {
  *err = "";
  free_csr(g, s);
  if (E >= (int64_t)1 << 31) { *err = "edge count exceeds int32 range"; return cudaErrorInvalidValue; }
  cudaError_t ce;
#define DA_TRY(x) do { ce = (x); if (ce != cudaSuccess) { *err = #x; goto fail; } } while (0)
  int32_t *key = nullptr, *val = nullptr, *key_sorted = nullptr, *bad = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int32_t bad_h = 0;
  int bits = 1;
  g->n = n; g->E = E;
  DA_TRY(tmp_alloc(&g->rowptr, sizeof(int32_t) * (size_t)(n + 1), s));
  DA_TRY(tmp_alloc(&g->col, sizeof(int32_t) * (size_t)(E > 0 ? E : 1), s));
  DA_TRY(tmp_alloc(&g->eid, sizeof(int32_t) * (size_t)(E > 0 ? E : 1), s));
  if (E == 0) {
    DA_TRY(cudaMemsetAsync(g->rowptr, 0, sizeof(int32_t) * (size_t)(n + 1), s));
    return cudaSuccess;
  }
  DA_TRY(tmp_alloc(&key, sizeof(int32_t) * (size_t)E, s));
  DA_TRY(tmp_alloc(&val, sizeof(int32_t) * (size_t)E, s));
  DA_TRY(tmp_alloc(&key_sorted, sizeof(int32_t) * (size_t)E, s));
  DA_TRY(tmp_alloc(&bad, sizeof(int32_t), s));
  DA_TRY(cudaMemsetAsync(bad, 0, sizeof(int32_t), s));
  narrow_edges_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(src, dst, E, n, key, val, bad);
  DA_TRY(cudaGetLastError());
  while ((1ll << bits) < (long long)n + 1 && bits < 31) ++bits;
  DA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key_sorted, val, g->eid, (int)E, 0, bits, s));
  DA_TRY(tmp_alloc(&tmp, tmp_bytes, s));
  DA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key_sorted, val, g->eid, (int)E, 0, bits, s));
  gather_cols_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(src, g->eid, E, n, g->col);
  DA_TRY(cudaGetLastError());
  rowptr_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>(key_sorted, E, n, g->rowptr);
  DA_TRY(cudaGetLastError());
  DA_TRY(cudaMemcpyAsync(&bad_h, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  DA_TRY(cudaStreamSynchronize(s));
  tmp_free(key, s); tmp_free(val, s); tmp_free(key_sorted, s); tmp_free(bad, s); tmp_free(tmp, s);
  if (bad_h) { *err = "edge_index entry outside [0, num_total)"; free_csr(g, s); return cudaErrorInvalidValue; }
  return cudaSuccess;
fail:
  tmp_free(key, s); tmp_free(val, s); tmp_free(key_sorted, s); tmp_free(bad, s); tmp_free(tmp, s);
  free_csr(g, s);
  return ce;
#undef DA_TRY
}

}  // namespace da

// ---------------------------------------------------------------------------------------------
// Scope row N3: Exphander topology on the device.  Given the per-graph random permutations (drawn
// by the caller with the reference's numpy Generator so seeds match, puzzle_dataset.py:133-152),
// writes the batched, symmetrised edge list in exactly the reference's order:
//   per graph: A = tile(perm, d/2) [+ perm[:n/2]],  Bv = [roll(perm, s) for s = 1..d/2] [+ perm[n/2:]]
//              senders = [A, Bv], receivers = [Bv, A];  node ids offset by g * n (PyG collation).
// The 250 MB int64 edge_index of a 32 x 900-node batch is then never built on, or copied from, the host.
// ---------------------------------------------------------------------------------------------
namespace da {
namespace {
__global__ void expander_edges_kernel(const int32_t* __restrict__ perm, int n, int degree, int n_graphs,
                                      int64_t* __restrict__ src, int64_t* __restrict__ dst) {
  const int half = degree / 2;
  const long long e_half = (long long)n * half + ((degree & 1) ? n / 2 : 0);
  const long long e_graph = 2 * e_half;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= e_half * n_graphs) return;
  const int g = (int)(idx / e_half);
  const long long k = idx % e_half;
  const int32_t* p = perm + (size_t)g * n;
  int a, b;
  if (k < (long long)n * half) {
    const int rep = (int)(k / n), pos = (int)(k % n);
    int q = pos - (rep + 1);            // np.roll(perm, s)[pos] == perm[(pos - s) mod n]
    q %= n; if (q < 0) q += n;
    a = p[pos]; b = p[q];
  } else {
    const int t = (int)(k - (long long)n * half);   // perfect matching of the odd-degree case
    a = p[t]; b = p[n / 2 + t];
  }
  const long long off = (long long)g * n;
  const long long base = (long long)g * e_graph;
  src[base + k] = a + off;           dst[base + k] = b + off;
  src[base + e_half + k] = b + off;  dst[base + e_half + k] = a + off;
}
}  // namespace

cudaError_t launch_expander_edges(const int32_t* perm, int n, int degree, int n_graphs, int64_t* src, int64_t* dst,
                                  cudaStream_t s) {
  const long long e_half = (long long)n * (degree / 2) + ((degree & 1) ? n / 2 : 0);
  const long long total = e_half * n_graphs;
  if (total <= 0) return cudaSuccess;
  expander_edges_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(perm, n, degree, n_graphs, src, dst);
  return cudaGetLastError();
}
}  // namespace da
