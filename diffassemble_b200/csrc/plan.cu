// Dense-tile attention planner (runs once per batch inside da_set_graph).
//
// Splits the edge multiset of the PyG-style batch into
//   (1) per-graph adjacency bitmaps for graphs that are big and dense enough to be worth running
//       as masked dense attention on the tensor cores, and
//   (2) a residual CSR (duplicates, virtual-node wiring, cross-graph edges, small / sparse graphs).
// The two parts partition the multiset exactly, so dense + residual == the reference's segment
// softmax over all in-edges (exophormer_gnn.py:198-200 wiring included).
#include <cub/cub.cuh>

#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace da {
namespace {

constexpr int MIN_DENSE_NODES = 48;   // smaller graphs stay on the CSR warp kernel
// residual in-degree above which a row gets a whole CTA per (row, head) (virtual hub nodes: ~1000 in-edges).  Rows of a few
// dozen in-edges -- every node of the 2..20-fragment graphs of the 3-D configuration -- stay on the warp-per-node kernel:
// with the threshold at 16 those graphs launched 2 400 CTAs of 1024 threads per layer for 17..20 edges each (64 us per
// launch, 8 launches per step = 5/6 of the whole step).
constexpr int HEAVY_MIN_DEG = 96;
constexpr int DENSITY_DIV = 16;       // dense if in-graph edges >= n^2 / 16

__global__ void count_ingraph_edges_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                           const int64_t* __restrict__ batch, int num_real,
                                           unsigned long long* __restrict__ cnt) {
  // one atomic per (warp, graph) instead of one per edge: PyG batches keep a graph's edges together, so an
  // un-aggregated version fires ~E atomics at a few dozen addresses -- slow by itself and, when the plan is built
  // on a side stream (prefetch), a storm that stalls the memory traffic of the sampling kernels running next to it
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  long long g = -1;
  if (e < E) {
    int64_t s = src[e], d = dst[e];
    if (s >= 0 && d >= 0 && s < num_real && d < num_real) {
      int64_t gd = batch[d];
      if (batch[s] == gd) g = (long long)gd;
    }
  }
  const unsigned peers = __match_any_sync(0xffffffffu, g);
  if (g >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&cnt[g], (unsigned long long)__popc(peers));
}

// flag[e] = 1 -> residual edge (goes to the CSR), 0 -> recorded in the bitmap.  int_of_ext (optional): the internal
// number of each real node (a permutation inside every graph); bitmap rows / columns are internal.
__global__ void classify_edges_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                      const int64_t* __restrict__ batch, int num_real,
                                      const int32_t* __restrict__ g_node0, const int64_t* __restrict__ g_bm_off,
                                      const int32_t* __restrict__ g_bm_words, uint32_t* __restrict__ bitmap,
                                      uint8_t* __restrict__ flag, const int32_t* __restrict__ int_of_ext) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = src[e], d = dst[e];
  uint8_t residual = 1;
  if (s >= 0 && d >= 0 && s < num_real && d < num_real) {
    int64_t g = batch[d];
    if (batch[s] == g && g_bm_off[g] >= 0) {
      if (int_of_ext) { s = int_of_ext[s]; d = int_of_ext[d]; }
      const int i = (int)(d - g_node0[g]), j = (int)(s - g_node0[g]);
      uint32_t* w = bitmap + g_bm_off[g] + (int64_t)i * g_bm_words[g] + (j >> 5);
      const uint32_t bit = 1u << (j & 31);
      const uint32_t old = atomicOr(w, bit);
      residual = (old & bit) ? 1 : 0;  // a duplicate of an edge already in the bitmap stays in the CSR
    }
  }
  if (flag) flag[e] = residual;
}

// Greedy walk along maximal neighbourhood overlap, one CTA per dense-tile graph: from the current node go to the unvisited
// neighbour that shares the most neighbours with it (bitmap row AND + popcount); when there is none, to the lowest
// unvisited node.  On the reference's Exphander graphs (ring position p joined to p +- 1 .. p +- d/2) the overlap with
// position p + k is d - 1 - k, strictly decreasing, so the walk follows the ring and the relabelled adjacency is a band.
// Any other graph just gets some permutation, which is harmless.  Writes ext_of_int / int_of_ext for the graph's nodes.
__global__ void __launch_bounds__(256)
ring_order_kernel(const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ g_node0, const int32_t* __restrict__ g_n,
                  const int64_t* __restrict__ g_bm_off, const int32_t* __restrict__ g_bm_words, const uint8_t* __restrict__ g_reorder,
                  int32_t* __restrict__ ext_of_int, int32_t* __restrict__ int_of_ext) {
  const int g = blockIdx.x;
  if (!g_reorder[g]) return;
  extern __shared__ uint32_t ro_sm[];
  const int n = g_n[g], words = g_bm_words[g], node0 = g_node0[g];
  const uint32_t* bm = bitmap + g_bm_off[g];
  uint32_t* visited = ro_sm;            // [words]
  uint32_t* cur_row = ro_sm + words;    // [words]
  __shared__ unsigned long long warp_best[8];
  __shared__ int cur_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int w = tid; w < words; w += 256) visited[w] = 0u;
  if (tid == 0) cur_s = 0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const int cur = cur_s;
    if (tid == 0) {
      visited[cur >> 5] |= 1u << (cur & 31);
      ext_of_int[node0 + k] = node0 + cur;
      int_of_ext[node0 + cur] = node0 + k;
    }
    for (int w = tid; w < words; w += 256) cur_row[w] = bm[(size_t)cur * words + w];
    __syncthreads();
    if (k == n - 1) break;
    unsigned long long best = 0ull;
    for (int v = tid; v < n; v += 256) {
      if (!((cur_row[v >> 5] >> (v & 31)) & 1u) || ((visited[v >> 5] >> (v & 31)) & 1u)) continue;
      const uint32_t* rv = bm + (size_t)v * words;
      int score = 0;
      for (int w = 0; w < words; ++w) score += __popc(cur_row[w] & rv[w]);
      const unsigned long long key = ((unsigned long long)(score + 1) << 32) | (unsigned long long)(0x7fffffff - v);
      best = key > best ? key : best;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) warp_best[wid] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned long long b = 0ull;
      for (int i = 0; i < 8; ++i) b = warp_best[i] > b ? warp_best[i] : b;
      int nxt = -1;
      if (b != 0ull) nxt = 0x7fffffff - (int)(b & 0xffffffffull);
      else {   // dead end: lowest unvisited node
        for (int w = 0; w < words && nxt < 0; ++w) {
          const uint32_t free_bits = ~visited[w];
          if (free_bits) { const int v = w * 32 + __ffs(free_bits) - 1; if (v < n) nxt = v; }
        }
      }
      cur_s = nxt;
    }
    __syncthreads();
  }
}

__global__ void iota_kernel(int32_t* __restrict__ a, int32_t* __restrict__ b, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = i; b[i] = i; }
}

__global__ void relabel_kernel(int64_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ int_of_ext, int num_real) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int64_t v = ids[i]; if (v >= 0 && v < num_real) ids[i] = int_of_ext[v]; }
}

// Per 128-row tile: which 64-source blocks of the graph hold at least one bit on a valid row (listed, in order), and
// which of those hold ALL bits (valid rows x the block's 64 columns).  One CTA of 128 threads per tile.
__global__ void __launch_bounds__(128)
tile_blocks_kernel(TileInfo* __restrict__ tiles, const uint32_t* __restrict__ bitmap, uint16_t* __restrict__ blk_list, int max_blocks,
                   unsigned long long* __restrict__ counters) {
  const int t = blockIdx.x, r = threadIdx.x;
  TileInfo ti = tiles[t];
  const int nblk = (ti.gn + 63) / 64;
  const bool valid = r < ti.rows;
  const uint32_t* row = bitmap + ti.bm_off + (size_t)(ti.row0 + r) * ti.bm_words;
  __shared__ int n_list_s;
  if (r == 0) n_list_s = 0;
  __syncthreads();
  int n_full = 0;
  for (int b = 0; b < nblk; ++b) {
    uint32_t w0 = 0u, w1 = 0u;
    if (valid) { w0 = row[2 * b]; w1 = row[2 * b + 1]; }
    const int any = __syncthreads_or(valid && (w0 | w1) != 0u);
    const int all = __syncthreads_and(!valid || (w0 & w1) == 0xffffffffu);
    if (any && r == 0) {
      blk_list[(size_t)t * max_blocks + n_list_s] = (uint16_t)((b << 1) | (all ? 1 : 0));
      n_list_s++;
      n_full += all ? 1 : 0;
    }
  }
  __syncthreads();
  if (r == 0) {
    if (n_list_s == 0) { blk_list[(size_t)t * max_blocks] = 0; n_list_s = 1; }   // the kernel always visits one block
    tiles[t].n_list = n_list_s;
    tiles[t].list_off = t * max_blocks;
    atomicAdd(&counters[0], (unsigned long long)nblk);
    atomicAdd(&counters[1], (unsigned long long)n_list_s);
    atomicAdd(&counters[2], (unsigned long long)n_full);
  }
}

__global__ void set_bits_kernel(const int64_t* __restrict__ word, const uint32_t* __restrict__ bit, int n,
                                uint32_t* __restrict__ bitmap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicOr(bitmap + word[i], bit[i]);
}

__global__ void check_sorted_kernel(const int64_t* __restrict__ batch, int n, int32_t* __restrict__ bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < n && batch[i + 1] < batch[i]) atomicExch(bad, 1);
  if (i < n && batch[i] < 0) atomicExch(bad, 1);
}

}  // namespace

void free_plan(DensePlan* p, cudaStream_t s) {
  if (!p) return;
  tmp_free(p->tiles, s); tmp_free(p->node_slot, s); tmp_free(p->bitmap, s); tmp_free(p->light, s); tmp_free(p->heavy, s); tmp_free(p->row_fused, s); tmp_free(p->light_nf, s); tmp_free(p->csr_rows, s); tmp_free(p->x_src, s); tmp_free(p->x_slot, s); tmp_free(p->f32_tile_flags[0], s); tmp_free(p->f32_tile_flags[1], s);
  tmp_free(p->ext_of_int, s); tmp_free(p->blk_list, s);
  free_csr(&p->residual, s);
  *p = DensePlan();
}

size_t dense_image_elems(int n_tiles, int H, int Cpad) { return (size_t)n_tiles * 128 * H * Cpad * 2; }

cudaError_t build_dense_plan(const int64_t* src, const int64_t* dst, int64_t E, const int64_t* batch, int num_real,
                             int num_total, DensePlan* plan, cudaStream_t s, const char** err, bool allow_reorder) {
  *err = "";
  free_plan(plan, s);
  cudaError_t ce = cudaSuccess;
#define DA_TRY(x) do { ce = (x); if (ce != cudaSuccess) { *err = #x; goto fail; } } while (0)
  std::vector<int64_t> hbatch(num_real);
  std::vector<unsigned long long> hcnt;
  std::vector<int32_t> g_node0, g_n, g_bm_words, g_tile0, node_slot(num_total, -1);
  std::vector<int64_t> g_bm_off;
  std::vector<TileInfo> tiles;
  std::vector<int32_t> extra_sources;
  unsigned long long* dcnt = nullptr;
  int32_t *d_g_node0 = nullptr, *d_g_bm_words = nullptr, *d_bad = nullptr;
  int64_t *d_g_bm_off = nullptr, *res_src = nullptr, *res_dst = nullptr, *d_nsel = nullptr;
  uint8_t* flag = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, tb2 = 0;
  int32_t bad_h = 0;
  int64_t nsel = 0, n_res = 0;
  int B = 0;
  size_t words = 0;
  int block64 = 0;
  int32_t* d_int_of_ext = nullptr; int32_t* d_g_n = nullptr; uint8_t* d_g_reorder = nullptr;
  unsigned long long* d_counters = nullptr;
  std::vector<uint8_t> g_reorder;
  int n_reorder = 0, words_max = 0;
  static const bool no_reorder_env = getenv("DA_NO_REORDER") != nullptr && getenv("DA_NO_REORDER")[0] == '1';

  DA_TRY(tmp_alloc(&d_bad, sizeof(int32_t), s));
  DA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), s));
  check_sorted_kernel<<<(num_real + 255) / 256, 256, 0, s>>>(batch, num_real, d_bad);
  DA_TRY(cudaGetLastError());
  DA_TRY(cudaMemcpyAsync(hbatch.data(), batch, sizeof(int64_t) * num_real, cudaMemcpyDeviceToHost, s));
  DA_TRY(cudaMemcpyAsync(&bad_h, d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  DA_TRY(cudaStreamSynchronize(s));
  if (bad_h) {  // not PyG collation order: no dense graphs, everything through the CSR
    B = 0;
  } else {
    B = (int)hbatch[num_real - 1] + 1;
  }
  g_node0.assign(B, 0); g_n.assign(B, 0); g_bm_words.assign(B, 0); g_bm_off.assign(B, -1); g_tile0.assign(B, -1);
  g_reorder.assign(B, 0);
  for (int i = 0; i < num_real && B > 0; ++i) g_n[hbatch[i]]++;
  for (int g = 1; g < B; ++g) g_node0[g] = g_node0[g - 1] + g_n[g - 1];
  if (B > 0 && E > 0) {
    DA_TRY(tmp_alloc(&dcnt, sizeof(unsigned long long) * B, s));
    DA_TRY(cudaMemsetAsync(dcnt, 0, sizeof(unsigned long long) * B, s));
    count_ingraph_edges_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(src, dst, E, batch, num_real, dcnt);
    DA_TRY(cudaGetLastError());
    hcnt.resize(B);
    DA_TRY(cudaMemcpyAsync(hcnt.data(), dcnt, sizeof(unsigned long long) * B, cudaMemcpyDeviceToHost, s));
    DA_TRY(cudaStreamSynchronize(s));
  } else {
    hcnt.assign(B, 0);
  }
  // host decisions: which graphs go dense, tile table, bitmap offsets
  for (int g = 0; g < B; ++g) {
    const long long n = g_n[g];
    if (n < MIN_DENSE_NODES || (long long)hcnt[g] * DENSITY_DIV < n * n) continue;
    const int ntile = (int)((n + 127) / 128);
    // worth reordering: more than two source blocks and at most half dense.  (A band of relative width w leaves
    // 1 - w - 128 / n of a tile's blocks empty; above ~50 % nothing is left to skip -- and the reference's odd-degree
    // Exphander graphs add a perfect matching p <-> p + n/2 that lands in the middle of what would be empty.)
    static const bool force_reorder = getenv("DA_FORCE_REORDER") != nullptr && getenv("DA_FORCE_REORDER")[0] == '1';
    if (allow_reorder && !no_reorder_env && n > 128 && ((long long)hcnt[g] * 2 <= n * n || (force_reorder && (long long)hcnt[g] * 10 < n * n * 9))) {
      g_reorder[g] = 1; ++n_reorder;
    }
    g_bm_words[g] = ntile * 4;   // one bit per image row of the graph (its nodes + the padding rows extra sources may use)
    g_bm_off[g] = (int64_t)words;
    g_tile0[g] = (int)tiles.size();
    words += (size_t)ntile * 128 * g_bm_words[g];
    for (int t = 0; t < ntile; ++t) {
      TileInfo ti;
      ti.node0 = g_node0[g] + t * 128;
      ti.rows = (int)std::min<long long>(128, n - t * 128);
      ti.gblock0 = block64;
      ti.gn = (int)n;
      ti.bm_words = g_bm_words[g];
      ti.row0 = t * 128;
      ti.bm_off = g_bm_off[g];
      const int tile_idx = (int)tiles.size();
      for (int r = 0; r < ti.rows; ++r) node_slot[ti.node0 + r] = tile_idx * 128 + r;
      tiles.push_back(ti);
    }
    block64 += ntile * 2;  // source blocks are laid out at image-row granularity: 2 per 128-row tile
    plan->n_dense_graphs++;
    plan->n_dense_edges += 0;
  }
  plan->n_tiles = (int)tiles.size();
  plan->bitmap_words = words;
  DA_TRY(tmp_alloc(&plan->node_slot, sizeof(int32_t) * (size_t)num_total, s));
  DA_TRY(cudaMemcpyAsync(plan->node_slot, node_slot.data(), sizeof(int32_t) * (size_t)num_total, cudaMemcpyHostToDevice, s));
  if (plan->n_tiles > 0) {
    DA_TRY(tmp_alloc(&plan->tiles, sizeof(TileInfo) * tiles.size(), s));   // uploaded after the promotion pass (gn may grow)
    DA_TRY(tmp_alloc(&plan->bitmap, sizeof(uint32_t) * words, s));
    DA_TRY(cudaMemsetAsync(plan->bitmap, 0, sizeof(uint32_t) * words, s));
    DA_TRY(tmp_alloc(&d_g_node0, sizeof(int32_t) * B, s));
    DA_TRY(tmp_alloc(&d_g_bm_words, sizeof(int32_t) * B, s));
    DA_TRY(tmp_alloc(&d_g_bm_off, sizeof(int64_t) * B, s));
    DA_TRY(cudaMemcpyAsync(d_g_node0, g_node0.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
    DA_TRY(cudaMemcpyAsync(d_g_bm_words, g_bm_words.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
    DA_TRY(cudaMemcpyAsync(d_g_bm_off, g_bm_off.data(), sizeof(int64_t) * B, cudaMemcpyHostToDevice, s));
  }
  if (plan->n_tiles > 0 && E > 0 && n_reorder > 0) {
    // ---- internal node order: bitmap in the caller's order -> greedy ring walk per graph -> bitmap rebuilt in internal order
    for (int g = 0; g < B; ++g) if (g_reorder[g] && g_bm_words[g] > words_max) words_max = g_bm_words[g];
    DA_TRY(tmp_alloc(&plan->ext_of_int, sizeof(int32_t) * (size_t)num_total, s));
    DA_TRY(tmp_alloc(&d_int_of_ext, sizeof(int32_t) * (size_t)num_total, s));
    DA_TRY(tmp_alloc(&d_g_n, sizeof(int32_t) * B, s));
    DA_TRY(tmp_alloc(&d_g_reorder, (size_t)B, s));
    DA_TRY(cudaMemcpyAsync(d_g_n, g_n.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
    DA_TRY(cudaMemcpyAsync(d_g_reorder, g_reorder.data(), (size_t)B, cudaMemcpyHostToDevice, s));
    iota_kernel<<<(num_total + 255) / 256, 256, 0, s>>>(plan->ext_of_int, d_int_of_ext, num_total);
    DA_TRY(cudaGetLastError());
    classify_edges_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(src, dst, E, batch, num_real, d_g_node0, d_g_bm_off,
                                                                     d_g_bm_words, plan->bitmap, nullptr, nullptr);
    DA_TRY(cudaGetLastError());
    ring_order_kernel<<<B, 256, sizeof(uint32_t) * 2 * (size_t)words_max, s>>>(plan->bitmap, d_g_node0, d_g_n, d_g_bm_off, d_g_bm_words,
                                                                             d_g_reorder, plan->ext_of_int, d_int_of_ext);
    DA_TRY(cudaGetLastError());
    DA_TRY(cudaMemsetAsync(plan->bitmap, 0, sizeof(uint32_t) * words, s));
    plan->n_reordered_graphs = n_reorder;
  }
  if (plan->n_tiles > 0 && E > 0) {
    DA_TRY(tmp_alloc(&flag, (size_t)E, s));
    classify_edges_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(src, dst, E, batch, num_real, d_g_node0, d_g_bm_off,
                                                                     d_g_bm_words, plan->bitmap, flag, d_int_of_ext);
    DA_TRY(cudaGetLastError());
    DA_TRY(tmp_alloc(&res_src, sizeof(int64_t) * (size_t)E, s));
    DA_TRY(tmp_alloc(&res_dst, sizeof(int64_t) * (size_t)E, s));
    DA_TRY(tmp_alloc(&d_nsel, sizeof(int64_t), s));
    DA_TRY(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, src, flag, res_src, d_nsel, (int)E, s));
    DA_TRY(cub::DeviceSelect::Flagged(nullptr, tb2, dst, flag, res_dst, d_nsel, (int)E, s));
    if (tb2 > tmp_bytes) tmp_bytes = tb2;
    DA_TRY(tmp_alloc(&tmp, tmp_bytes, s));
    DA_TRY(cub::DeviceSelect::Flagged(tmp, tmp_bytes, src, flag, res_src, d_nsel, (int)E, s));
    DA_TRY(cub::DeviceSelect::Flagged(tmp, tmp_bytes, dst, flag, res_dst, d_nsel, (int)E, s));
    DA_TRY(cudaMemcpyAsync(&nsel, d_nsel, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    DA_TRY(cudaStreamSynchronize(s));
    n_res = nsel;
    if (d_int_of_ext && n_res > 0) {   // the residual CSR is in internal numbering too
      relabel_kernel<<<(unsigned)((n_res + 255) / 256), 256, 0, s>>>(res_src, n_res, d_int_of_ext, num_real);
      relabel_kernel<<<(unsigned)((n_res + 255) / 256), 256, 0, s>>>(res_dst, n_res, d_int_of_ext, num_real);
      DA_TRY(cudaGetLastError());
    }
    plan->n_dense_edges = E - n_res;
    ce = build_csr_compressed(res_src, res_dst, n_res, num_total, &plan->residual, s, err);
    if (ce != cudaSuccess) goto fail;
  } else {
    plan->n_dense_edges = 0;
    ce = build_csr_compressed(src, dst, E, num_total, &plan->residual, s, err);
    if (ce != cudaSuccess) goto fail;
  }
  DA_TRY(cudaStreamSynchronize(s));
  // (the pass walks the residual on the host: skipped when that is not "small", i.e. a mostly sparse batch)
  if (plan->n_tiles > 0 && plan->residual.E > 0 && plan->residual.E <= (int64_t)8 << 20 && !getenv("DA_NO_PROMOTE")) {
    // ---- promotion pass (host): residual in-edges of dense-tile rows with multiplicity one move into the bitmap
    // on extra columns (see DensePlan); the residual CSR is rewritten without them
    const int64_t Er = plan->residual.E;
    std::vector<int32_t> rp((size_t)num_total + 1), colv((size_t)Er), new_rp((size_t)num_total + 1), new_col, xs, xslot;
    std::vector<float> wv((size_t)Er), new_w;
    std::vector<int64_t> bword;
    std::vector<uint32_t> bbit;
    DA_TRY(copy_sync(rp.data(), plan->residual.rowptr, sizeof(int32_t) * rp.size(), cudaMemcpyDeviceToHost, s));
    DA_TRY(copy_sync(colv.data(), plan->residual.col, sizeof(int32_t) * (size_t)Er, cudaMemcpyDeviceToHost, s));
    DA_TRY(copy_sync(wv.data(), plan->residual.weight, sizeof(float) * (size_t)Er, cudaMemcpyDeviceToHost, s));
    new_col.reserve((size_t)Er); new_w.reserve((size_t)Er);
    std::vector<std::vector<int32_t>> xsrc_of((size_t)B);   // per graph: sources in extra-column order
    for (int i = 0; i < num_total; ++i) {
      new_rp[i] = (int32_t)new_col.size();
      const int g = (i < num_real && B > 0) ? (int)hbatch[i] : -1;
      const bool dense_row = g >= 0 && g_bm_off[g] >= 0;
      for (int e = rp[i]; e < rp[i + 1]; ++e) {
        bool promoted = false;
        if (dense_row && wv[e] == 1.0f) {
          std::vector<int32_t>& xv = xsrc_of[g];
          const int n = g_n[g], cap = (int)(g_bm_words[g] * 32) - n;
          int idx = -1;
          for (size_t k = 0; k < xv.size(); ++k) if (xv[k] == colv[e]) { idx = (int)k; break; }
          if (idx < 0 && (int)xv.size() < cap) { idx = (int)xv.size(); xv.push_back(colv[e]); }
          if (idx >= 0) {
            const int colx = n + idx;
            bword.push_back(g_bm_off[g] + (int64_t)(i - g_node0[g]) * g_bm_words[g] + (colx >> 5));
            bbit.push_back(1u << (colx & 31));
            promoted = true;
          }
        }
        if (!promoted) { new_col.push_back(colv[e]); new_w.push_back(wv[e]); }
      }
    }
    new_rp[num_total] = (int32_t)new_col.size();
    if (!bword.empty()) {
      for (int g = 0; g < B; ++g) {
        if (xsrc_of[g].empty()) continue;
        const int n = g_n[g];
        for (size_t k = 0; k < xsrc_of[g].size(); ++k) { xs.push_back(xsrc_of[g][k]); xslot.push_back(g_tile0[g] * 128 + n + (int)k); }
        const int ntile = g_bm_words[g] / 4;
        for (int t = 0; t < ntile; ++t) tiles[g_tile0[g] + t].gn = n + (int)xsrc_of[g].size();
      }
      int64_t* d_w = nullptr; uint32_t* d_b = nullptr;
      DA_TRY(tmp_alloc(&d_w, sizeof(int64_t) * bword.size(), s));
      DA_TRY(tmp_alloc(&d_b, sizeof(uint32_t) * bbit.size(), s));
      DA_TRY(cudaMemcpyAsync(d_w, bword.data(), sizeof(int64_t) * bword.size(), cudaMemcpyHostToDevice, s));
      DA_TRY(cudaMemcpyAsync(d_b, bbit.data(), sizeof(uint32_t) * bbit.size(), cudaMemcpyHostToDevice, s));
      set_bits_kernel<<<(unsigned)((bword.size() + 255) / 256), 256, 0, s>>>(d_w, d_b, (int)bword.size(), plan->bitmap);
      DA_TRY(cudaGetLastError());
      DA_TRY(cudaStreamSynchronize(s));
      tmp_free(d_w, s); tmp_free(d_b, s);
      plan->n_extra = (int)xs.size();
      plan->n_promoted_edges = (int64_t)bword.size();
      DA_TRY(tmp_alloc(&plan->x_src, sizeof(int32_t) * xs.size(), s));
      DA_TRY(tmp_alloc(&plan->x_slot, sizeof(int32_t) * xs.size(), s));
      DA_TRY(copy_sync(plan->x_src, xs.data(), sizeof(int32_t) * xs.size(), cudaMemcpyHostToDevice, s));
      DA_TRY(copy_sync(plan->x_slot, xslot.data(), sizeof(int32_t) * xs.size(), cudaMemcpyHostToDevice, s));
      // the rewritten residual is never larger: overwrite in place
      DA_TRY(copy_sync(plan->residual.rowptr, new_rp.data(), sizeof(int32_t) * new_rp.size(), cudaMemcpyHostToDevice, s));
      if (!new_col.empty()) {
        DA_TRY(copy_sync(plan->residual.col, new_col.data(), sizeof(int32_t) * new_col.size(), cudaMemcpyHostToDevice, s));
        DA_TRY(copy_sync(plan->residual.weight, new_w.data(), sizeof(float) * new_w.size(), cudaMemcpyHostToDevice, s));
      }
      plan->residual.E = (int64_t)new_col.size();
      plan->n_dense_edges += plan->n_promoted_edges;
    }
    extra_sources = xs;
  }
  if (plan->n_tiles > 0) {
    int max_blocks = 1;
    for (TileInfo& t : tiles) { t.n_list = 0; t.list_off = 0; max_blocks = std::max(max_blocks, (t.gn + 63) / 64); }
    DA_TRY(copy_sync(plan->tiles, tiles.data(), sizeof(TileInfo) * tiles.size(), cudaMemcpyHostToDevice, s));
    // per-tile lists of the source blocks that hold at least one edge (the final bitmap: in-graph + promoted bits)
    plan->max_blocks = max_blocks;
    unsigned long long hc[3] = {0, 0, 0};
    DA_TRY(tmp_alloc(&plan->blk_list, sizeof(uint16_t) * tiles.size() * (size_t)max_blocks, s));
    DA_TRY(tmp_alloc(&d_counters, sizeof(hc), s));
    DA_TRY(cudaMemsetAsync(d_counters, 0, sizeof(hc), s));
    tile_blocks_kernel<<<plan->n_tiles, 128, 0, s>>>(plan->tiles, plan->bitmap, plan->blk_list, max_blocks, d_counters);
    DA_TRY(cudaGetLastError());
    DA_TRY(copy_sync(hc, d_counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
    plan->n_blocks_total = (int64_t)hc[0]; plan->n_blocks_listed = (int64_t)hc[1]; plan->n_blocks_full = (int64_t)hc[2];
  }
  {  // degree classes of the residual CSR (node order kept, real nodes before virtual rows)
    std::vector<int32_t> rp((size_t)num_total + 1), light, heavy, light_nf;
    std::vector<uint8_t> fused((size_t)num_total, 0);
    DA_TRY(copy_sync(rp.data(), plan->residual.rowptr, sizeof(int32_t) * rp.size(), cudaMemcpyDeviceToHost, s));
    plan->n_fused = 0;
    plan->real_rows_clean = plan->n_tiles > 0;
    for (int i = 0; i < num_real; ++i)
      if (node_slot[i] < 0 || rp[i + 1] != rp[i]) { plan->real_rows_clean = false; break; }
    for (int i = 0; i < num_total; ++i) {
      if (i == num_real) {
        plan->n_light_real = (int)light.size(); plan->n_heavy_real = (int)heavy.size();
        plan->n_light_nf_real = (int)light_nf.size();
      }
      const int deg = rp[i + 1] - rp[i];
      if (deg <= HEAVY_MIN_DEG) {
        light.push_back(i);
        if (node_slot[i] >= 0 && deg <= DA_FUSE_MAX_RESIDUAL) { fused[i] = 1; ++plan->n_fused; }
        else light_nf.push_back(i);
      } else heavy.push_back(i);
    }
    if (num_total == num_real) {
      plan->n_light_real = (int)light.size(); plan->n_heavy_real = (int)heavy.size();
      plan->n_light_nf_real = (int)light_nf.size();
    }
    plan->n_light = (int)light.size(); plan->n_heavy = (int)heavy.size(); plan->n_light_nf = (int)light_nf.size();
    plan->csr_rows_independent = true;
    for (int32_t i : heavy) if (node_slot[i] >= 0) plan->csr_rows_independent = false;
    for (int32_t i : light_nf) if (node_slot[i] >= 0) plan->csr_rows_independent = false;
    {
      std::vector<int32_t> rows(heavy);   // long rows first
      rows.insert(rows.end(), light_nf.begin(), light_nf.end());
      plan->n_csr_rows = (int)rows.size();
      plan->n_csr_rows_real = 0;
      for (int32_t i : rows) if (i < num_real) plan->n_csr_rows_real++;
      DA_TRY(tmp_alloc(&plan->csr_rows, sizeof(int32_t) * (rows.size() + 1), s));
      DA_TRY(copy_sync(plan->csr_rows, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, s));
    }
    DA_TRY(tmp_alloc(&plan->row_fused, fused.size() + 1, s));
    DA_TRY(copy_sync(plan->row_fused, fused.data(), fused.size(), cudaMemcpyHostToDevice, s));
    DA_TRY(tmp_alloc(&plan->light_nf, sizeof(int32_t) * (light_nf.size() + 1), s));
    DA_TRY(copy_sync(plan->light_nf, light_nf.data(), sizeof(int32_t) * light_nf.size(), cudaMemcpyHostToDevice, s));
    {  // which 128-row tiles contain rows whose fp32 Q / K / V some CSR kernel reads?
      std::vector<int32_t> colv((size_t)(plan->residual.E > 0 ? plan->residual.E : 1));
      if (plan->residual.E > 0)
        DA_TRY(copy_sync(colv.data(), plan->residual.col, sizeof(int32_t) * (size_t)plan->residual.E, cudaMemcpyDeviceToHost, s));
      const int n_mt = (num_total + 127) / 128;
      for (int v = 0; v < 2; ++v) {
        std::vector<uint8_t> fl((size_t)n_mt, 0);
        const int n_targets = v == 0 ? num_total : num_real;
        for (int i = 0; i < n_targets; ++i) {
          // fused rows take Q from TMEM inside the dense kernel: only the CSR kernels read fp32 Q
          if (rp[i + 1] > rp[i] && !fused[i]) fl[i >> 7] |= 1;
          // heavy rows read dense-tile sources from the K / V operand images, light rows from the fp32 row
          const bool heavy_row = rp[i + 1] - rp[i] > HEAVY_MIN_DEG;
          for (int e = rp[i]; e < rp[i + 1]; ++e)
            if (!(heavy_row && node_slot[colv[e]] >= 0)) fl[colv[e] >> 7] |= 2;
        }
        for (int32_t x : extra_sources) fl[x >> 7] |= 2;   // the per-layer gather reads their fp32 K / V rows
        for (int i = 0; i < num_total; ++i)
          if (node_slot[i] < 0) fl[i >> 7] = 3;   // rows outside the dense tiles are served by the CSR kernels only
        DA_TRY(tmp_alloc(&plan->f32_tile_flags[v], (size_t)n_mt, s));
        DA_TRY(copy_sync(plan->f32_tile_flags[v], fl.data(), (size_t)n_mt, cudaMemcpyHostToDevice, s));
      }
    }
    DA_TRY(tmp_alloc(&plan->light, sizeof(int32_t) * (light.size() + 1), s));
    DA_TRY(tmp_alloc(&plan->heavy, sizeof(int32_t) * (heavy.size() + 1), s));
    DA_TRY(copy_sync(plan->light, light.data(), sizeof(int32_t) * light.size(), cudaMemcpyHostToDevice, s));
    DA_TRY(copy_sync(plan->heavy, heavy.data(), sizeof(int32_t) * heavy.size(), cudaMemcpyHostToDevice, s));
  }
  tmp_free(dcnt, s); tmp_free(d_g_node0, s); tmp_free(d_g_bm_words, s); tmp_free(d_bad, s); tmp_free(d_g_bm_off, s);
  tmp_free(res_src, s); tmp_free(res_dst, s); tmp_free(d_nsel, s); tmp_free(flag, s); tmp_free(tmp, s);
  tmp_free(d_int_of_ext, s); tmp_free(d_g_n, s); tmp_free(d_g_reorder, s); tmp_free(d_counters, s);
  return cudaSuccess;
fail:
  tmp_free(dcnt, s); tmp_free(d_g_node0, s); tmp_free(d_g_bm_words, s); tmp_free(d_bad, s); tmp_free(d_g_bm_off, s);
  tmp_free(res_src, s); tmp_free(res_dst, s); tmp_free(d_nsel, s); tmp_free(flag, s); tmp_free(tmp, s);
  tmp_free(d_int_of_ext, s); tmp_free(d_g_n, s); tmp_free(d_g_reorder, s); tmp_free(d_counters, s);
  free_plan(plan, s);
  return ce;
#undef DA_TRY
}

}  // namespace da
