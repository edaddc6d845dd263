// graph.cu — batching of per-scene graphs and the destination-sorted CSR the LaneConv gather consumes.
// Replaces utils.to_long (utils.py:88-96), graph_gather's offset/cat loops (lanegcn.py:191-208) and turns
// the order-dependent index_add_ scatters of lanegcn.py:333-354 into a fixed-order CSR.
#include "common.cuh"

// ------------------------------------------------------------------ index widening + scene offsets
template <typename T>
__global__ void k_offset_indices(const T* __restrict__ local, const int64_t* __restrict__ seg_start,
                                 const int64_t* __restrict__ seg_add, int64_t* __restrict__ out) {
  const int s = blockIdx.x;
  const int64_t beg = seg_start[s], end = seg_start[s + 1], add = seg_add[s];
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) out[i] = (int64_t)local[i] + add;
}

extern "C" int lgcn_offset_indices(const void* local, int idx_bytes, const int64_t* seg_start,
                                   const int64_t* seg_add, int n_segments, int64_t total, int64_t* out,
                                   void* stream) {
  LGCN_CHECK_ARG(idx_bytes == 2 || idx_bytes == 4 || idx_bytes == 8, "offset_indices: idx_bytes %d", idx_bytes);
  LGCN_CHECK_ARG(n_segments >= 0 && total >= 0, "offset_indices: negative size");
  if (n_segments == 0 || total == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bytes == 2)
    k_offset_indices<int16_t><<<n_segments, 128, 0, st>>>((const int16_t*)local, seg_start, seg_add, out);
  else if (idx_bytes == 4)
    k_offset_indices<int32_t><<<n_segments, 128, 0, st>>>((const int32_t*)local, seg_start, seg_add, out);
  else
    k_offset_indices<int64_t><<<n_segments, 128, 0, st>>>((const int64_t*)local, seg_start, seg_add, out);
  LGCN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ meta = cat(turn, control, intersect)
__global__ void k_pack_meta(const float2* __restrict__ turn, const float* __restrict__ control,
                            const float* __restrict__ intersect, float4* __restrict__ meta, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 t = turn[i];
  meta[i] = make_float4(t.x, t.y, control[i], intersect[i]);
}

extern "C" int lgcn_pack_meta(const float* turn, const float* control, const float* intersect, float* meta,
                              int64_t n, void* stream) {
  if (n <= 0) return 0;
  k_pack_meta<<<lgcn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)turn, control, intersect,
                                                                   (float4*)meta, n);
  LGCN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ exclusive scan
// n is at most a few hundred thousand rows.  Three tiny launches: (1) per-block sums of 4096-element chunks,
// (2) one block scans the <= 1024 chunk sums, (3) every block scans its chunk with its base.  All loads are
// the array is at most ~1 MB, i.e. L2-resident between the passes.
#define SCAN_BLOCK 256
#define SCAN_ITEMS 16  // per thread -> 4096 elements per block

__device__ __forceinline__ int32_t block_inclusive_scan(int32_t v, int32_t* warp_tot /* [32] */, int32_t* total) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) warp_tot[w] = v;
  __syncthreads();
  if (w == 0) {
    int32_t x = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t u = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += u;
    }
    warp_tot[lane] = x;
  }
  __syncthreads();
  if (total) *total = warp_tot[(blockDim.x >> 5) - 1];
  const int32_t r = v + (w ? warp_tot[w - 1] : 0);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block_sums(const int32_t* __restrict__ cnt, int32_t* __restrict__ sums, int64_t n) {
  __shared__ int32_t wt[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS;
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int64_t j = base + (int64_t)i * SCAN_BLOCK + threadIdx.x;
    if (j < n) s += cnt[j];
  }
  int32_t total;
  block_inclusive_scan(s, wt, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(int32_t* __restrict__ sums, int nb, int32_t* __restrict__ out_total) {
  __shared__ int32_t wt[32];
  const int32_t v = threadIdx.x < nb ? sums[threadIdx.x] : 0;
  int32_t total;
  const int32_t inc = block_inclusive_scan(v, wt, &total);
  if (threadIdx.x < nb) sums[threadIdx.x] = inc - v;  // exclusive base of each chunk
  if (threadIdx.x == 0) *out_total = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(const int32_t* __restrict__ cnt, const int32_t* __restrict__ sums,
                                                           int32_t* __restrict__ out, int64_t n) {
  __shared__ int32_t wt[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS;
  int32_t run = sums[blockIdx.x];
  // thread t owns SCAN_ITEMS CONSECUTIVE elements (strided reads hit L2: the array is at most ~1 MB)
  int32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int64_t j = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
    v[i] = j < n ? cnt[j] : 0;
    s += v[i];
  }
  const int32_t inc = block_inclusive_scan(s, wt, nullptr);
  run += inc - s;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int64_t j = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
    if (j < n) out[j] = run;
    run += v[i];
  }
}

// out[0..n] (n+1 entries, out[n] = total).  `scratch` (>= 1025 int32, caller-provided so the call stays
// re-entrant) holds the chunk sums.
int lgcn_launch_exclusive_scan(const int32_t* cnt, int32_t* out, int64_t n, int32_t* scratch, cudaStream_t st) {
  const int64_t per = (int64_t)SCAN_BLOCK * SCAN_ITEMS;
  const int64_t nb = (n + per - 1) / per;
  LGCN_CHECK_ARG(nb <= 1024, "exclusive_scan: %lld elements exceed the 4M-element limit", (long long)n);
  if (nb == 0) {
    LGCN_CUDA_OK(cudaMemsetAsync(out, 0, 4, st));
    return 0;
  }
  k_scan_block_sums<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(cnt, scratch, n);
  LGCN_LAUNCH_OK();
  k_scan_sums<<<1, 1024, 0, st>>>(scratch, (int)nb, out + n);
  LGCN_LAUNCH_OK();
  k_scan_apply<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(cnt, scratch, out, n);
  LGCN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ merged destination-sorted CSR
struct EdgeSets {
  const int64_t* u[LGCN_MAX_KEYS];
  const int64_t* v[LGCN_MAX_KEYS];
  int64_t start[LGCN_MAX_KEYS + 1];  // prefix of lengths: edge id e in [start[k], start[k+1]) belongs to key k
  int n_keys;
};

__device__ __forceinline__ int key_of(const EdgeSets& es, int64_t e) {
  int k = 0;
#pragma unroll 1
  while (k + 1 < es.n_keys && e >= es.start[k + 1]) ++k;
  return k;
}

__global__ void k_csr_hist(EdgeSets es, int64_t n_nodes, int64_t n_src, int32_t* __restrict__ cnt, int32_t* __restrict__ err) {
  const int64_t E = es.start[es.n_keys];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = key_of(es, e);
    const int64_t u = es.u[k][e - es.start[k]], v = es.v[k][e - es.start[k]];
    if (u < 0 || u >= n_nodes || v < 0 || v >= n_src) {
      atomicExch(err, 1);
      continue;
    }
    atomicAdd(&cnt[u], 1);
  }
}

__global__ void k_csr_place(EdgeSets es, int64_t n_nodes, int64_t n_src, const int32_t* __restrict__ rowptr,
                            int32_t* __restrict__ cursor, int32_t* __restrict__ slot_edge) {
  const int64_t E = es.start[es.n_keys];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = key_of(es, e);
    const int64_t u = es.u[k][e - es.start[k]], v = es.v[k][e - es.start[k]];
    if (u < 0 || u >= n_nodes || v < 0 || v >= n_src) continue;
    const int32_t pos = atomicAdd(&cursor[u], 1);
    slot_edge[rowptr[u] + pos] = (int32_t)e;
  }
}

// One thread per destination row: order the row's edge ids ascending (== key order, then edge-list order:
// the stable-by-destination order CPU index_add_ accumulates in) and emit col = v*(K+1) + (k+1).
// Rows are short (about a dozen entries on lane graphs), so an in-place insertion sort is the right tool;
// the atomics above only decide a scratch order that this pass erases, so the CSR is deterministic.
__global__ void k_csr_finish(EdgeSets es, int64_t n_nodes, int plain, const int32_t* __restrict__ rowptr,
                             int32_t* __restrict__ slot_edge, int32_t* __restrict__ col) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int32_t beg = rowptr[r], end = rowptr[r + 1];
  for (int32_t i = beg + 1; i < end; ++i) {
    const int32_t x = slot_edge[i];
    int32_t j = i - 1;
    while (j >= beg && slot_edge[j] > x) {
      slot_edge[j + 1] = slot_edge[j];
      --j;
    }
    slot_edge[j + 1] = x;
  }
  const int32_t nb = es.n_keys + 1;
  for (int32_t i = beg; i < end; ++i) {
    const int64_t e = slot_edge[i];
    const int k = key_of(es, e);
    const int64_t v = es.v[k][e - es.start[k]];
    col[i] = plain ? (int32_t)v : (int32_t)(v * nb + (k + 1));
  }
}

extern "C" int64_t lgcn_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  // cnt/cursor int32[n_nodes] + slot_edge int32[n_edges] + scan scratch int32[1025]
  return lgcn_align_up(4 * n_nodes, 256) + lgcn_align_up(4 * n_edges, 256) + 4352 + 256;
}

static int csr_build_impl(const int64_t* const* h_u, const int64_t* const* h_v, const int64_t* h_len, int n_keys,
                          int64_t n_nodes, int64_t n_src, int plain, int32_t* rowptr, int32_t* col, void* workspace,
                          int32_t* err_flag, cudaStream_t st) {
  LGCN_CHECK_ARG(n_keys >= 0 && n_keys <= LGCN_MAX_KEYS, "csr_build: n_keys %d out of range", n_keys);
  LGCN_CHECK_ARG(n_nodes >= 0 && n_src >= 0, "csr_build: negative size");
  EdgeSets es;
  es.n_keys = n_keys;
  es.start[0] = 0;
  for (int k = 0; k < n_keys; ++k) {
    LGCN_CHECK_ARG(h_len[k] >= 0, "csr_build: negative length for key %d", k);
    es.u[k] = h_u[k];
    es.v[k] = h_v[k];
    es.start[k + 1] = es.start[k] + h_len[k];
  }
  const int64_t E = es.start[n_keys];
  LGCN_CHECK_ARG(E < (int64_t)1 << 31, "csr_build: %lld edges exceed int32", (long long)E);
  LGCN_CHECK_ARG(n_src * (int64_t)(plain ? 1 : n_keys + 1) < (int64_t)1 << 31, "csr_build: block index exceeds int32");
  int32_t* cnt = (int32_t*)workspace;
  int32_t* slot_edge = (int32_t*)((char*)workspace + lgcn_align_up(4 * n_nodes, 256));
  int32_t* scan_scratch = (int32_t*)((char*)slot_edge + lgcn_align_up(4 * E, 256));
  LGCN_CUDA_OK(cudaMemsetAsync(cnt, 0, 4 * (size_t)n_nodes, st));
  LGCN_CUDA_OK(cudaMemsetAsync(err_flag, 0, 4, st));
  const unsigned eb = E ? min(lgcn_cdiv(E, 256), 148u * 16u) : 0u;
  if (eb) {
    k_csr_hist<<<eb, 256, 0, st>>>(es, n_nodes, n_src, cnt, err_flag);
    LGCN_LAUNCH_OK();
  }
  if (lgcn_launch_exclusive_scan(cnt, rowptr, n_nodes, scan_scratch, st)) return -2;
  if (eb) {
    LGCN_CUDA_OK(cudaMemsetAsync(cnt, 0, 4 * (size_t)n_nodes, st));
    k_csr_place<<<eb, 256, 0, st>>>(es, n_nodes, n_src, rowptr, cnt, slot_edge);
    LGCN_LAUNCH_OK();
    k_csr_finish<<<lgcn_cdiv(n_nodes, 128), 128, 0, st>>>(es, n_nodes, plain, rowptr, slot_edge, col);
    LGCN_LAUNCH_OK();
  }
  return 0;
}

extern "C" int lgcn_csr_build(const int64_t* const* h_u, const int64_t* const* h_v, const int64_t* h_len,
                              int n_keys, int64_t n_nodes, int32_t* rowptr, int32_t* col, void* workspace,
                              int32_t* err_flag, void* stream) {
  return csr_build_impl(h_u, h_v, h_len, n_keys, n_nodes, n_nodes, 0, rowptr, col, workspace, err_flag, (cudaStream_t)stream);
}

extern "C" int lgcn_scatter_csr_build(const int64_t* dst, const int64_t* src, int64_t n_edges, int64_t n_dst,
                                      int64_t n_src, int32_t* rowptr, int32_t* col, void* workspace, int32_t* err_flag,
                                      void* stream) {
  const int64_t* u[1] = {dst};
  const int64_t* v[1] = {src};
  return csr_build_impl(u, v, &n_edges, 1, n_dst, n_src, 1, rowptr, col, workspace, err_flag, (cudaStream_t)stream);
}
