// graph.cu — batching of per-scene graphs and the destination-sorted CSR the LaneConv gather consumes.
// Replaces utils.to_long (utils.py:88-96), graph_gather's offset/cat loops (lanegcn.py:191-208) and turns
// the order-dependent index_add_ scatters of lanegcn.py:333-354 into a fixed-order CSR.
#include "common.cuh"

// ------------------------------------------------------------------ index widening + scene offsets
template <typename T>
__global__ void k_offset_indices(const T* __restrict__ local, const int64_t* __restrict__ seg_start,
                                 const int64_t* __restrict__ seg_add, int n_segments, int64_t* __restrict__ out) {
  // one warp per segment; CTAs of 16 warps (see kIndexThreads in common.cuh)
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_segments) return;
  const int64_t beg = seg_start[s], end = seg_start[s + 1], add = seg_add[s];
  for (int64_t i = beg + (threadIdx.x & 31); i < end; i += 32) out[i] = (int64_t)local[i] + add;
}

extern "C" int lgcn_offset_indices(const void* local, int idx_bytes, const int64_t* seg_start,
                                   const int64_t* seg_add, int n_segments, int64_t total, int64_t* out,
                                   void* stream) {
  LGCN_CHECK_ARG(idx_bytes == 2 || idx_bytes == 4 || idx_bytes == 8, "offset_indices: idx_bytes %d", idx_bytes);
  LGCN_CHECK_ARG(n_segments >= 0 && total >= 0, "offset_indices: negative size");
  if (n_segments == 0 || total == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bytes == 2)
    k_offset_indices<int16_t><<<lgcn_cdiv(n_segments, kIndexThreads / 32), kIndexThreads, 0, st>>>((const int16_t*)local, seg_start, seg_add, n_segments, out);
  else if (idx_bytes == 4)
    k_offset_indices<int32_t><<<lgcn_cdiv(n_segments, kIndexThreads / 32), kIndexThreads, 0, st>>>((const int32_t*)local, seg_start, seg_add, n_segments, out);
  else
    k_offset_indices<int64_t><<<lgcn_cdiv(n_segments, kIndexThreads / 32), kIndexThreads, 0, st>>>((const int64_t*)local, seg_start, seg_add, n_segments, out);
  LGCN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ zero fill
__global__ void k_zero(uint32_t* __restrict__ p, int64_t n_words) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0u;
}
int lgcn_zero_async(void* p, int64_t bytes, cudaStream_t st) {
  if (bytes <= 0) return 0;
  LGCN_CHECK_ARG(bytes % 4 == 0 && ((uintptr_t)p & 3) == 0, "zero_async: unaligned");
  k_zero<<<min(lgcn_cdiv(bytes / 4, 256), 148u * 4u), 256, 0, st>>>((uint32_t*)p, bytes / 4);
  LGCN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ meta = cat(turn, control, intersect)
__global__ void k_pack_meta(const float2* __restrict__ turn, const float* __restrict__ control,
                            const float* __restrict__ intersect, float4* __restrict__ meta, int64_t n_cap,
                            const int32_t* __restrict__ n_dev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lgcn_devn(n_dev, n_cap)) return;
  const float2 t = turn[i];
  meta[i] = make_float4(t.x, t.y, control[i], intersect[i]);
}

int lgcn_launch_pack_meta(const float* turn, const float* control, const float* intersect, float* meta, int64_t n_cap,
                          const int32_t* n_dev, cudaStream_t st) {
  if (n_cap <= 0) return 0;
  k_pack_meta<<<lgcn_cdiv(n_cap, 256), 256, 0, st>>>((const float2*)turn, control, intersect, (float4*)meta, n_cap, n_dev);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_pack_meta(const float* turn, const float* control, const float* intersect, float* meta,
                              int64_t n, void* stream) {
  return lgcn_launch_pack_meta(turn, control, intersect, meta, n, nullptr, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ actor_gather (lanegcn.py:155-168)
// out[a, c, t] = in[a, t, c]: the per-actor transpose of the concatenated [A, T, C] history (T = 20 steps, C = 3:
// dx, dy, valid) into the channels-first layout ActorNet's Conv1d stack expects.  Pure data movement (bit-exact).
__global__ void k_actor_transpose(const float* __restrict__ in, float* __restrict__ out, int64_t n_cap,
                                  const int32_t* __restrict__ n_dev, int T, int C) {
  const int64_t n = lgcn_devn(n_dev, n_cap);
  const int64_t per = (int64_t)T * C, total = n * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = i / per;
    const int r = (int)(i - a * per), c = r / T, t = r - c * T;   // i indexes the OUTPUT (coalesced stores)
    out[i] = in[a * per + (int64_t)t * C + c];
  }
}

int lgcn_launch_actor_transpose(const float* in, float* out, int64_t n_cap, const int32_t* n_dev, int T, int C,
                                cudaStream_t st) {
  if (n_cap <= 0) return 0;
  k_actor_transpose<<<min(lgcn_cdiv(n_cap * T * C, 256), 148u * 8u), 256, 0, st>>>(in, out, n_cap, n_dev, T, C);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_actor_gather(const float* feats, float* out, int64_t n_actors, int n_steps, int n_channels,
                                 void* stream) {
  LGCN_CHECK_ARG(n_actors >= 0 && n_steps > 0 && n_channels > 0, "actor_gather: sizes");
  LGCN_CHECK_ARG(n_actors == 0 || (feats && out), "actor_gather: NULL argument");
  return lgcn_launch_actor_transpose(feats, out, n_actors, nullptr, n_steps, n_channels, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ exclusive scan
// n is at most a few hundred thousand rows.  Three tiny launches: (1) per-block sums of 4096-element chunks,
// (2) one block scans the <= 1024 chunk sums, (3) every block scans its chunk with its base.  The array is at most
// a few MB, i.e. L2-resident between the passes.  The element count may live in device memory (n_dev): the grids are
// sized by the capacity and blocks past the live range have nothing to do.
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

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block_sums(const int32_t* __restrict__ cnt, int32_t* __restrict__ sums,
                                                                int64_t n_cap, const int32_t* __restrict__ n_dev) {
  __shared__ int32_t wt[32];
  const int64_t n = lgcn_devn(n_dev, n_cap);
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

__global__ void __launch_bounds__(1024) k_scan_sums(int32_t* __restrict__ sums, int nb, int32_t* __restrict__ out,
                                                    int64_t n_cap, const int32_t* __restrict__ n_dev) {
  __shared__ int32_t wt[32];
  const int32_t v = threadIdx.x < nb ? sums[threadIdx.x] : 0;
  int32_t total;
  const int32_t inc = block_inclusive_scan(v, wt, &total);
  if (threadIdx.x < nb) sums[threadIdx.x] = inc - v;  // exclusive base of each chunk
  if (threadIdx.x == 0) out[lgcn_devn(n_dev, n_cap)] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(const int32_t* __restrict__ cnt, const int32_t* __restrict__ sums,
                                                           int32_t* __restrict__ out, int64_t n_cap,
                                                           const int32_t* __restrict__ n_dev) {
  __shared__ int32_t wt[32];
  const int64_t n = lgcn_devn(n_dev, n_cap);
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
int lgcn_launch_exclusive_scan(const int32_t* cnt, int32_t* out, int64_t n_cap, const int32_t* n_dev, int32_t* scratch,
                               cudaStream_t st) {
  const int64_t per = (int64_t)SCAN_BLOCK * SCAN_ITEMS;
  const int64_t nb = (n_cap + per - 1) / per;
  LGCN_CHECK_ARG(nb <= 1024, "exclusive_scan: %lld elements exceed the 4M-element limit", (long long)n_cap);
  if (nb == 0) {
    if (lgcn_zero_async(out, 4, st)) return -2;
    return 0;
  }
  k_scan_block_sums<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(cnt, scratch, n_cap, n_dev);
  LGCN_LAUNCH_OK();
  k_scan_sums<<<1, 1024, 0, st>>>(scratch, (int)nb, out, n_cap, n_dev);
  LGCN_LAUNCH_OK();
  k_scan_apply<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(cnt, scratch, out, n_cap, n_dev);
  LGCN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ merged destination-sorted CSR
// The edge sets either come by value (host arrays of device pointers: lgcn_csr_build) or are described in device
// memory (lgcn_forward: the batched int64 indices and their segment table are produced on the device and their sizes
// never visit the host).
struct EdgeSets {
  const int64_t* u[LGCN_MAX_KEYS];
  const int64_t* v[LGCN_MAX_KEYS];
  int64_t start[LGCN_MAX_KEYS + 1];  // prefix of lengths: edge id e in [start[k], start[k+1]) belongs to key k
  int n_keys;
};

__device__ __forceinline__ int key_of(const EdgeSets& es, int64_t e) {
  // boundaries are ascending: the key is the number of boundaries at or below e (independent loads, no search chain)
  const int n_keys = es.n_keys;
  int k = 0;
#pragma unroll
  for (int q = 1; q < LGCN_MAX_KEYS; ++q) k += (q < n_keys && e >= es.start[q]) ? 1 : 0;
  return k;
}

// Device-side description of the batched edge sets written by lgcn_offset_indices: `e64` holds, per key k, the
// segments u_k (seg_stride of them, one per scene slot) then v_k; seg_start[] are element offsets.
__global__ void k_edge_sets_from_segs(const int64_t* __restrict__ e64, const int64_t* __restrict__ seg_start,
                                      int seg_stride, int n_keys, EdgeSets* __restrict__ es) {
  const int k = threadIdx.x;
  if (k < n_keys) {
    es->u[k] = e64 + seg_start[(int64_t)(2 * k) * seg_stride];
    es->v[k] = e64 + seg_start[(int64_t)(2 * k + 1) * seg_stride];
  }
  if (k <= n_keys) es->start[k] = seg_start[(int64_t)(2 * k) * seg_stride] / 2;
  if (k == 0) es->n_keys = n_keys;
}

template <bool DEV>
__global__ void k_csr_hist(const EdgeSets es_val, const EdgeSets* __restrict__ es_dev, int64_t n_cap,
                           const int32_t* __restrict__ n_dev, int64_t n_src_cap, int32_t* __restrict__ cnt,
                           int32_t* __restrict__ err) {
  const EdgeSets& es = DEV ? *es_dev : es_val;
  const int64_t n_nodes = lgcn_devn(n_dev, n_cap), n_src = n_dev ? n_nodes : n_src_cap;
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

template <bool DEV>
__global__ void k_csr_place(const EdgeSets es_val, const EdgeSets* __restrict__ es_dev, int64_t n_cap,
                            const int32_t* __restrict__ n_dev, int64_t n_src_cap, const int32_t* __restrict__ rowptr,
                            int32_t* __restrict__ cursor, int32_t* __restrict__ slot_edge) {
  const EdgeSets& es = DEV ? *es_dev : es_val;
  const int64_t n_nodes = lgcn_devn(n_dev, n_cap), n_src = n_dev ? n_nodes : n_src_cap;
  const int64_t E = es.start[es.n_keys];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = key_of(es, e);
    const int64_t u = es.u[k][e - es.start[k]], v = es.v[k][e - es.start[k]];
    if (u < 0 || u >= n_nodes || v < 0 || v >= n_src) continue;
    const int32_t pos = atomicAdd(&cursor[u], 1);
    slot_edge[rowptr[u] + pos] = (int32_t)e;
  }
}

// One thread per CSR SLOT: the slot's edge id e gives its row (u) and source (v); its final position inside the row is
// the number of edge ids of that row below e (ids are distinct), i.e. rows come out ordered by edge id == key order,
// then edge-list order: the stable-by-destination order CPU index_add_ accumulates in.  col = v*(K+1) + (k+1).
// The atomics above only decide the scratch order that this pass erases, so the CSR is deterministic.
// (The first version walked a row per thread with an insertion sort: rows with a merge upstream carry up to ~26 entries,
// nearly every warp had one, and the kernel took 124 us stand-alone — 200+ beside ActorNet — for 2.4 M edges.)
template <bool DEV>
__global__ void __launch_bounds__(kIndexThreads)
k_csr_finish(const EdgeSets es_val, const EdgeSets* __restrict__ es_dev, int64_t n_cap,
             const int32_t* __restrict__ n_dev, int plain, const int32_t* __restrict__ rowptr,
             const int32_t* __restrict__ slot_edge, int32_t* __restrict__ col) {
  const EdgeSets& es = DEV ? *es_dev : es_val;
  const int32_t n_slots = rowptr[lgcn_devn(n_dev, n_cap)];
  const int n_keys = es.n_keys;
  const int32_t nb = n_keys + 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = slot_edge[i];
    const int k = key_of(es, e);
    const int64_t off = e - es.start[k];
    const int64_t u = es.u[k][off], v = es.v[k][off];
    const int32_t beg = rowptr[u], end = rowptr[u + 1];
    int32_t rank = 0;
    for (int32_t j = beg; j < end; ++j) rank += slot_edge[j] < (int32_t)e ? 1 : 0;
    col[beg + rank] = plain ? (int32_t)v : (int32_t)(v * nb + (k + 1));
  }
}

extern "C" int64_t lgcn_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  // cnt/cursor int32[n_nodes] + slot_edge int32[n_edges] + scan scratch int32[1025] + device EdgeSets
  return lgcn_align_up(4 * n_nodes, 256) + lgcn_align_up(4 * n_edges, 256) + 4352 + 512 + 256;
}

template <bool DEV>
static int csr_launch(const EdgeSets& es, const EdgeSets* es_dev, int64_t E_cap, int64_t n_cap, const int32_t* n_dev,
                      int64_t n_src, int plain, int32_t* rowptr, int32_t* col, void* workspace, int32_t* err_flag,
                      cudaStream_t st) {
  int32_t* cnt = (int32_t*)workspace;
  int32_t* slot_edge = (int32_t*)((char*)workspace + lgcn_align_up(4 * n_cap, 256));
  int32_t* scan_scratch = (int32_t*)((char*)slot_edge + lgcn_align_up(4 * E_cap, 256));
  if (lgcn_zero_async(cnt, 4 * n_cap, st) || lgcn_zero_async(err_flag, 4, st)) return -2;
  const unsigned eb = E_cap ? min(lgcn_cdiv(E_cap, kIndexThreads), 148u * 8u) : 0u;
  if (eb) {
    k_csr_hist<DEV><<<eb, kIndexThreads, 0, st>>>(es, es_dev, n_cap, n_dev, n_src, cnt, err_flag);
    LGCN_LAUNCH_OK();
  }
  if (lgcn_launch_exclusive_scan(cnt, rowptr, n_cap, n_dev, scan_scratch, st)) return -2;
  if (eb) {
    if (lgcn_zero_async(cnt, 4 * n_cap, st)) return -2;
    k_csr_place<DEV><<<eb, kIndexThreads, 0, st>>>(es, es_dev, n_cap, n_dev, n_src, rowptr, cnt, slot_edge);
    LGCN_LAUNCH_OK();
    k_csr_finish<DEV><<<eb, kIndexThreads, 0, st>>>(es, es_dev, n_cap, n_dev, plain, rowptr, slot_edge, col);
    LGCN_LAUNCH_OK();
  }
  return 0;
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
  return csr_launch<false>(es, nullptr, E, n_nodes, nullptr, n_src, plain, rowptr, col, workspace, err_flag, st);
}

// CSR of the edge sets lgcn_offset_indices wrote (e64 + segment table, seg_stride = scene slots per (key, u|v)):
// node count in device memory, edge capacity E_cap (half of the index capacity).  Same workspace layout.
int lgcn_launch_csr_from_segs(const int64_t* e64, const int64_t* seg_start, int seg_stride, int n_keys, int64_t E_cap,
                              int64_t n_cap, const int32_t* n_dev, int32_t* rowptr, int32_t* col, void* workspace,
                              int32_t* err_flag, cudaStream_t st) {
  LGCN_CHECK_ARG(n_keys >= 0 && n_keys <= LGCN_MAX_KEYS, "csr_from_segs: n_keys %d out of range", n_keys);
  LGCN_CHECK_ARG(E_cap < (int64_t)1 << 31 && n_cap * (int64_t)(n_keys + 1) < (int64_t)1 << 31, "csr_from_segs: sizes exceed int32");
  EdgeSets* es_dev = (EdgeSets*)((char*)workspace + lgcn_align_up(4 * n_cap, 256) + lgcn_align_up(4 * E_cap, 256) + 4352);
  k_edge_sets_from_segs<<<1, 32, 0, st>>>(e64, seg_start, seg_stride, n_keys, es_dev);
  LGCN_LAUNCH_OK();
  EdgeSets dummy;
  dummy.n_keys = 0;
  return csr_launch<true>(dummy, es_dev, E_cap, n_cap, n_dev, n_cap, 0, rowptr, col, workspace, err_flag, st);
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
