// dilate.cu — multi-scale edge dilation on the GPU: the boolean CSR x CSR squaring of data.py:520-534
// (`mat = mat * mat` five times: hops 2,4,8,16,32), bit-exact including scipy's per-row COLUMN ORDER.
//
// scipy's csr_matmat (SMMP) emits, for row i, the columns k in REVERSE first-discovery order, where discovery
// walks j over row i in stored order and k over row j in stored order (restated and pinned in
// oracle/graph_oracle.py::dilate_smmp).  The scale-0 matrix comes from csr_matrix((data,(u,v))): columns ascending,
// duplicates merged.  Works on the BATCHED graph: the adjacency is block diagonal over scenes and so are its
// powers, so one launch sequence dilates every scene of a batch.
//
// One thread per row (rows of lane graphs have 1-3 entries; a row of a power has at most a few dozen), three
// passes per squaring: upper bound per row -> scan -> discover into scratch (linear-search dedup inside the row's
// own scratch segment) -> scan of the true counts -> emit reversed.  Integer only: deterministic and exact.
#include "common.cuh"

namespace {

__global__ void k_d_hist(const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t n_edges, int64_t n_nodes,
                         int32_t* __restrict__ cnt, int32_t* __restrict__ err) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = u[e], b = v[e];
    if (a < 0 || a >= n_nodes || b < 0 || b >= n_nodes) {
      atomicExch(err, 1);
      continue;
    }
    atomicAdd(&cnt[a], 1);
  }
}

__global__ void k_d_place(const int64_t* __restrict__ u, const int64_t* __restrict__ v, int64_t n_edges, int64_t n_nodes,
                          const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor, int32_t* __restrict__ col) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = u[e], b = v[e];
    if (a < 0 || a >= n_nodes || b < 0 || b >= n_nodes) continue;
    col[rowptr[a] + atomicAdd(&cursor[a], 1)] = (int32_t)b;
  }
}

// per row: sort ascending, drop duplicates in place, report the unique count (the atomics above only chose a
// scratch order that the sort erases)
__global__ void k_d_sort_unique(int64_t n_nodes, const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                int32_t* __restrict__ ucnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int32_t beg = rowptr[r], end = rowptr[r + 1];
  for (int32_t i = beg + 1; i < end; ++i) {
    const int32_t x = col[i];
    int32_t j = i - 1;
    while (j >= beg && col[j] > x) {
      col[j + 1] = col[j];
      --j;
    }
    col[j + 1] = x;
  }
  int32_t w = beg;
  for (int32_t i = beg; i < end; ++i)
    if (i == beg || col[i] != col[w - 1]) col[w++] = col[i];
  ucnt[r] = w - beg;
}

__global__ void k_d_compact(int64_t n_nodes, const int32_t* __restrict__ rowptr_raw, const int32_t* __restrict__ col_raw,
                            const int32_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int32_t src = rowptr_raw[r], dst = rowptr[r], n = rowptr[r + 1] - dst;
  for (int32_t i = 0; i < n; ++i) col[dst + i] = col_raw[src + i];
}

// upper bound of row i of A*A: sum over j in row i of deg(j)
__global__ void k_d_bound(int64_t n_nodes, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                          int32_t* __restrict__ ub) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  int32_t s = 0;
  for (int32_t p = rowptr[r]; p < rowptr[r + 1]; ++p) {
    const int32_t j = col[p];
    s += rowptr[j + 1] - rowptr[j];
  }
  ub[r] = s;
}

// first-discovery order of row i of A*A into scratch[off[i]..]; cnt[i] = number of distinct columns
__global__ void k_d_discover(int64_t n_nodes, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                             const int32_t* __restrict__ off, int32_t* __restrict__ scratch, int32_t* __restrict__ cnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  int32_t* mine = scratch + off[r];
  int32_t n = 0;
  for (int32_t p = rowptr[r]; p < rowptr[r + 1]; ++p) {
    const int32_t j = col[p];
    for (int32_t q = rowptr[j]; q < rowptr[j + 1]; ++q) {
      const int32_t k = col[q];
      bool seen = false;
      for (int32_t t = 0; t < n; ++t) seen |= (mine[t] == k);
      if (!seen) mine[n++] = k;
    }
  }
  cnt[r] = n;
}

// emit row i reversed (scipy's linked-list head insertion) + the COO arrays u (= row) and v (= col) as int64
__global__ void k_d_emit(int64_t n_nodes, const int32_t* __restrict__ off, const int32_t* __restrict__ scratch,
                         const int32_t* __restrict__ rowptr_out, int32_t* __restrict__ col_out, int64_t* __restrict__ u_out,
                         int64_t* __restrict__ v_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int32_t dst = rowptr_out[r], n = rowptr_out[r + 1] - dst;
  const int32_t* mine = scratch + off[r];
  for (int32_t i = 0; i < n; ++i) {
    const int32_t k = mine[n - 1 - i];
    col_out[dst + i] = k;
    if (u_out) u_out[dst + i] = r;
    if (v_out) v_out[dst + i] = k;
  }
}

int read_i32(const int32_t* d, int64_t* h, cudaStream_t st) {
  int32_t x = 0;
  LGCN_CUDA_OK(cudaMemcpyAsync(&x, d, 4, cudaMemcpyDeviceToHost, st));
  LGCN_CUDA_OK(cudaStreamSynchronize(st));
  *h = x;
  return 0;
}

}  // namespace

// workspace (int32): a[n+1] | b[n+1] | c[n] | scan scratch[1088] | raw/scratch columns[cap]
extern "C" int64_t lgcn_dilate_workspace_bytes(int64_t n_nodes, int64_t cap) {
  return 4 * (3 * lgcn_align_up(n_nodes + 1, 64) + 1088 + lgcn_align_up(cap, 64)) + 256;
}

extern "C" int lgcn_dilate_csr0(const int64_t* u, const int64_t* v, int64_t n_edges, int64_t n_nodes, int32_t* rowptr,
                                int32_t* col, void* workspace, int64_t* h_nnz, void* stream) {
  LGCN_CHECK_ARG(n_nodes >= 0 && n_edges >= 0 && n_edges < ((int64_t)1 << 31), "dilate_csr0: bad sizes");
  LGCN_CHECK_ARG(h_nnz, "dilate_csr0: h_nnz is required");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t na = lgcn_align_up(n_nodes + 1, 64);
  int32_t* raw_ptr = (int32_t*)workspace;   // rowptr with duplicates
  int32_t* cnt = raw_ptr + na;              // histogram / cursor, then unique counts
  int32_t* err = cnt + na;                  // [0] error flag (uses the c[] region)
  int32_t* scan_scratch = err + na;
  int32_t* raw_col = scan_scratch + 1088;
  LGCN_CUDA_OK(cudaMemsetAsync(cnt, 0, 4 * (size_t)(n_nodes + 1), st));
  LGCN_CUDA_OK(cudaMemsetAsync(err, 0, 4, st));
  const unsigned eb = n_edges ? min(lgcn_cdiv(n_edges, 256), 148u * 16u) : 0u;
  if (eb) {
    k_d_hist<<<eb, 256, 0, st>>>(u, v, n_edges, n_nodes, cnt, err);
    LGCN_LAUNCH_OK();
  }
  if (lgcn_launch_exclusive_scan(cnt, raw_ptr, n_nodes, nullptr, scan_scratch, st)) return -2;
  if (eb) {
    LGCN_CUDA_OK(cudaMemsetAsync(cnt, 0, 4 * (size_t)(n_nodes + 1), st));
    k_d_place<<<eb, 256, 0, st>>>(u, v, n_edges, n_nodes, raw_ptr, cnt, raw_col);
    LGCN_LAUNCH_OK();
  }
  if (n_nodes) {
    k_d_sort_unique<<<lgcn_cdiv(n_nodes, 128), 128, 0, st>>>(n_nodes, raw_ptr, raw_col, cnt);
    LGCN_LAUNCH_OK();
  }
  if (lgcn_launch_exclusive_scan(cnt, rowptr, n_nodes, nullptr, scan_scratch, st)) return -2;
  if (n_nodes) {
    k_d_compact<<<lgcn_cdiv(n_nodes, 128), 128, 0, st>>>(n_nodes, raw_ptr, raw_col, rowptr, col);
    LGCN_LAUNCH_OK();
  }
  int64_t bad = 0;
  if (read_i32(err, &bad, st)) return -2;
  LGCN_CHECK_ARG(bad == 0, "dilate_csr0: edge index out of range [0, n_nodes)");
  return read_i32(rowptr + n_nodes, h_nnz, st);
}

extern "C" int lgcn_dilate_bound(const int32_t* rowptr, const int32_t* col, int64_t n_nodes, void* workspace,
                                 int64_t* h_bound, void* stream) {
  LGCN_CHECK_ARG(h_bound, "dilate_bound: h_bound is required");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t na = lgcn_align_up(n_nodes + 1, 64);
  int32_t* off = (int32_t*)workspace;  // a[]: scratch offsets (kept for lgcn_dilate_square)
  int32_t* ub = off + 2 * na;          // c[]
  int32_t* scan_scratch = ub + na;
  if (n_nodes) {
    k_d_bound<<<lgcn_cdiv(n_nodes, 128), 128, 0, st>>>(n_nodes, rowptr, col, ub);
    LGCN_LAUNCH_OK();
  }
  if (lgcn_launch_exclusive_scan(ub, off, n_nodes, nullptr, scan_scratch, st)) return -2;
  return read_i32(off + n_nodes, h_bound, st);
}

extern "C" int lgcn_dilate_square(const int32_t* rowptr, const int32_t* col, int64_t n_nodes, int64_t cap,
                                  int32_t* rowptr_out, int32_t* col_out, int64_t* u_out, int64_t* v_out, void* workspace,
                                  int64_t* h_nnz, void* stream) {
  LGCN_CHECK_ARG(h_nnz, "dilate_square: h_nnz is required");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t na = lgcn_align_up(n_nodes + 1, 64);
  int32_t* off = (int32_t*)workspace;  // filled by lgcn_dilate_bound on the SAME workspace
  int32_t* cnt = off + na;             // b[]
  int32_t* scan_scratch = off + 3 * na;
  int32_t* scratch = scan_scratch + 1088;
  int64_t bound = 0;
  if (read_i32(off + n_nodes, &bound, st)) return -2;
  LGCN_CHECK_ARG(bound <= cap, "dilate_square: capacity %lld < bound %lld (call lgcn_dilate_bound first)", (long long)cap,
                 (long long)bound);
  if (n_nodes) {
    k_d_discover<<<lgcn_cdiv(n_nodes, 128), 128, 0, st>>>(n_nodes, rowptr, col, off, scratch, cnt);
    LGCN_LAUNCH_OK();
  }
  if (lgcn_launch_exclusive_scan(cnt, rowptr_out, n_nodes, nullptr, scan_scratch, st)) return -2;
  if (n_nodes) {
    k_d_emit<<<lgcn_cdiv(n_nodes, 128), 128, 0, st>>>(n_nodes, off, scratch, rowptr_out, col_out, u_out, v_out);
    LGCN_LAUNCH_OK();
  }
  return read_i32(rowptr_out + n_nodes, h_nnz, st);
}
