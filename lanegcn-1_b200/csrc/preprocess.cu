// preprocess.cu — graph construction right before the forward path, on the GPU:
//   * scale-0 pre / suc edge lists from the lane topology (data.py:272-295): in-lane chains plus the lane-boundary links,
//     in exactly the reference's order (per lane: chain, then one link per predecessor / successor pair);
//   * left / right node edges of `preprocess()` (preprocess_data.py:287-392): for every node the NEAREST node among the
//     lanes reachable as "left (right) neighbour, or a predecessor / successor of that neighbour", kept if closer than
//     cross_dist and heading within 45 degrees.  The reference builds the dense N x N distance matrix (O(N^2) memory:
//     40 GB at 100 k nodes); here one warp scans a node's candidates, O(N) memory.
// Integer outputs are bit-exact with the reference run on CPU (tests/golden/make_golden.py): distances use the same
// separately rounded fp32 operations as torch (sub, mul, add, sqrt), ties go to the lower index like torch.min on CPU.
#include "common.cuh"

namespace {

// first node of every lane from the ascending node -> lane map: lane_start[l] = min{n : lane[n] >= l}, [n_lanes + 1]
__global__ void k_lane_start(const int64_t* __restrict__ lane, int64_t n_nodes, int64_t n_lanes, int32_t* __restrict__ lane_start) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n > n_nodes) return;
  const int64_t cur = n < n_nodes ? lane[n] : n_lanes, prev = n > 0 ? lane[n - 1] : -1;
  for (int64_t l = prev + 1; l <= cur && l <= n_lanes; ++l) lane_start[l] = (int32_t)n;
}

__device__ __forceinline__ int64_t lower_bound_first(const int64_t* __restrict__ pairs, int64_t n_pairs, int64_t key) {
  int64_t lo = 0, hi = n_pairs;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (pairs[2 * mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// data.py:272-295.  PRE: u = idcs[1:], v = idcs[:-1], then per pair (i, j): u = first(i), v = last(j).
//                   SUC: u = idcs[:-1], v = idcs[1:], then per pair (i, j): u = last(i), v = first(j).
// pairs [P,2] int64 sorted by their first column (the reference appends them lane by lane).
template <bool SUC>
__global__ void k_scale0(const int64_t* __restrict__ lane, const int32_t* __restrict__ lane_start, int64_t n_nodes,
                         const int64_t* __restrict__ pairs, int64_t n_pairs, int64_t* __restrict__ u, int64_t* __restrict__ v,
                         int32_t* __restrict__ err) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_nodes) {
    const int64_t i = lane[t];
    if (t != lane_start[i]) {   // not the first node of its lane: the chain edge that ends / starts here
      const int64_t pos = t - i - 1 + lower_bound_first(pairs, n_pairs, i);
      u[pos] = SUC ? t - 1 : t;
      v[pos] = SUC ? t : t - 1;
    }
  } else if (t < n_nodes + n_pairs) {
    const int64_t p = t - n_nodes, i = pairs[2 * p], j = pairs[2 * p + 1];
    if (p > 0 && pairs[2 * (p - 1)] > i) atomicExch(err, 1);   // not sorted by lane
    const int64_t pos = lane_start[i + 1] - (i + 1) + p;
    u[pos] = SUC ? lane_start[i + 1] - 1 : lane_start[i];
    v[pos] = SUC ? lane_start[j] : lane_start[j + 1] - 1;
  }
}

// reach[a][b] (bitmap, n_lanes x words) = side[a][b] | exists c: side[a][c] & (pre[c][b] | suc[c][b])
// (preprocess_data.py:318-321: (mat.pre + mat.suc + mat) > 0.5).  One thread per side pair.
__global__ void k_lane_reach(const int64_t* __restrict__ side, int64_t n_side, const int64_t* __restrict__ pre, int64_t n_pre,
                             const int64_t* __restrict__ suc, int64_t n_suc, int64_t words, uint32_t* __restrict__ reach) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_side) return;
  const int64_t a = side[2 * p], c = side[2 * p + 1];
  uint32_t* row = reach + a * words;
  atomicOr(row + (c >> 5), 1u << (c & 31));
  for (int64_t q = 0; q < n_pre; ++q)
    if (pre[2 * q] == c) atomicOr(row + (pre[2 * q + 1] >> 5), 1u << (pre[2 * q + 1] & 31));
  for (int64_t q = 0; q < n_suc; ++q)
    if (suc[2 * q] == c) atomicOr(row + (suc[2 * q + 1] >> 5), 1u << (suc[2 * q + 1] & 31));
}

// One warp per node i: nearest node j with reach[lane i][lane j] (ties: lowest j), then the reference's two filters.
// keep[i] = 1 and nearest[i] = j when the edge (u = i, v = j) exists.
__global__ void __launch_bounds__(256)
k_side_nearest(const float2* __restrict__ ctrs, const float2* __restrict__ feats, const int64_t* __restrict__ lane,
               const uint32_t* __restrict__ reach, int64_t words, int64_t n_nodes, float cross_dist,
               int32_t* __restrict__ keep, int32_t* __restrict__ nearest) {
  const int lane_id = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_nodes) return;
  const float2 c = ctrs[i];
  const uint32_t* row = reach + lane[i] * words;
  float best = 1e6f;       // the reference writes 1e6 into the masked entries and takes the row minimum
  int64_t best_j = 0;      // argmin of an all-1e6 row is column 0 (it fails the distance test anyway)
  for (int64_t j = lane_id; j < n_nodes; j += 32) {
    const int64_t lj = lane[j];
    float d = 1e6f;
    if (row[lj >> 5] >> (lj & 31) & 1u) {
      const float2 o = ctrs[j];
      const float dx = __fsub_rn(c.x, o.x), dy = __fsub_rn(c.y, o.y);
      d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    }
    if (d < best) {   // strictly smaller: within a lane's stride the lower j was seen first
      best = d;
      best_j = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int64_t oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    if (ob < best || (ob == best && oj < best_j)) {
      best = ob;
      best_j = oj;
    }
  }
  if (lane_id == 0) {
    int ok = best < cross_dist;
    if (ok) {   // |heading(i) - heading(j)| (wrapped into [0, pi]) < pi / 4          preprocess_data.py:337-344
      const float2 f1 = feats[i], f2 = feats[best_j];
      const float pi = 3.14159274101257324f;   // float32(np.pi): the tensors are fp32, the python scalars are cast
      float dt = fabsf(__fsub_rn(atan2f(f1.y, f1.x), atan2f(f2.y, f2.x)));
      if (dt > pi) dt = fabsf(__fsub_rn(dt, 6.28318548202514648f));
      ok = dt < 0.785398185253143311f;
    }
    keep[i] = ok;
    nearest[i] = (int32_t)best_j;
  }
}

__global__ void k_side_emit(const int32_t* __restrict__ keep, const int32_t* __restrict__ pos, const int32_t* __restrict__ nearest,
                            int64_t n_nodes, int64_t* __restrict__ u, int64_t* __restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_nodes && keep[i]) {
    u[pos[i]] = i;
    v[pos[i]] = nearest[i];
  }
}

}  // namespace

extern "C" int64_t lgcn_scale0_workspace_bytes(int64_t n_lanes) { return lgcn_align_up(4 * (n_lanes + 2), 256) + 256; }

extern "C" int lgcn_scale0_edges(const int64_t* lane_idcs, int64_t n_nodes, int64_t n_lanes, const int64_t* pairs,
                                 int64_t n_pairs, int is_suc, int64_t* u, int64_t* v, void* workspace, int32_t* err_flag,
                                 void* stream) {
  LGCN_CHECK_ARG(n_nodes >= 0 && n_lanes >= 0 && n_pairs >= 0, "scale0_edges: negative size");
  if (n_nodes == 0) return 0;
  LGCN_CHECK_ARG(lane_idcs && u && v && workspace && err_flag && (n_pairs == 0 || pairs), "scale0_edges: NULL argument");
  LGCN_CHECK_ARG(n_nodes < ((int64_t)1 << 31), "scale0_edges: node count exceeds int32");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* lane_start = (int32_t*)workspace;
  if (int rc = lgcn_zero_async(err_flag, 4, st)) return rc;
  k_lane_start<<<lgcn_cdiv(n_nodes + 1, 256), 256, 0, st>>>(lane_idcs, n_nodes, n_lanes, lane_start);
  LGCN_LAUNCH_OK();
  if (is_suc)
    k_scale0<true><<<lgcn_cdiv(n_nodes + n_pairs, 256), 256, 0, st>>>(lane_idcs, lane_start, n_nodes, pairs, n_pairs, u, v, err_flag);
  else
    k_scale0<false><<<lgcn_cdiv(n_nodes + n_pairs, 256), 256, 0, st>>>(lane_idcs, lane_start, n_nodes, pairs, n_pairs, u, v, err_flag);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int64_t lgcn_side_edges_workspace_bytes(int64_t n_nodes, int64_t n_lanes) {
  const int64_t words = (n_lanes + 31) / 32;
  return lgcn_align_up(4 * n_lanes * words, 256) + 3 * lgcn_align_up(4 * (n_nodes + 1), 256) + 4352 + 256;
}

extern "C" int lgcn_side_edges(const float* ctrs, const float* feats, const int64_t* lane_idcs, int64_t n_nodes,
                               int64_t n_lanes, const int64_t* side_pairs, int64_t n_side, const int64_t* pre_pairs,
                               int64_t n_pre, const int64_t* suc_pairs, int64_t n_suc, float cross_dist, int64_t* u,
                               int64_t* v, void* workspace, int64_t* h_count, void* stream) {
  LGCN_CHECK_ARG(n_nodes >= 0 && n_lanes >= 0 && n_side >= 0 && n_pre >= 0 && n_suc >= 0, "side_edges: negative size");
  LGCN_CHECK_ARG(h_count, "side_edges: NULL h_count");
  *h_count = 0;
  if (n_nodes == 0 || n_side == 0) return 0;   // `if len(pairs) > 0` (preprocess_data.py:316): no pairs, no edges
  LGCN_CHECK_ARG(ctrs && feats && lane_idcs && side_pairs && u && v && workspace, "side_edges: NULL argument");
  LGCN_CHECK_ARG(n_nodes < ((int64_t)1 << 31), "side_edges: node count exceeds int32");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t words = (n_lanes + 31) / 32;
  char* w = (char*)workspace;
  uint32_t* reach = (uint32_t*)w;
  w += lgcn_align_up(4 * n_lanes * words, 256);
  int32_t* keep = (int32_t*)w;
  w += lgcn_align_up(4 * (n_nodes + 1), 256);
  int32_t* nearest = (int32_t*)w;
  w += lgcn_align_up(4 * (n_nodes + 1), 256);
  int32_t* pos = (int32_t*)w;
  w += lgcn_align_up(4 * (n_nodes + 1), 256);
  int32_t* scratch = (int32_t*)w;
  if (int rc = lgcn_zero_async(reach, 4 * n_lanes * words, st)) return rc;
  k_lane_reach<<<lgcn_cdiv(n_side, 128), 128, 0, st>>>(side_pairs, n_side, pre_pairs, n_pre, suc_pairs, n_suc, words, reach);
  LGCN_LAUNCH_OK();
  k_side_nearest<<<lgcn_cdiv(n_nodes, 8), 256, 0, st>>>((const float2*)ctrs, (const float2*)feats, lane_idcs, reach, words,
                                                       n_nodes, cross_dist, keep, nearest);
  LGCN_LAUNCH_OK();
  if (lgcn_launch_exclusive_scan(keep, pos, n_nodes, nullptr, scratch, st)) return -2;
  k_side_emit<<<lgcn_cdiv(n_nodes, 256), 256, 0, st>>>(keep, pos, nearest, n_nodes, u, v);
  LGCN_LAUNCH_OK();
  int32_t total = 0;
  LGCN_CUDA_OK(cudaMemcpyAsync(&total, pos + n_nodes, 4, cudaMemcpyDeviceToHost, st));
  LGCN_CUDA_OK(cudaStreamSynchronize(st));
  *h_count = total;
  return 0;
}
