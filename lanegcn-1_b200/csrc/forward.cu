// forward.cu — lgcn_forward: lanegcn.py:134-141 of Net.forward (graph_gather -> MapNet -> A2M -> M2M -> M2A -> A2A) as
// ONE C call that enqueues ~100 kernels with no host synchronisation.  Every batch-dependent row count (nodes,
// actors, pairs of the three Att lists) is read from device memory; buffers and grids are sized by capacities, so the
// sequence is the same for every batch that fits them and can be captured in a CUDA graph once per capacity bucket
// (the reference synchronises once per scene per Att layer: lanegcn.py:680-681).
#include <string.h>

#include <mutex>

#include "common.cuh"

#define CC ((int64_t)LGCN_C * LGCN_C)

// internal entry points of api.cu
LinearArgs lgcn_lin1(const float* x, const int32_t* idx, const float* W, const float* gamma, const float* beta,
                     const float* res, int flags, float* out, int64_t m, const int32_t* m_dev);
int lgcn_att_split_weights(const float* wpack, float* WH, float* WL, cudaStream_t st);
int lgcn_att_layer(const float* agts_in, float* agts_out, const float* ctx, const float* agt_ctrs, const float* ctx_ctrs,
                   const int32_t* hi, const int32_t* wi, const int32_t* rowptr, int64_t n_agt, const int32_t* n_agt_dev,
                   int64_t n_ctx, int64_t n_pairs, const int32_t* n_pairs_dev, const float* wpack, const float* WH,
                   const float* WL, void* workspace, cudaStream_t st, const LgcnFork* fk);
int lgcn_laneconv_stack_presplit(float* feat, float* other, float* xa, void* plan, int64_t n_edges, int n_keys,
                                 int n_blocks, const float* wpack, const float* w_hi, const float* w_lo, int64_t n_nodes,
                                 const int32_t* n_dev, cudaStream_t st);

namespace {

// ---- fork / join events of the optional parallel branches (LgcnForwardArgs.aux_streams): three per device, created on
// first use and kept for the life of the process like the cached function attributes (they carry no data: a record /
// wait pair only orders streams, and inside a capture it only adds graph edges)
constexpr int kMaxDev = 64;
cudaEvent_t g_fork_ev[kMaxDev][3];
bool g_fork_ready[kMaxDev];
std::mutex g_fork_mu;
int fork_of(const LgcnForwardArgs& a, LgcnFork* fk) {
  memset(fk, 0, sizeof(*fk));
  if (!a.aux_streams[0] && !a.aux_streams[1]) return 0;
  int dev = 0;
  LGCN_CUDA_OK(cudaGetDevice(&dev));
  LGCN_CHECK_ARG(dev >= 0 && dev < kMaxDev, "forward: device %d", dev);
  {
    std::lock_guard<std::mutex> lock(g_fork_mu);
    if (!g_fork_ready[dev]) {
      for (int i = 0; i < 3; ++i) LGCN_CUDA_OK(cudaEventCreateWithFlags(&g_fork_ev[dev][i], cudaEventDisableTiming));
      g_fork_ready[dev] = true;
    }
  }
  for (int i = 0; i < 2; ++i) fk->aux[i] = (cudaStream_t)a.aux_streams[i];
  for (int i = 0; i < 3; ++i) fk->ev[i] = g_fork_ev[dev][i];
  return 0;
}

// ---- tf32 hi / lo images of the weights (float offsets into `prepared`)
struct Prepared {
  int64_t in_hi, in_lo, seg_hi, seg_lo, meta_hi, meta_lo, map_hi, map_lo, m2m_hi, m2m_lo, att_hi[6], att_lo[6], total;
};
Prepared prepared_layout(int n_scales) {
  const int n_keys = 2 * n_scales + 2;
  const int64_t stack = lgcn_align_up(4 * lgcn_laneconv_wpack_floats(n_keys), 256);
  Prepared p;
  int64_t o = 0;
  auto take = [&](int64_t n) { const int64_t at = o; o += lgcn_align_up(n, 256); return at; };
  p.in_hi = take(CC); p.in_lo = take(CC);
  p.seg_hi = take(CC); p.seg_lo = take(CC);
  p.meta_hi = take(CC); p.meta_lo = take(CC);
  p.map_hi = take(stack); p.map_lo = take(stack);
  p.m2m_hi = take(stack); p.m2m_lo = take(stack);
  for (int i = 0; i < 6; ++i) {
    p.att_hi[i] = take(8 * CC);
    p.att_lo[i] = take(8 * CC);
  }
  p.total = o;
  return p;
}

// ---- workspace (byte offsets)
struct Layout {
  int64_t e64, meta, rowptr, col, csr_ws, plan, t0, hid, xa, att_ws, total;
  int64_t p_rowptr[3], p_ws[3], p_hi[3], p_wi[3], p_tot;
  int64_t E_cap;
  int n_keys;
};
Layout layout_of(const LgcnForwardArgs& a) {
  Layout L;
  L.n_keys = 2 * a.n_scales + 2;
  L.E_cap = a.cap_index / 2;
  const int64_t N = a.cap_nodes, A = a.cap_actors, rows = N > A ? N : A;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { const int64_t at = o; o += lgcn_align_up(bytes > 0 ? bytes : 1, 1024); return at; };
  L.e64 = take(8 * a.cap_index);
  L.meta = take(16 * N);
  L.rowptr = take(4 * (N + 1));
  L.col = take(4 * L.E_cap);
  L.csr_ws = take(lgcn_csr_workspace_bytes(N, L.E_cap));
  L.plan = take(lgcn_laneconv_plan_bytes(N, L.E_cap, L.n_keys));
  const int64_t n_agt[3] = {N, A, A};
  for (int i = 0; i < 3; ++i) {
    L.p_rowptr[i] = take(4 * (n_agt[i] + 1));
    L.p_ws[i] = take(lgcn_pairs_workspace_bytes(n_agt[i], a.cap_scenes));
    L.p_hi[i] = take(4 * a.cap_pairs[i]);
    L.p_wi[i] = take(4 * a.cap_pairs[i]);
  }
  L.p_tot = take(64);
  L.t0 = take(N * LGCN_C * 4);
  L.hid = take(N * LGCN_C * 4);
  L.xa = take(lgcn_laneconv_fused_aux_bytes(L.E_cap));
  int64_t pmax = a.cap_pairs[0];
  for (int i = 1; i < 3; ++i) pmax = a.cap_pairs[i] > pmax ? a.cap_pairs[i] : pmax;
  L.att_ws = take(lgcn_att_workspace_bytes(rows, pmax));
  L.total = o;
  return L;
}

int check_args(const LgcnForwardArgs* a) {
  LGCN_CHECK_ARG(a, "forward: NULL args");
  LGCN_CHECK_ARG(a->cap_nodes > 0 && a->cap_actors > 0 && a->cap_index >= 0 && a->cap_scenes >= 1, "forward: capacities");
  LGCN_CHECK_ARG(a->n_scales >= 1 && 2 * a->n_scales + 2 <= LGCN_MAX_KEYS, "forward: n_scales %d", a->n_scales);
  LGCN_CHECK_ARG(a->cap_index % 2 == 0, "forward: cap_index must be even (u and v per edge)");
  for (int i = 0; i < 3; ++i) LGCN_CHECK_ARG(a->cap_pairs[i] > 0, "forward: cap_pairs[%d]", i);
  return 0;
}

}  // namespace

// reg[a, k, t, :] <- reg[a, k, t, :] . rot[b] + orig[b], b = scene of actor a: the per-scene loop of lanegcn.py:145-150
// (torch.matmul(reg[i], rot[i]) + orig[i]) over the batched tensor, in place.
__global__ void k_world_transform(float2* __restrict__ reg, const int32_t* __restrict__ actor_off, int n_scenes,
                                  const float* __restrict__ rot, const float* __restrict__ orig, int64_t a_cap,
                                  const int32_t* __restrict__ n_dev, int pts) {
  const int64_t total = lgcn_devn(n_dev, a_cap) * pts;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = scene_of(actor_off, n_scenes, (int32_t)(i / pts));
    const float4 r = make_float4(rot[4 * b], rot[4 * b + 1], rot[4 * b + 2], rot[4 * b + 3]);   // r00 r01 r10 r11
    const float2 o = make_float2(orig[2 * b], orig[2 * b + 1]);
    const float2 x = reg[i];
    reg[i] = make_float2(__fadd_rn(__fadd_rn(__fmul_rn(x.x, r.x), __fmul_rn(x.y, r.z)), o.x),
                         __fadd_rn(__fadd_rn(__fmul_rn(x.x, r.y), __fmul_rn(x.y, r.w)), o.y));
  }
}

extern "C" int lgcn_world_transform(float* reg, const int32_t* actor_off, int n_scenes, const float* rot, const float* orig,
                                    int64_t n_actors, const int32_t* n_actors_dev, int points_per_actor, void* stream) {
  LGCN_CHECK_ARG(n_actors >= 0 && n_scenes >= 1 && points_per_actor > 0, "world_transform: sizes");
  if (n_actors == 0) return 0;
  LGCN_CHECK_ARG(reg && actor_off && rot && orig, "world_transform: NULL argument");
  k_world_transform<<<min(lgcn_cdiv(n_actors * points_per_actor, 256), 148u * 8u), 256, 0, (cudaStream_t)stream>>>(
      (float2*)reg, actor_off, n_scenes, rot, orig, n_actors, n_actors_dev, points_per_actor);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int64_t lgcn_forward_prepared_bytes(int n_scales) { return prepared_layout(n_scales).total * 4 + 1024; }

extern "C" int lgcn_forward_prepare(const LgcnForwardWeights* w, int n_scales, void* stream) {
#if LGCN_HAVE_TC
  LGCN_CHECK_ARG(w && w->prepared && w->map_input && w->map_seg && w->map_fuse && w->a2m_meta && w->m2m_fuse,
                 "forward_prepare: NULL argument");
  LGCN_CHECK_ARG(n_scales >= 1 && 2 * n_scales + 2 <= LGCN_MAX_KEYS, "forward_prepare: n_scales %d", n_scales);
  cudaStream_t st = (cudaStream_t)stream;
  const Prepared p = prepared_layout(n_scales);
  float* base = (float*)w->prepared;
  const int64_t stack = 4 * lgcn_laneconv_wpack_floats(2 * n_scales + 2);
  // the three stand-alone 128x128 matrices: MapNet input.2 / seg.2 and the first 128 input columns of A2M.meta
  const float* single[3] = {w->map_input + 3 * LGCN_C, w->map_seg + 3 * LGCN_C, w->a2m_meta};
  const int64_t ldw[3] = {LGCN_C, LGCN_C, LGCN_C + 4};
  const int64_t hi[3] = {p.in_hi, p.seg_hi, p.meta_hi}, lo[3] = {p.in_lo, p.seg_lo, p.meta_lo};
  for (int b = 0; b < 3; ++b) {
    LgcnSplitList one;
    one.n_blocks = 1;
    one.p[0] = single[b];
    one.ldw[0] = ldw[b];
    if (int rc = lgcn_split_blocks_many(one, base + hi[b], base + lo[b], st)) return rc;
  }
  if (int rc = lgcn_split_fused(w->map_fuse, base + p.map_hi, base + p.map_lo, stack, st)) return rc;
  if (int rc = lgcn_split_fused(w->m2m_fuse, base + p.m2m_hi, base + p.m2m_lo, stack, st)) return rc;
  for (int i = 0; i < 6; ++i) {
    LGCN_CHECK_ARG(w->att[i], "forward_prepare: NULL att[%d]", i);
    if (int rc = lgcn_att_split_weights(w->att[i], base + p.att_hi[i], base + p.att_lo[i], st)) return rc;
  }
  return 0;
#else
  (void)w; (void)n_scales; (void)stream;
  LGCN_CHECK_ARG(false, "forward_prepare: built without the tcgen05 engine");
  return -1;
#endif
}

extern "C" int64_t lgcn_forward_workspace_bytes(const LgcnForwardArgs* args) {
  if (check_args(args)) return -1;
  return layout_of(*args).total + 1024;
}

extern "C" void* lgcn_forward_buffer(const LgcnForwardArgs* args, int which) {
  if (check_args(args) || !args->workspace) return nullptr;
  const Layout L = layout_of(*args);
  char* ws = (char*)args->workspace;
  switch (which) {
    case 0: return ws + L.rowptr;
    case 1: return ws + L.col;
    case 2: return ws + L.e64;
    case 3: case 4: case 5: return ws + L.p_hi[which - 3];
    case 6: case 7: case 8: return ws + L.p_wi[which - 6];
    case 9: case 10: case 11: return ws + L.p_rowptr[which - 9];
    default: return nullptr;
  }
}

extern "C" int lgcn_forward(const LgcnForwardArgs* args, void* stream) {
#if LGCN_HAVE_TC
  if (check_args(args)) return -1;
  const LgcnForwardArgs& a = *args;
  LGCN_CHECK_ARG(lgcn_get_gemm_engine() == 1, "forward needs the tcgen05 engine");
  LGCN_CHECK_ARG(a.dims && a.node_off && a.actor_off && a.node_ctrs && a.node_feats && a.turn && a.control && a.intersect &&
                     a.actor_ctrs && a.segs && a.nodes && a.actors && a.status && a.workspace && a.w.prepared,
                 "forward: NULL argument");
  LGCN_CHECK_ARG(a.cap_index == 0 || a.local_idx, "forward: NULL local_idx");
  cudaStream_t st = (cudaStream_t)stream;
  const Layout L = layout_of(a);
  const Prepared P = prepared_layout(a.n_scales);
  char* ws = (char*)a.workspace;
  const float* prep = (const float*)a.w.prepared;
  const int64_t N = a.cap_nodes, A = a.cap_actors;
  const int32_t* n_nodes = a.dims;       // device-side live sizes
  const int32_t* n_actors = a.dims + 1;
  int64_t* e64 = (int64_t*)(ws + L.e64);
  float* meta = (float*)(ws + L.meta);
  int32_t* rowptr = (int32_t*)(ws + L.rowptr);
  int32_t* col = (int32_t*)(ws + L.col);
  void* plan = ws + L.plan;
  float* t0 = (float*)(ws + L.t0);
  float* hid = (float*)(ws + L.hid);
  float* xa = (float*)(ws + L.xa);
  void* att_ws = ws + L.att_ws;
  int32_t* p_tot = (int32_t*)(ws + L.p_tot);
  const int n_seg = 2 * L.n_keys * a.cap_scenes;
  LgcnFork fk;
  if (int rc = fork_of(a, &fk)) return rc;

  if (a.stages & LGCN_STAGE_GRAPH) {
    if (int rc = lgcn_zero_async(a.status, 8 * sizeof(int32_t), st)) return rc;
    // the three pair lists depend on the centres only (lanegcn.py:672-689; both Att layers of a block share theirs) and
    // not on the edge lists: with an auxiliary stream they are built beside the CSR / plan chain
    if (fk.fork(0, st) || fk.fork(1, st)) return -2;
    {
      const float* agt_c[3] = {a.node_ctrs, a.actor_ctrs, a.actor_ctrs};
      const float* ctx_c[3] = {a.actor_ctrs, a.node_ctrs, a.actor_ctrs};
      const int32_t* agt_o[3] = {a.node_off, a.actor_off, a.actor_off};
      const int32_t* ctx_o[3] = {a.actor_off, a.node_off, a.actor_off};
      const int64_t n_agt[3] = {N, A, A};
      const int32_t* n_agt_dev[3] = {n_nodes, n_actors, n_actors};
      const int64_t n_ctx[3] = {A, N, A};
      for (int i = 0; i < 3; ++i) {   // the A2M list on the first branch, M2A and A2A on the second
        cudaStream_t ps = fk.on(i == 0 ? 0 : 1, st);
        if (int rc = lgcn_launch_pairs(agt_c[i], ctx_c[i], agt_o[i], ctx_o[i], a.cap_scenes, n_agt[i], n_agt_dev[i],
                                       a.dist_th[i], a.keep_pair_quirk, (int32_t*)(ws + L.p_rowptr[i]), ws + L.p_ws[i],
                                       a.cap_pairs[i], (int32_t*)(ws + L.p_hi[i]), (int32_t*)(ws + L.p_wi[i]), p_tot + i,
                                       a.status, a.status + 1 + i, LGCN_ST_OVERFLOW_A2M << i, LGCN_ST_EMPTY_A2M << i, n_ctx[i], ps))
          return rc;
      }
    }
    // utils.to_long + the offset / cat loops of graph_gather (lanegcn.py:191-208)
    if (int rc = lgcn_offset_indices(a.local_idx, a.idx_bytes, a.segs, a.segs + n_seg + 1, n_seg, a.cap_index, e64, stream))
      return rc;
    if (int rc = lgcn_launch_pack_meta(a.turn, a.control, a.intersect, meta, N, n_nodes, st)) return rc;
    if (int rc = lgcn_launch_csr_from_segs(e64, a.segs, a.cap_scenes, L.n_keys, L.E_cap, N, n_nodes, rowptr, col,
                                           ws + L.csr_ws, a.status + 4, st))
      return rc;
    if (int rc = lgcn_launch_plan_build(rowptr, col, L.n_keys, N, n_nodes, L.E_cap, plan, st)) return rc;
    if (fk.join(0, st) || fk.join(1, st)) return -2;
  }

  if (a.stages & LGCN_STAGE_MAPNET) {
    // feat = relu(input(ctrs) + seg(feats))                                                lanegcn.py:324-327
    const float* mlp[2] = {a.w.map_input, a.w.map_seg};
    const float* src[2] = {a.node_ctrs, a.node_feats};
    for (int i = 0; i < 2; ++i) {
      const float* w = mlp[i];   // W1[128,2] | b1[128] | W2[128,128] | gamma | beta
      const bool head_in = !(lgcn_debug_get() & 131072);   // the K=2 head computed inside the Linear's kernel
      if (!head_in)
        if (int rc = lgcn_launch_mlp2_in(src[i], nullptr, nullptr, nullptr, w, w + 2 * LGCN_C, hid, N, n_nodes, st)) return rc;
      LinearArgs l = lgcn_lin1(hid, nullptr, w + 3 * LGCN_C, w + 3 * LGCN_C + CC, w + 4 * LGCN_C + CC, i ? t0 : nullptr,
                               i ? (LGCN_EPI_GN | LGCN_EPI_RES | LGCN_EPI_RELU2) : LGCN_EPI_GN, i ? a.nodes : t0, N, n_nodes);
      l.w_hi = prep + (i ? P.seg_hi : P.in_hi);
      l.w_lo = prep + (i ? P.seg_lo : P.in_lo);
      if (head_in) {
        l.head_w = w;   // W1 [128][2] | b1 [128]
        l.head_p = src[i];
      }
      if (int rc = lgcn_launch_linear(l, st)) return rc;
    }
    if (int rc = lgcn_laneconv_stack_presplit(a.nodes, t0, xa, plan, L.E_cap, L.n_keys, 4, a.w.map_fuse, prep + P.map_hi,
                                              prep + P.map_lo, N, n_nodes, st))
      return rc;
  }

  if (a.stages & LGCN_STAGE_A2M) {
    // feat = relu(GN(Linear_132->128(cat(feat, turn, control, intersect))))                 lanegcn.py:387-395
    LinearArgs m = lgcn_lin1(a.nodes, nullptr, a.w.a2m_meta, a.w.a2m_meta + (int64_t)LGCN_C * (LGCN_C + 4),
                             a.w.a2m_meta + (int64_t)LGCN_C * (LGCN_C + 4) + LGCN_C, nullptr, LGCN_EPI_GN | LGCN_EPI_RELU1,
                             t0, N, n_nodes);
    m.xs = meta;
    m.ks = 4;
    m.w_hi = prep + P.meta_hi;
    m.w_lo = prep + P.meta_lo;
    if (int rc = lgcn_launch_linear(m, st)) return rc;
    for (int i = 0; i < 2; ++i)
      if (int rc = lgcn_att_layer(i ? a.nodes : t0, a.nodes, a.actors, a.node_ctrs, a.actor_ctrs, (int32_t*)(ws + L.p_hi[0]),
                                  (int32_t*)(ws + L.p_wi[0]), (int32_t*)(ws + L.p_rowptr[0]), N, n_nodes, A,
                                  a.cap_pairs[0], p_tot + 0, a.w.att[i], prep + P.att_hi[i], prep + P.att_lo[i], att_ws, st, &fk))
        return rc;
  }

  if (a.stages & LGCN_STAGE_M2M)
    if (int rc = lgcn_laneconv_stack_presplit(a.nodes, t0, xa, plan, L.E_cap, L.n_keys, 4, a.w.m2m_fuse, prep + P.m2m_hi,
                                              prep + P.m2m_lo, N, n_nodes, st))
      return rc;

  if (a.stages & LGCN_STAGE_M2A)
    for (int i = 0; i < 2; ++i)
      if (int rc = lgcn_att_layer(a.actors, a.actors, a.nodes, a.actor_ctrs, a.node_ctrs, (int32_t*)(ws + L.p_hi[1]),
                                  (int32_t*)(ws + L.p_wi[1]), (int32_t*)(ws + L.p_rowptr[1]), A, n_actors, N,
                                  a.cap_pairs[1], p_tot + 1, a.w.att[2 + i], prep + P.att_hi[2 + i], prep + P.att_lo[2 + i],
                                  att_ws, st, &fk))
        return rc;

  if (a.stages & LGCN_STAGE_A2A)
    for (int i = 0; i < 2; ++i)
      if (int rc = lgcn_att_layer(a.actors, a.actors, a.actors, a.actor_ctrs, a.actor_ctrs, (int32_t*)(ws + L.p_hi[2]),
                                  (int32_t*)(ws + L.p_wi[2]), (int32_t*)(ws + L.p_rowptr[2]), A, n_actors, A,
                                  a.cap_pairs[2], p_tot + 2, a.w.att[4 + i], prep + P.att_hi[4 + i], prep + P.att_lo[4 + i],
                                  att_ws, st, &fk))
        return rc;
  return 0;
#else
  (void)args; (void)stream;
  LGCN_CHECK_ARG(false, "forward: built without the tcgen05 engine");
  return -1;
#endif
}
