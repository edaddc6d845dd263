// api.cu — C-ABI glue: error reporting, engine selection, lgcn_linear128 argument checking, and the two
// fused per-module sequences (LaneConv stack, Att layer) that keep the host at one call per module.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void lgcn_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t lgcn_launch_count(void) { return g_launches.load(); }

// ---- per-kernel timing: event pairs on the launching stream, summed per kind by lgcn_prof_collect / lgcn_prof_peek.
// Mode 1 brackets EAGER launches (captures are left alone); mode 2 brackets launches INSIDE a stream capture with
// external event-record nodes, so every replay of that graph re-records them: the kernels are timed in the very
// sequence (branch overlap, stream priorities) the product runs.
struct ProfEv { cudaEvent_t a, b; int kind; };
static std::vector<ProfEv> g_prof_pool;
static size_t g_prof_used = 0;
static int g_prof_on = 0;
LgcnProfScope::LgcnProfScope(int kind, cudaStream_t s) : slot(-1), st(s), external(false) {
  if (!g_prof_on) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) return;
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  if (capturing != (g_prof_on == 2)) return;
  external = capturing;
  if (g_prof_used == g_prof_pool.size()) {
    ProfEv e;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    g_prof_pool.push_back(e);
  }
  slot = (int)g_prof_used++;
  g_prof_pool[slot].kind = kind;
  cudaEventRecordWithFlags(g_prof_pool[slot].a, st, external ? cudaEventRecordExternal : cudaEventRecordDefault);
}
LgcnProfScope::~LgcnProfScope() {
  if (slot >= 0) cudaEventRecordWithFlags(g_prof_pool[slot].b, st, external ? cudaEventRecordExternal : cudaEventRecordDefault);
}
extern "C" int lgcn_prof_enable(int on) {
  const int prev = g_prof_on;
  g_prof_on = on < 0 || on > 2 ? 0 : on;
  return prev;
}
static int prof_sum(double* ms_by_kind, int64_t* launches_by_kind) {
  for (int k = 0; k < LGCN_PROF_KINDS; ++k) {
    ms_by_kind[k] = 0.0;
    launches_by_kind[k] = 0;
  }
  for (size_t i = 0; i < g_prof_used; ++i) {
    float ms = 0.f;
    LGCN_CUDA_OK(cudaEventSynchronize(g_prof_pool[i].b));
    LGCN_CUDA_OK(cudaEventElapsedTime(&ms, g_prof_pool[i].a, g_prof_pool[i].b));
    ms_by_kind[g_prof_pool[i].kind] += ms;
    launches_by_kind[g_prof_pool[i].kind] += 1;
  }
  return 0;
}
extern "C" int lgcn_prof_peek(double* ms_by_kind, int64_t* launches_by_kind) { return prof_sum(ms_by_kind, launches_by_kind); }
extern "C" int lgcn_prof_collect(double* ms_by_kind, int64_t* launches_by_kind) {
  const int rc = prof_sum(ms_by_kind, launches_by_kind);
  g_prof_used = 0;
  return rc;
}
static int g_engine = -1;  // -1: not decided yet
static int g_debug = 0;
int lgcn_debug_get() { return g_debug; }
extern "C" int lgcn_debug_flags(int flags) {
  const int prev = g_debug;
  g_debug = flags;
  return prev;
}

void lgcn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* lgcn_last_error(void) { return g_err; }
extern "C" int lgcn_version(void) { return 100; }

extern "C" int lgcn_get_gemm_engine(void) {
  if (g_engine < 0) {
    const char* e = getenv("LGCN_GEMM_ENGINE");
    g_engine = (e && strcmp(e, "simt") == 0) ? 0 : LGCN_DEFAULT_ENGINE;
  }
  return g_engine;
}
extern "C" int lgcn_set_gemm_engine(int engine) {
  const int prev = lgcn_get_gemm_engine();
  g_engine = (engine && LGCN_HAVE_TC) ? 1 : 0;  // engine 1 only exists when gemm_tc.cu was built in
  return prev;
}

int lgcn_launch_linear(const LinearArgs& a, cudaStream_t st) {
#if LGCN_HAVE_TC
  if (lgcn_get_gemm_engine() == 1) return lgcn_launch_linear_tc(a, st);
#endif
  return lgcn_launch_linear_simt(a, st);
}

static int check_linear(const LinearArgs& a) {
  LGCN_CHECK_ARG(a.n_src >= 1 && a.n_src <= 3, "linear128: n_src %d not in 1..3", a.n_src);
  LGCN_CHECK_ARG(a.ks == 0 || a.ks == 4, "linear128: ks %d (only 0 or 4: W rows must stay 16-byte aligned)", a.ks);
  LGCN_CHECK_ARG(a.ks == 0 || a.xs, "linear128: ks > 0 without xs");
  LGCN_CHECK_ARG(a.n_out_blocks >= 1, "linear128: n_out_blocks %d", a.n_out_blocks);
  LGCN_CHECK_ARG(a.flags == 0 || a.n_out_blocks == 1, "linear128: epilogue flags need n_out_blocks == 1");
  LGCN_CHECK_ARG(!(a.flags & LGCN_EPI_GN) || (a.gamma && a.beta), "linear128: GN without gamma/beta");
  LGCN_CHECK_ARG(!(a.flags & LGCN_EPI_RES) || a.res, "linear128: RES without res");
  LGCN_CHECK_ARG(a.ldo >= (int64_t)a.n_out_blocks * LGCN_C && a.ldo % 4 == 0, "linear128: bad ldo %lld", (long long)a.ldo);
  for (int s = 0; s < a.n_src; ++s) LGCN_CHECK_ARG(a.a[s], "linear128: source %d is NULL", s);
  LGCN_CHECK_ARG(a.W && a.out, "linear128: NULL W/out");
  return 0;
}

static int linear128_impl(const float* a0, const int32_t* idx0, const float* a1, const int32_t* idx1, const float* a2,
                          const int32_t* idx2, int n_src, const float* xs, int ks, const float* W, int n_out_blocks,
                          const float* gamma, const float* beta, const float* res, int flags, float* out, int64_t ldo,
                          int64_t m, void* workspace, void* stream) {
  LinearArgs a;
  memset(&a, 0, sizeof(a));
  a.a[0] = a0; a.a[1] = a1; a.a[2] = a2;
  a.idx[0] = idx0; a.idx[1] = idx1; a.idx[2] = idx2;
  a.n_src = n_src; a.xs = xs; a.ks = ks; a.W = W; a.n_out_blocks = n_out_blocks;
  a.gamma = gamma; a.beta = beta; a.res = res; a.flags = flags; a.out = out; a.ldo = ldo; a.m = m;
  a.dbg = g_debug;
  a.split_ws = workspace;
  if (check_linear(a)) return -1;
  return lgcn_launch_linear(a, (cudaStream_t)stream);
}

extern "C" int64_t lgcn_linear128_workspace_bytes(void) {
#if LGCN_HAVE_TC
  return lgcn_linear_split_bytes();
#else
  return 0;
#endif
}

extern "C" int lgcn_linear128_ws(const float* a0, const int32_t* idx0, const float* a1, const int32_t* idx1,
                                 const float* a2, const int32_t* idx2, int n_src, const float* xs, int ks,
                                 const float* W, int n_out_blocks, const float* gamma, const float* beta,
                                 const float* res, int flags, float* out, int64_t ldo, int64_t m, void* workspace,
                                 void* stream) {
  LGCN_CHECK_ARG(workspace || lgcn_get_gemm_engine() == 0, "linear128_ws: NULL workspace");
  return linear128_impl(a0, idx0, a1, idx1, a2, idx2, n_src, xs, ks, W, n_out_blocks, gamma, beta, res, flags, out, ldo, m,
                        workspace, stream);
}

extern "C" int lgcn_linear128(const float* a0, const int32_t* idx0, const float* a1, const int32_t* idx1,
                              const float* a2, const int32_t* idx2, int n_src, const float* xs, int ks,
                              const float* W, int n_out_blocks, const float* gamma, const float* beta,
                              const float* res, int flags, float* out, int64_t ldo, int64_t m, void* stream) {
  return linear128_impl(a0, idx0, a1, idx1, a2, idx2, n_src, xs, ks, W, n_out_blocks, gamma, beta, res, flags, out, ldo, m,
                        nullptr, stream);
}

LinearArgs lgcn_lin1(const float* x, const int32_t* idx, const float* W, const float* gamma, const float* beta,
                     const float* res, int flags, float* out, int64_t m, const int32_t* m_dev) {
  LinearArgs a;
  memset(&a, 0, sizeof(a));
  a.a[0] = x; a.idx[0] = idx; a.n_src = 1; a.W = W; a.n_out_blocks = 1;
  a.gamma = gamma; a.beta = beta; a.res = res; a.flags = flags; a.out = out; a.ldo = LGCN_C; a.m = m; a.m_dev = m_dev;
  a.dbg = g_debug;
  return a;
}
static LinearArgs lin1(const float* x, const int32_t* idx, const float* W, const float* gamma, const float* beta,
                       const float* res, int flags, float* out, int64_t m) {
  return lgcn_lin1(x, idx, W, gamma, beta, res, flags, out, m, nullptr);
}

// ------------------------------------------------------------------ LaneConv stack
#define CC ((int64_t)LGCN_C * LGCN_C)

extern "C" int64_t lgcn_laneconv_wpack_floats(int n_keys) { return (int64_t)(n_keys + 1) * CC + CC + 4 * LGCN_C; }

extern "C" int64_t lgcn_laneconv_workspace_bytes(int64_t n_nodes, int n_keys) {
  // Y [n, (K+1)*128] | h [n,128] | W_hi, W_lo [(K+1)*128, 128] (tf32 split of the current block's projection)
  return lgcn_align_up(n_nodes * (int64_t)(n_keys + 1) * LGCN_C * 4, 1024) + lgcn_align_up(n_nodes * LGCN_C * 4, 1024) +
         2 * lgcn_align_up((int64_t)(n_keys + 1) * CC * 4, 1024) + 1024;
}

extern "C" int lgcn_laneconv_stack(float* feat, const int32_t* rowptr, const int32_t* col, int n_keys,
                                   int n_blocks, const float* wpack, int64_t n_nodes, void* workspace,
                                   void* stream) {
  LGCN_CHECK_ARG(n_keys >= 0 && n_keys <= LGCN_MAX_KEYS, "laneconv_stack: n_keys %d", n_keys);
  LGCN_CHECK_ARG(feat && rowptr && wpack && workspace, "laneconv_stack: NULL argument");
  if (n_nodes <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = n_keys + 1;
  float* Y = (float*)workspace;
  float* h = (float*)((char*)workspace + lgcn_align_up(n_nodes * (int64_t)nb * LGCN_C * 4, 1024));
  float* w_hi = (float*)((char*)h + lgcn_align_up(n_nodes * LGCN_C * 4, 1024));
  float* w_lo = (float*)((char*)w_hi + lgcn_align_up((int64_t)nb * CC * 4, 1024));
#if LGCN_HAVE_TC
  const bool wide_tc = lgcn_get_gemm_engine() == 1 && nb >= 2 && !(g_debug & 16);
#else
  const bool wide_tc = false;
#endif
  for (int i = 0; i < n_blocks; ++i) {
    const float* w = wpack + (int64_t)i * lgcn_laneconv_wpack_floats(n_keys);
    const float* wctr2 = w + (int64_t)nb * CC;
    const float* gn_g = wctr2 + CC;
    const float* gn_b = gn_g + LGCN_C;
    const float* c2_g = gn_b + LGCN_C;
    const float* c2_b = c2_g + LGCN_C;
    // (1) all nb projections of every node in one wide GEMM: Y[n, k*128:(k+1)*128] = feat[n] . W_k^T
    LinearArgs wide = lin1(feat, nullptr, w, nullptr, nullptr, nullptr, 0, Y, n_nodes);
    wide.n_out_blocks = nb;
    wide.ldo = (int64_t)nb * LGCN_C;
    if (wide_tc) {  // static weights: split into tf32 hi/lo once per block, then TMA feeds them to the tensor pipe
#if LGCN_HAVE_TC
      if (int rc = lgcn_split_tf32(w, w_hi, w_lo, (int64_t)nb * CC, st)) return rc;
      LgcnProfScope ps(LGCN_PROF_WIDE, st);
      if (int rc = lgcn_launch_wide_tc(wide, w_hi, w_lo, st)) return rc;
#endif
    } else {
      LgcnProfScope ps(LGCN_PROF_WIDE, st);
      if (int rc = lgcn_launch_linear(wide, st)) return rc;
    }
    // (2) temp = ctr + sum over destination-sorted edges; h = relu(GN(temp))
    {
      LgcnProfScope ps(LGCN_PROF_GATHER, st);
      if (int rc = lgcn_laneconv_gather_gn_relu(Y, nb, rowptr, col, gn_g, gn_b, h, n_nodes, stream)) return rc;
    }
    // (3) feat = relu(GN(h . Wctr2^T) + feat)     (res == current feat; written in place)
    LinearArgs c2 = lin1(h, nullptr, wctr2, c2_g, c2_b, feat, LGCN_EPI_GN | LGCN_EPI_RES | LGCN_EPI_RELU2, feat, n_nodes);
    {
      LgcnProfScope ps(LGCN_PROF_CTR2, st);
      if (int rc = lgcn_launch_linear(c2, st)) return rc;
    }
  }
  return 0;
}

// Planned (aggregate-first) stack: every block is ONE kernel (laneconv_fused.cu); the features ping-pong between
// `feat` and a workspace buffer because a block reads neighbour rows while other tiles already write theirs.
// workspace: other feature buffer | aux rows | tf32 hi / lo copies of up to LGCN_MAX_PLANNED_BLOCKS blocks' wpack
#define LGCN_MAX_PLANNED_BLOCKS 8
extern "C" int64_t lgcn_laneconv_planned_workspace_bytes(int64_t n_nodes, int64_t n_edges, int n_keys) {
#if LGCN_HAVE_TC
  return lgcn_align_up(n_nodes * LGCN_C * 4, 1024) + lgcn_laneconv_fused_aux_bytes(n_edges) +
         2 * lgcn_align_up(LGCN_MAX_PLANNED_BLOCKS * lgcn_laneconv_wpack_floats(n_keys) * 4, 1024) + 1024;
#else
  (void)n_nodes; (void)n_edges; (void)n_keys;
  return 0;
#endif
}

// blocks of the stack on weights that are ALREADY split (w_hi / w_lo: tf32 hi / lo images of the whole wpack, norm
// vectors included at their places); `other` [n,128] and `xa` (aux rows) are scratch.  n_dev: live row count in
// device memory (n_nodes is then the capacity).
int lgcn_laneconv_stack_presplit(float* feat, float* other, float* xa, void* plan, int64_t n_edges, int n_keys,
                                 int n_blocks, const float* wpack, const float* w_hi, const float* w_lo, int64_t n_nodes,
                                 const int32_t* n_dev, cudaStream_t st) {
#if LGCN_HAVE_TC
  const int nb = n_keys + 1;
  const int64_t per = lgcn_laneconv_wpack_floats(n_keys);
  for (int i = 0; i < n_blocks; ++i) {
    const float* w = wpack + (int64_t)i * per;   // Wcat | Wctr2 | 4 norm vectors
    const float* gn = w + (int64_t)(nb + 1) * CC;
    const float* src = (i & 1) ? other : feat;
    float* dst = (i & 1) ? feat : other;
    LgcnProfScope ps(LGCN_PROF_FUSED, st);
    if (int rc = lgcn_launch_laneconv_fused(src, dst, plan, n_nodes, n_dev, n_edges, n_keys, w_hi + (int64_t)i * per,
                                            w_lo + (int64_t)i * per, gn, xa, 1, st))
      return rc;
  }
  if (n_blocks & 1) LGCN_CUDA_OK(cudaMemcpyAsync(feat, other, n_nodes * LGCN_C * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
#else
  (void)feat; (void)other; (void)xa; (void)plan; (void)n_edges; (void)n_keys; (void)n_blocks; (void)wpack; (void)w_hi;
  (void)w_lo; (void)n_nodes; (void)n_dev; (void)st;
  LGCN_CHECK_ARG(false, "laneconv_stack_planned: built without the tcgen05 engine");
  return -1;
#endif
}

extern "C" int lgcn_laneconv_stack_planned(float* feat, void* plan, int64_t n_edges, int n_keys, int n_blocks,
                                           const float* wpack, int64_t n_nodes, void* workspace, void* stream) {
#if LGCN_HAVE_TC
  LGCN_CHECK_ARG(n_keys >= 0 && n_keys <= LGCN_MAX_KEYS, "laneconv_stack_planned: n_keys %d", n_keys);
  LGCN_CHECK_ARG(n_blocks >= 0 && n_blocks <= LGCN_MAX_PLANNED_BLOCKS, "laneconv_stack_planned: n_blocks %d", n_blocks);
  LGCN_CHECK_ARG(feat && plan && wpack && workspace, "laneconv_stack_planned: NULL argument");
  LGCN_CHECK_ARG(lgcn_get_gemm_engine() == 1, "laneconv_stack_planned needs the tcgen05 engine");
  if (n_nodes <= 0 || n_blocks == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per = lgcn_laneconv_wpack_floats(n_keys);
  float* other = (float*)workspace;
  float* xa = (float*)((char*)other + lgcn_align_up(n_nodes * LGCN_C * 4, 1024));
  float* w_hi = (float*)((char*)xa + lgcn_laneconv_fused_aux_bytes(n_edges));
  float* w_lo = (float*)((char*)w_hi + lgcn_align_up(LGCN_MAX_PLANNED_BLOCKS * per * 4, 1024));
  // the whole pack (weights and, harmlessly, the norm vectors between them) is split by one launch
  if (int rc = lgcn_split_fused(wpack, w_hi, w_lo, (int64_t)n_blocks * per, st)) return rc;
  return lgcn_laneconv_stack_presplit(feat, other, xa, plan, n_edges, n_keys, n_blocks, wpack, w_hi, w_lo, n_nodes,
                                      nullptr, st);
#else
  (void)feat; (void)plan; (void)n_edges; (void)n_keys; (void)n_blocks; (void)wpack; (void)n_nodes; (void)workspace; (void)stream;
  LGCN_CHECK_ARG(false, "laneconv_stack_planned: built without the tcgen05 engine");
  return -1;
#endif
}

// ------------------------------------------------------------------ Att layer
struct AttW {
  const float *d0w, *d0b, *d2w, *d2g, *d2b, *qw, *qg, *qb, *c0w, *c0g, *c0b, *c1w, *aw, *ng, *nb, *lw, *lg, *lb;
};
static AttW att_unpack(const float* p) {
  AttW w;
  w.d0w = p; p += 2 * LGCN_C;
  w.d0b = p; p += LGCN_C;
  w.d2w = p; p += CC;
  w.d2g = p; p += LGCN_C;
  w.d2b = p; p += LGCN_C;
  w.qw = p; p += CC;
  w.qg = p; p += LGCN_C;
  w.qb = p; p += LGCN_C;
  w.c0w = p; p += 3 * CC;
  w.c0g = p; p += LGCN_C;
  w.c0b = p; p += LGCN_C;
  w.c1w = p; p += CC;
  w.aw = p; p += CC;
  w.ng = p; p += LGCN_C;
  w.nb = p; p += LGCN_C;
  w.lw = p; p += CC;
  w.lg = p; p += LGCN_C;
  w.lb = p; p += LGCN_C;
  return w;
}
extern "C" int64_t lgcn_att_wpack_floats(void) { return 2 * LGCN_C + LGCN_C + 8 * CC + 10 * LGCN_C; }

extern "C" int64_t lgcn_att_workspace_bytes(int64_t n_agt, int64_t n_pairs) {
  // 3 pair-row buffers | 2 agent-row buffers | the layer's 8 weight blocks split into tf32 hi / lo
  return 3 * lgcn_align_up(n_pairs * LGCN_C * 4, 1024) + 2 * lgcn_align_up(n_agt * LGCN_C * 4, 1024) + 2 * 8 * CC * 4 + 1024;
}

// The six weight matrices of an Att layer as 8 tf32 hi / lo blocks (ONE launch):
// 0 dist.2 | 1 query | 2-4 ctx.0 (its three K=128 slices) | 5 ctx.1 | 6 agt | 7 linear
int lgcn_att_split_weights(const float* wpack, float* WH, float* WL, cudaStream_t st) {
  const AttW w = att_unpack(wpack);
  LgcnSplitList sl;
  const float* ps[8] = {w.d2w, w.qw, w.c0w, w.c0w + LGCN_C, w.c0w + 2 * LGCN_C, w.c1w, w.aw, w.lw};
  for (int b = 0; b < 8; ++b) {
    sl.p[b] = ps[b];
    sl.ldw[b] = (b >= 2 && b <= 4) ? 3 * LGCN_C : LGCN_C;
  }
  sl.n_blocks = 8;
  return lgcn_split_blocks_many(sl, WH, WL, st);
}

// One Att layer with every size either by value or in device memory (n_agt_dev / n_pairs_dev != NULL: n_agt / n_pairs
// are capacities).  WH / WL: the layer's pre-split weights (lgcn_att_split_weights) or NULL on the fp32 SIMT engine.
// scratch: P0 P1 P2 [n_pairs,128] | A0 A1 [n_agt,128] (lgcn_att_workspace_bytes layout).  n_ctx == 0 selects the
// reference's early-out (lanegcn.py:664-670) and is a host-side decision.
int lgcn_att_layer(const float* agts_in, float* agts_out, const float* ctx, const float* agt_ctrs, const float* ctx_ctrs,
                   const int32_t* hi, const int32_t* wi, const int32_t* rowptr, int64_t n_agt, const int32_t* n_agt_dev,
                   int64_t n_ctx, int64_t n_pairs, const int32_t* n_pairs_dev, const float* wpack, const float* WH,
                   const float* WL, void* workspace, cudaStream_t st, const LgcnFork* fk) {
  if (n_agt <= 0) return 0;
  LgcnProfScope ps(LGCN_PROF_ATT, st);
  static const LgcnFork serial = {{nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  if (!fk) fk = &serial;
  const AttW w = att_unpack(wpack);
  const int64_t pb = lgcn_align_up(n_pairs * LGCN_C * 4, 1024), ab = lgcn_align_up(n_agt * LGCN_C * 4, 1024);
  float* P0 = (float*)workspace;
  float* P1 = (float*)((char*)workspace + pb);
  float* P2 = (float*)((char*)workspace + 2 * pb);
  float* A0 = (float*)((char*)workspace + 3 * pb);
  float* A1 = (float*)((char*)workspace + 3 * pb + ab);
  const int kLin = LGCN_EPI_GN | LGCN_EPI_RES | LGCN_EPI_RELU2;
  const bool pre = WH != nullptr && WL != nullptr;
  auto with_w = [&](LinearArgs a, int blk) {
    if (pre) {
      a.w_hi = WH + (int64_t)blk * CC;
      a.w_lo = WL + (int64_t)blk * CC;
    }
    return a;
  };
  auto agt_rows = [&](const float* x, const int32_t* idx, const float* W, const float* g, const float* b, const float* res,
                      int flags, float* out) { return lgcn_lin1(x, idx, W, g, b, res, flags, out, n_agt, n_agt_dev); };
  auto pair_rows = [&](const float* x, const int32_t* idx, const float* W, const float* g, const float* b, const float* res,
                       int flags, float* out) { return lgcn_lin1(x, idx, W, g, b, res, flags, out, n_pairs, n_pairs_dev); };
  if (n_ctx == 0) {  // lanegcn.py:664-670 — no self.norm on this path
    LinearArgs a = with_w(agt_rows(agts_in, nullptr, w.aw, nullptr, nullptr, nullptr, LGCN_EPI_RELU1, A0), 6);
    if (int rc = lgcn_launch_linear(a, st)) return rc;
    LinearArgs l = with_w(agt_rows(A0, nullptr, w.lw, w.lg, w.lb, agts_in, kLin, agts_out), 7);
    return lgcn_launch_linear(l, st);
  }
  LGCN_CHECK_ARG(n_pairs > 0, "att_forward: no agent/context pair within the distance threshold in any scene "
                              "(the reference raises at lanegcn.py:688: torch.cat of an empty list)");
  LGCN_CHECK_ARG(ctx && agt_ctrs && ctx_ctrs && hi && wi && rowptr, "att_forward: NULL argument");
  // Three independent chains start a layer: dist (pair rows), query, and the agent-side Linear `agt`; with auxiliary
  // streams the last two run beside the first (the layer's critical path is then 6 kernels instead of 8).
  if (fk->fork(0, st) || fk->fork(1, st)) return -2;
  // query = relu(GN(L(agts[hi])))  — a per-row function: computed per agent when that is fewer rows   :696
  // (decided on the capacities when the sizes live on the device)
  const float* q;
  const int32_t* qidx;
  if (n_agt <= n_pairs) {
    LinearArgs qa = with_w(agt_rows(agts_in, nullptr, w.qw, w.qg, w.qb, nullptr, LGCN_EPI_GN | LGCN_EPI_RELU1, A0), 1);
    if (int rc = lgcn_launch_linear(qa, fk->on(0, st))) return rc;
    q = A0; qidx = hi;
  } else {
    LinearArgs qp = with_w(pair_rows(agts_in, hi, w.qw, w.qg, w.qb, nullptr, LGCN_EPI_GN | LGCN_EPI_RELU1, P2), 1);
    if (int rc = lgcn_launch_linear(qp, fk->on(0, st))) return rc;
    q = P2; qidx = nullptr;
  }
  // agt(agts), the Linear of lanegcn.py:702
  LinearArgs ag = with_w(agt_rows(agts_in, nullptr, w.aw, nullptr, nullptr, nullptr, 0, A1), 6);
  if (int rc = lgcn_launch_linear(ag, fk->on(1, st))) return rc;
  // dist = relu(GN(L(relu(L2(agt_ctrs[hi] - ctx_ctrs[wi])))))                      lanegcn.py:693-694
  // (with pre-split weights the K=2 head is computed inside dist.2's kernel: its rows never exist in memory)
  const bool head_in = pre && !(lgcn_debug_get() & 131072);
  if (!head_in)
    if (int rc = lgcn_launch_mlp2_in(agt_ctrs, hi, ctx_ctrs, wi, w.d0w, w.d0b, P0, n_pairs, n_pairs_dev, st)) return rc;
  LinearArgs d2 = with_w(pair_rows(P0, nullptr, w.d2w, w.d2g, w.d2b, nullptr, LGCN_EPI_GN | LGCN_EPI_RELU1, P1), 0);
  if (head_in) {
    d2.head_w = w.d0w;   // d0w [128][2] | d0b [128] are adjacent in the pack
    d2.head_p = agt_ctrs; d2.head_ip = hi;
    d2.head_q = ctx_ctrs; d2.head_iq = wi;
  }
  if (int rc = lgcn_launch_linear(d2, st)) return rc;
  if (fk->join(0, st)) return -2;
  // ctx = L(relu(GN(L384(cat(dist, query, ctx[wi])))))  — split-K over the three sources, no cat   :698-700
  // with pre-split weights ctx.1 is chained inside ctx.0's kernel (its block follows ctx.0's three in WH / WL)
  const bool chain_c1 = pre && !(lgcn_debug_get() & 65536);
  LinearArgs c0 = with_w(pair_rows(P1, nullptr, w.c0w, w.c0g, w.c0b, nullptr, LGCN_EPI_GN | LGCN_EPI_RELU1, P0), 2);
  c0.n_src = 3;
  c0.a[1] = q; c0.idx[1] = qidx;
  c0.a[2] = ctx; c0.idx[2] = wi;
  c0.chain = chain_c1 ? 1 : 0;
  c0.flags2 = 0;
  if (int rc = lgcn_launch_linear(c0, st)) return rc;
  const float* ctx_out = P0;   // P0 is free again: dist.2 has consumed the K=2 head's rows
  if (!chain_c1) {
    LinearArgs c1 = with_w(pair_rows(P0, nullptr, w.c1w, nullptr, nullptr, nullptr, 0, P1), 5);
    if (int rc = lgcn_launch_linear(c1, st)) return rc;
    ctx_out = P1;
  }
  if (fk->join(1, st)) return -2;
  // agts = relu(GN(agt(agts) + scatter(ctx by hi)))                                                  :702-705
  if (int rc = lgcn_launch_segsum_gn_relu(A1, ctx_out, rowptr, w.ng, w.nb, A0, n_agt, n_agt_dev, st)) return rc;
  // agts = relu(GN(linear(agts)) + res)                                                              :707-709
  LinearArgs l = with_w(agt_rows(A0, nullptr, w.lw, w.lg, w.lb, agts_in, kLin, agts_out), 7);
  return lgcn_launch_linear(l, st);
}

extern "C" int lgcn_att_forward(const float* agts_in, float* agts_out, const float* ctx, const float* agt_ctrs,
                                const float* ctx_ctrs, const int32_t* hi, const int32_t* wi,
                                const int32_t* rowptr, int64_t n_agt, int64_t n_ctx, int64_t n_pairs,
                                const float* wpack, void* workspace, void* stream) {
  LGCN_CHECK_ARG(agts_in && agts_out && wpack && workspace, "att_forward: NULL argument");
  if (n_agt <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t pb = lgcn_align_up(n_pairs * LGCN_C * 4, 1024), ab = lgcn_align_up(n_agt * LGCN_C * 4, 1024);
  float* WH = (float*)((char*)workspace + 3 * pb + 2 * ab);
  float* WL = WH + 8 * CC;
  const bool pre = lgcn_get_gemm_engine() == 1;
  if (pre)
    if (int rc = lgcn_att_split_weights(wpack, WH, WL, st)) return rc;
  return lgcn_att_layer(agts_in, agts_out, ctx, agt_ctrs, ctx_ctrs, hi, wi, rowptr, n_agt, nullptr, n_ctx, n_pairs, nullptr,
                        wpack, pre ? WH : nullptr, pre ? WL : nullptr, workspace, st, nullptr);
}
