// common.cuh — shared helpers of liblgcn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lgcn.h"
#include "../../include/lgcn_debug.h"

#define LGCN_WARP 32
#define LGCN_GN_EPS 1e-5f

void lgcn_set_error(const char* fmt, ...);

#define LGCN_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      lgcn_set_error(__VA_ARGS__);     \
      return -1;                       \
    }                                  \
  } while (0)

#define LGCN_CUDA_OK(expr)                                                                         \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      lgcn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));        \
      return -2;                                                                                   \
    }                                                                                              \
  } while (0)

void lgcn_count_launch();
#define LGCN_LAUNCH_OK()                   \
  do {                                     \
    lgcn_count_launch();                   \
    LGCN_CUDA_OK(cudaPeekAtLastError());   \
  } while (0)

// optional per-kernel timing (lgcn_prof_*): CUDA events recorded on the launching stream around a launch
enum { LGCN_PROF_WIDE = 0, LGCN_PROF_GATHER = 1, LGCN_PROF_CTR2 = 2, LGCN_PROF_ATT = 3, LGCN_PROF_FUSED = 4, LGCN_PROF_BLOCK_KERNEL = 5, LGCN_PROF_KINDS = 8 };
struct LgcnProfScope {
  int slot;
  cudaStream_t st;
  bool external;
  LgcnProfScope(int kind, cudaStream_t st);
  ~LgcnProfScope();
};

// CTA size of the index kernels that run next to ActorNet at the head of the forward (graph build, pair lists).  Two
// ActorNet CTAs hold 224 of an SM's 228 KB of shared memory, and every CTA reserves 1 KB of it: only TWO more CTAs fit
// beside them whatever their own footprint, so the threads have to come from the CTA size (128-thread CTAs ran 256
// threads per SM: k_csr_finish 186 us).
constexpr int kIndexThreads = 512;

// Programmatic dependent launch (PDL).  A kernel launched through lgcn_launch_pdl may START while its predecessor on the
// stream is still draining: its CTAs take SMs as they free up and run their prologue (barrier init, tensor-memory
// allocation, norm vectors) up to lgcn_pdl_wait(), which returns once the predecessor has completed and flushed.
// Rules: (1) such a kernel executes lgcn_pdl_wait() in EVERY thread before it reads or writes anything another kernel
// of the stream touches (weights are static: fine before); (2) kernels that may precede one call lgcn_pdl_trigger()
// first thing (without it the dependent simply starts when they end).  Both are no-ops in a normal launch.
// Debug flag 16384 (include/lgcn_debug.h) turns the attribute off.
__device__ __forceinline__ void lgcn_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void lgcn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
int lgcn_debug_get();
template <typename... KArgs, typename... Args>
static inline cudaError_t lgcn_launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (lgcn_debug_get() & 16384) ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Optional parallel branches inside one library call: independent kernels are forked from the call's stream onto up to two
// caller-provided auxiliary streams and joined before anything reads their results / before the call returns (plain
// event record + wait: capturable, the branches become parallel paths of the CUDA graph).  aux[i] == NULL: stay serial.
struct LgcnFork {
  cudaStream_t aux[2];
  cudaEvent_t ev[3];   // [0] fork point on the main stream, [1 + i] end of branch i
  bool has(int i) const { return aux[i] != nullptr; }
  int fork(int i, cudaStream_t st) const {
    if (!has(i)) return 0;
    if (cudaEventRecord(ev[0], st) != cudaSuccess || cudaStreamWaitEvent(aux[i], ev[0], 0) != cudaSuccess) return -2;
    return 0;
  }
  int join(int i, cudaStream_t st) const {
    if (!has(i)) return 0;
    if (cudaEventRecord(ev[1 + i], aux[i]) != cudaSuccess || cudaStreamWaitEvent(st, ev[1 + i], 0) != cudaSuccess) return -2;
    return 0;
  }
  cudaStream_t on(int i, cudaStream_t st) const { return has(i) ? aux[i] : st; }
};

static inline int64_t lgcn_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline unsigned lgcn_cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// Row counts of the forward path may live in DEVICE memory (so that a whole forward is one CUDA graph per capacity
// bucket and no host round trip carries a data-dependent size): n = n_dev ? min(*n_dev, cap) : cap.
__device__ __forceinline__ int64_t lgcn_devn(const int32_t* n_dev, int64_t cap) {
  if (!n_dev) return cap;
  const int64_t n = *n_dev;
  return n < cap ? (n < 0 ? 0 : n) : cap;
}

// scene owning global row r: largest b with off[b] <= r (off is ascending, may contain empty scenes)
__device__ __forceinline__ int scene_of(const int32_t* __restrict__ off, int n_scenes, int32_t r) {
  int lo = 0, hi = n_scenes;  // invariant: off[lo] <= r < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= r) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 128-bit streaming load: data that is read once (rows of the wide projection) should not displace L1.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// GroupNorm with ONE group over the 128 channels held by a warp (4 per lane) — layers.py:72, eps 1e-5,
// biased variance; two-pass (mean, then centred squares) for accuracy.
__device__ __forceinline__ float4 warp_gn128(float4 x, float4 g, float4 b) {
  const float mean = warp_sum((x.x + x.y) + (x.z + x.w)) * (1.0f / 128.0f);
  float4 d = make_float4(x.x - mean, x.y - mean, x.z - mean, x.w - mean);
  const float var = warp_sum((d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w)) * (1.0f / 128.0f);
  const float rstd = 1.0f / sqrtf(var + LGCN_GN_EPS);
  return make_float4(d.x * rstd * g.x + b.x, d.y * rstd * g.y + b.y, d.z * rstd * g.z + b.z,
                     d.w * rstd * g.w + b.w);
}

__device__ __forceinline__ float4 relu4(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}

// ---- internal launchers shared between translation units (all enqueue on `st`, return 0 / <0)
struct LinearArgs {
  const float* a[3];
  const int32_t* idx[3];
  int n_src;
  const float* xs;
  int ks;
  const float* W;
  int n_out_blocks;
  const float* gamma;
  const float* beta;
  const float* res;
  int flags;
  float* out;
  int64_t ldo;
  int64_t m;
  const int32_t* m_dev;  // optional: live row count in device memory (m is then the capacity); one output block only
  int dbg;  // ablation switches for profiling (lgcn_debug_flags); 0 in production
  // optional: the weights already split into tf32 hi / lo blocks [n_src*128, 128] (lgcn_split_blocks_many); when
  // NULL the tcgen05 path splits W into a scratch slot at launch
  const float* w_hi;
  const float* w_lo;
  // optional caller-provided scratch for that split (lgcn_linear128_workspace_bytes); without it the tcgen05 path uses
  // the library's lazily allocated per-device scratch ring (lgcn_linear128 only)
  void* split_ws;
  // optional (tcgen05, one output block): a SECOND Linear chained inside the kernel, y = epi2(W2 . epi1(...)): its tf32
  // images are the block that follows the sources' blocks in w_hi / w_lo; flags2 = its LGCN_EPI_* (no GroupNorm)
  int chain;
  int flags2;
  // optional (tcgen05, pre-split weights): source 0 is COMPUTED instead of read, relu(W1 . (p[ip ? ip[m] : m] - q[iq ? iq[m] : m]) + b1)
  // with head_w = W1 [128][2] | b1 [128] (the nn.Linear(2, 128) + ReLU heads; q may be NULL); a[0] / idx[0] are ignored
  const float* head_w;
  const float* head_p;
  const int32_t* head_ip;
  const float* head_q;
  const int32_t* head_iq;
};
int lgcn_debug_get();
int lgcn_launch_linear_simt(const LinearArgs& a, cudaStream_t st);
int lgcn_launch_linear_tc(const LinearArgs& a, cudaStream_t st);  // tcgen05 3xTF32 (gemm_tc.cu)
int lgcn_launch_linear(const LinearArgs& a, cudaStream_t st);     // engine dispatch
// LaneConv wide projection on tcgen05 with the A tile in TMEM and TMA-fed pre-split weights (gemm_tc_wide.cu)
int lgcn_split_tf32(const float* w, float* hi, float* lo, int64_t n, cudaStream_t st);
int lgcn_launch_wide_tc(const LinearArgs& a, const float* w_hi, const float* w_lo, cudaStream_t st);
int64_t lgcn_laneconv_fused_aux_bytes(int64_t n_edges);
int lgcn_launch_linear_fused(const LinearArgs& a, cudaStream_t st);
int64_t lgcn_linear_split_bytes();
// split up to 8 weight blocks W_b[n, 0..127] = p[b][n * ldw[b] + 0..127] (n < 128) into hi / lo [n_blocks*128, 128]
struct LgcnSplitList {
  const float* p[8];
  int64_t ldw[8];
  int n_blocks;
};
int lgcn_split_blocks_many(const LgcnSplitList& l, float* hi, float* lo, cudaStream_t st);
// hi / cross images of a flat run of 128-float rows for the aggregate-first kernel (n floats, n % 128 == 0)
int lgcn_split_fused(const float* w, float* hi, float* lo, int64_t n, cudaStream_t st);
int lgcn_launch_laneconv_fused(const float* x, float* out, void* plan, int64_t n_nodes, const int32_t* n_dev,
                               int64_t n_edges, int n_keys, const float* w_hi, const float* w_lo, const float* gn,
                               float* xa, int chain, cudaStream_t st);
// exclusive scan: out[0..n] (n+1 entries) from cnt[0..n); scratch >= 1025 int32
int lgcn_launch_exclusive_scan(const int32_t* cnt, int32_t* out, int64_t n_cap, const int32_t* n_dev, int32_t* scratch,
                               cudaStream_t st);
// zero `bytes` (a multiple of 4) bytes at p with a KERNEL: inside a captured graph a kernel node carries the capture
// stream's priority, a memset node does not (measured: memset nodes of the high-priority branch queued behind every
// CTA of the low-priority ActorNet kernel)
int lgcn_zero_async(void* p, int64_t bytes, cudaStream_t st);
// ---- launchers whose row counts may live in device memory (n_dev != NULL: n_cap is the capacity that sizes the grid)
int lgcn_launch_pack_meta(const float* turn, const float* control, const float* intersect, float* meta, int64_t n_cap,
                          const int32_t* n_dev, cudaStream_t st);
int lgcn_launch_actor_transpose(const float* in, float* out, int64_t n_cap, const int32_t* n_dev, int T, int C,
                                cudaStream_t st);
int lgcn_launch_csr_from_segs(const int64_t* e64, const int64_t* seg_start, int seg_stride, int n_keys, int64_t E_cap,
                              int64_t n_cap, const int32_t* n_dev, int32_t* rowptr, int32_t* col, void* workspace,
                              int32_t* err_flag, cudaStream_t st);
int lgcn_launch_plan_build(const int32_t* rowptr, const int32_t* col, int n_keys, int64_t n_nodes, const int32_t* n_dev,
                           int64_t n_edges, void* plan, cudaStream_t st);
int lgcn_launch_pairs(const float* agt_ctrs, const float* ctx_ctrs, const int32_t* agt_off, const int32_t* ctx_off,
                      int n_scenes, int64_t n_agt_cap, const int32_t* n_agt_dev, float th, int keep_quirk,
                      int32_t* rowptr, void* workspace, int64_t p_cap, int32_t* hi32, int32_t* wi32, int32_t* p_total,
                      int32_t* status, int32_t* p_exact, int overflow_bit, int empty_bit, int64_t n_ctx_hint, cudaStream_t st);
int lgcn_launch_mlp2_in(const float* p, const int32_t* ip, const float* q, const int32_t* iq, const float* W1,
                        const float* b1, float* h, int64_t m_cap, const int32_t* m_dev, cudaStream_t st);
int lgcn_launch_segsum_gn_relu(const float* a, const float* c, const int32_t* rowptr, const float* gamma,
                               const float* beta, float* out, int64_t n_cap, const int32_t* n_dev, cudaStream_t st);
