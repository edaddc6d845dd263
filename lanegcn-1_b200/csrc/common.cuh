// common.cuh — shared helpers of liblgcn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lgcn.h"

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
enum { LGCN_PROF_WIDE = 0, LGCN_PROF_GATHER = 1, LGCN_PROF_CTR2 = 2, LGCN_PROF_ATT = 3, LGCN_PROF_FUSED = 4, LGCN_PROF_KINDS = 8 };
struct LgcnProfScope {
  int slot;
  cudaStream_t st;
  LgcnProfScope(int kind, cudaStream_t st);
  ~LgcnProfScope();
};

static inline int64_t lgcn_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline unsigned lgcn_cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

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
  int dbg;  // ablation switches for profiling (lgcn_debug_flags); 0 in production
  // optional: the weights already split into tf32 hi / lo blocks [n_src*128, 128] (lgcn_split_blocks_many); when
  // NULL the tcgen05 path splits W into a scratch slot at launch
  const float* w_hi;
  const float* w_lo;
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
// split up to 8 weight blocks W_b[n, 0..127] = p[b][n * ldw[b] + 0..127] (n < 128) into hi / lo [n_blocks*128, 128]
struct LgcnSplitList {
  const float* p[8];
  int64_t ldw[8];
  int n_blocks;
};
int lgcn_split_blocks_many(const LgcnSplitList& l, float* hi, float* lo, cudaStream_t st);
int lgcn_launch_laneconv_fused(const float* x, float* out, void* plan, int64_t n_nodes, int64_t n_edges, int n_keys,
                               const float* w_hi, const float* w_lo, const float* gn, float* xa, int chain,
                               cudaStream_t st);
// exclusive scan: out[0..n] (n+1 entries) from cnt[0..n); scratch >= 1025 int32
int lgcn_launch_exclusive_scan(const int32_t* cnt, int32_t* out, int64_t n, int32_t* scratch, cudaStream_t st);
