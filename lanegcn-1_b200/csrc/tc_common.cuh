// tc_common.cuh — PTX wrappers shared by the tcgen05 GEMM kernels (sm_100a): mbarriers, proxy / tcgen05 fences,
// UMMA shared-memory descriptors, tcgen05.mma / commit / ld, the tf32 hi/lo split, the tensor-map encoder.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tc {

// One lane of a CONVERGED warp.  Issue tcgen05.mma / commit / TMA under `if (elect_one())`, not `if (lane == 0)`:
// under a lane-id branch ptxas treats the uniform-register operands of UTCHMMA / UTMASTG as divergent and wraps
// every single instruction in an ELECT + BRA.U.ANY waterfall loop (measured: the MMA issue thread, not the tensor
// pipe, bounded the GEMM); elect.sync is recognised as a uniform single-thread region.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol bug must surface as a launch failure, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
// Pure polling variant (mbarrier.test_wait never suspends the thread).
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void st_shared_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_shared_f2(uint32_t addr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 ld_shared_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp:91): start>>4 | LBO=1 |
// SBO = 1024 B (8 rows x 128 B) | version 1 | layout SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem: row -> lane, k -> column, one 32-bit element per column] . B[smem desc]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor, kind::tf32: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), both K-major, N>>3 at bit 17,
// M>>4 at bit 24                                                           (cute/arch/mma_sm100_desc.hpp:412)
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem: row -> lane, TWO bf16 per 32-bit column (k even in the low half)] . B[smem desc]^T, K = 16
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor, kind::f16 with bf16 operands: D=F32 (1<<4), A=BF16 (1<<7), B=BF16 (1<<10), both K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// instruction descriptor, kind::f16 with fp16 operands (format 0), fp32 accumulation
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// {hi, lo} -> two fp16 (round to nearest even, saturating to +-65504 instead of inf): `lo` in the low half
__device__ __forceinline__ uint32_t pack_f16(float hi, float lo) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// {hi, lo} -> one 32-bit word of two bf16 (round to nearest even): `lo` in the low half
__device__ __forceinline__ uint32_t pack_bf16(float hi, float lo) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TMEM_ST32(taddr, v, base)                                                                             \
  asm volatile(                                                                                               \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"  \
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"                                         \
      ::"r"(v[base + 0]), "r"(v[base + 1]), "r"(v[base + 2]), "r"(v[base + 3]), "r"(v[base + 4]),             \
        "r"(v[base + 5]), "r"(v[base + 6]), "r"(v[base + 7]), "r"(v[base + 8]), "r"(v[base + 9]),             \
        "r"(v[base + 10]), "r"(v[base + 11]), "r"(v[base + 12]), "r"(v[base + 13]), "r"(v[base + 14]),        \
        "r"(v[base + 15]), "r"(v[base + 16]), "r"(v[base + 17]), "r"(v[base + 18]), "r"(v[base + 19]),        \
        "r"(v[base + 20]), "r"(v[base + 21]), "r"(v[base + 22]), "r"(v[base + 23]), "r"(v[base + 24]),        \
        "r"(v[base + 25]), "r"(v[base + 26]), "r"(v[base + 27]), "r"(v[base + 28]), "r"(v[base + 29]),        \
        "r"(v[base + 30]), "r"(v[base + 31]), "r"(taddr)                                                      \
      : "memory")

#define TMEM_ST16(taddr, v, base)                                                                             \
  asm volatile(                                                                                               \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"  \
      ::"r"(v[base + 0]), "r"(v[base + 1]), "r"(v[base + 2]), "r"(v[base + 3]), "r"(v[base + 4]),             \
        "r"(v[base + 5]), "r"(v[base + 6]), "r"(v[base + 7]), "r"(v[base + 8]), "r"(v[base + 9]),             \
        "r"(v[base + 10]), "r"(v[base + 11]), "r"(v[base + 12]), "r"(v[base + 13]), "r"(v[base + 14]),        \
        "r"(v[base + 15]), "r"(taddr)                                                                         \
      : "memory")

#define TMEM_ST8(taddr, v, base)                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"                       \
               ::"r"(v[base + 0]), "r"(v[base + 1]), "r"(v[base + 2]), "r"(v[base + 3]), "r"(v[base + 4]),    \
                 "r"(v[base + 5]), "r"(v[base + 6]), "r"(v[base + 7]), "r"(taddr)                             \
               : "memory")

#define TMEM_LD16(v, base, taddr)                                                                             \
  asm volatile(                                                                                               \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"  \
      : "=r"(v[base + 0]), "=r"(v[base + 1]), "=r"(v[base + 2]), "=r"(v[base + 3]), "=r"(v[base + 4]),        \
        "=r"(v[base + 5]), "=r"(v[base + 6]), "=r"(v[base + 7]), "=r"(v[base + 8]), "=r"(v[base + 9]),        \
        "=r"(v[base + 10]), "=r"(v[base + 11]), "=r"(v[base + 12]), "=r"(v[base + 13]), "=r"(v[base + 14]),   \
        "=r"(v[base + 15])                                                                                    \
      : "r"(taddr))

#define TMEM_LD32(v, base, taddr)                                                                             \
  asm volatile(                                                                                               \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,"  \
      "%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                      \
      : "=r"(v[base + 0]), "=r"(v[base + 1]), "=r"(v[base + 2]), "=r"(v[base + 3]), "=r"(v[base + 4]),        \
        "=r"(v[base + 5]), "=r"(v[base + 6]), "=r"(v[base + 7]), "=r"(v[base + 8]), "=r"(v[base + 9]),        \
        "=r"(v[base + 10]), "=r"(v[base + 11]), "=r"(v[base + 12]), "=r"(v[base + 13]), "=r"(v[base + 14]),   \
        "=r"(v[base + 15]), "=r"(v[base + 16]), "=r"(v[base + 17]), "=r"(v[base + 18]), "=r"(v[base + 19]),   \
        "=r"(v[base + 20]), "=r"(v[base + 21]), "=r"(v[base + 22]), "=r"(v[base + 23]), "=r"(v[base + 24]),   \
        "=r"(v[base + 25]), "=r"(v[base + 26]), "=r"(v[base + 27]), "=r"(v[base + 28]), "=r"(v[base + 29]),   \
        "=r"(v[base + 30]), "=r"(v[base + 31])                                                                \
      : "r"(taddr))

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
  return __uint_as_float(t);
}
// hi = tf32(x), lo = tf32(x - hi).  16-byte shared stores at the SWIZZLE_128B position of (row, 16-byte chunk c)
// inside a [rows][128 B] K-chunk block: chunk index XOR (row & 7).
__device__ __forceinline__ void split_store(uint32_t hi_blk, uint32_t lo_blk, int row, int c, float4 x) {
  float4 h, l;
  h.x = tf32_rna(x.x); l.x = tf32_rna(x.x - h.x);
  h.y = tf32_rna(x.y); l.y = tf32_rna(x.y - h.y);
  h.z = tf32_rna(x.z); l.z = tf32_rna(x.z - h.z);
  h.w = tf32_rna(x.w); l.w = tf32_rna(x.w - h.w);
  const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
  st_shared_f4(hi_blk + off, h);
  st_shared_f4(lo_blk + off, l);
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();
// 2-D fp32 tensor map over out[m, cols] (row stride ldo floats) with a 32 x 32 box, SWIZZLE_128B (TMA stores)
int make_out_map(CUtensorMap* map, float* out, int64_t cols, int64_t rows, int64_t ldo);
// generic 2-D fp32 map: [rows, cols] with row stride ld floats, box {box_cols, box_rows}, SWIZZLE_128B
int make_map_2d(CUtensorMap* map, const float* base, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows);
// 2-D fp32 map with a 16-float (64 B) x box_rows box, SWIZZLE_64B
int make_map_2d_sw64(CUtensorMap* map, const float* base, int64_t cols, int64_t rows, int64_t ld, int box_rows);
// the same over a bf16 tensor (cols / ld / box_cols in bf16 elements)
int make_map_2d_bf16(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows);
// Per-DEVICE cached state (one process may drive several GPUs: the Python layer keys its workspaces per device and
// wraps calls in torch.cuda.device(dev)).  num_sms(): SM count of the CURRENT device.  first_use(family): true
// exactly once per (current device, kernel family) — cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device.
enum { kFamLinearTc = 0, kFamWideTc = 1, kFamFused = 2, kFamFusedV2 = 3, kFamCount = 4 };
constexpr int kMaxDevices = 64;
int current_device();
int num_sms();
bool first_use(int family);

}  // namespace tc
