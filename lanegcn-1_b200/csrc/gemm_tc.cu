// gemm_tc.cu — tcgen05 (5th-gen tensor core) implementation of lgcn_linear128 (engine 1), sm_100a only.
//
//   out[m, ob*128 + c] = epi( sum_s A_s[idx_s[m], :] . W[ob*128 + c, s*128:(s+1)*128] )      fp32 in, fp32 out
//
// Precision: 3xTF32.  Every fp32 operand x is split into hi = tf32(x) and lo = tf32(x - hi) (both rounded to
// nearest), and   D = sum A_hi.B_hi  +  sum (A_lo.B_hi + A_hi.B_lo)   is accumulated in fp32 in TMEM — in TWO
// accumulators.  Measured on B200 (tools/error_budget.py): the tensor core truncates toward zero once per MMA
// instruction (about -1.3e-8 relative per instruction), so folding the 32 cross-term instructions into the
// main accumulator triples its bias; a separate cross accumulator (2^-11 smaller, so its own truncation is
// negligible) is added in the epilogue with one IEEE fp32 add.  Result: ~21 mantissa bits, inside the
// 1e-4 / 1e-5 budget where plain TF32 is not.
//
// Structure (one persistent CTA per SM, 288 threads, warp-specialised):
//   warps 0-3  producers.  Load the 128-row A tile (optionally row-gathered) with coalesced 128-bit loads,
//              split it into hi/lo and store both into shared memory in the K-major SWIZZLE_128B layout that
//              the UMMA smem descriptor expects (4 K-chunks of 32 floats; chunk = [128 rows][128 B]).  A
//              stays RESIDENT for all output blocks of the tile (the wide LaneConv projection has 15).  Then
//              stream the weight tile of each output block the same way, one 32-float K-chunk per pipeline
//              stage (hi+lo = 32 KB/stage, 2 stages).
//   warp 8     MMA issuer (one elected lane): per stage 4 k-steps x 3 tcgen05.mma.kind::tf32 (M=128, N=128,
//              K=8), accumulators in TMEM (2 stages x [main 128 | cross 128] columns: double-buffered so the
//              epilogue of tile i overlaps the MMAs of tile i+1); tcgen05.commit hands smem stages back and
//              publishes accumulators.
//   warps 4-7  epilogue.  tcgen05.ld the accumulator row (thread == row, so GroupNorm(1) statistics are
//              thread-local), apply GN / ReLU / residual / ReLU, stage 32x32 fp32 blocks in swizzled smem and
//              write them with TMA bulk tensor stores (coalesced 128 B rows, M-tail clipped by the tensor map).
// Shared memory: A hi+lo 128 KB | B 2 x 32 KB | store staging 32 KB | barriers.  TMEM: all 512 columns.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kChunkBytes = kTileM * 128;           // one K-chunk (32 floats) of a 128-row operand: 16 KB
constexpr int kSmemAHi = 0;
constexpr int kSmemALo = 4 * kChunkBytes;           // 64 KB
constexpr int kSmemB = 8 * kChunkBytes;             // 128 KB: stages of [hi 16 KB | lo 16 KB]
constexpr int kBStages = 2;
constexpr int kSmemOut = kSmemB + kBStages * 2 * kChunkBytes;   // 192 KB: 4 warps x 2 x 4 KB
constexpr int kSmemBar = kSmemOut + 4 * 2 * 4096;   // 224 KB
constexpr int kSmemTotal = kSmemBar + 256;
constexpr int kSmemAlloc = kSmemTotal + 1024;       // slack for manual 1024-byte alignment
constexpr int kNumThreads = 288;
constexpr uint32_t kTmemCols = 512;   // 2 stages x (main + cross accumulator) x 128 columns
// instruction descriptor, kind::tf32: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), both K-major,
// N=128 (16<<17), M=128 (8<<24)                                            (cute/arch/mma_sm100_desc.hpp:412)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp:91): start>>4 | LBO=1 |
// SBO = 1024 B (8 rows x 128 B) | version 1 | layout SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

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

// hi = tf32(x), lo = tf32(x - hi), both rounded to nearest (ties away).  16-byte stores at the SWIZZLE_128B position of
// (row, 16-byte chunk c) inside a [rows][128 B] K-chunk block: chunk index XOR (row & 7).
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
  return __uint_as_float(t);
}
__device__ __forceinline__ void split_store(uint8_t* hi_blk, uint8_t* lo_blk, int row, int c, float4 x) {
  float4 h, l;
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x.x)); h.x = __uint_as_float(t); l.x = tf32_rna(x.x - h.x);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x.y)); h.y = __uint_as_float(t); l.y = tf32_rna(x.y - h.y);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x.z)); h.z = __uint_as_float(t); l.z = tf32_rna(x.z - h.z);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x.w)); h.w = __uint_as_float(t); l.w = tf32_rna(x.w - h.w);
  const int off = row * 128 + ((c ^ (row & 7)) << 4);
  *reinterpret_cast<float4*>(hi_blk + off) = h;
  *reinterpret_cast<float4*>(lo_blk + off) = l;
}

__global__ void __launch_bounds__(kNumThreads, 1)
k_linear_tc(const LinearArgs a, const __grid_constant__ CUtensorMap out_map) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // barriers (8 bytes each) and the TMEM base address slot
  const uint32_t bar_a_full = sbase + kSmemBar + 0, bar_a_empty = sbase + kSmemBar + 8;
  const uint32_t bar_b_full = sbase + kSmemBar + 16;    // [kBStages]
  const uint32_t bar_b_empty = sbase + kSmemBar + 32;   // [kBStages]
  const uint32_t bar_acc_full = sbase + kSmemBar + 48;  // [2]
  const uint32_t bar_acc_empty = sbase + kSmemBar + 64; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + 96);

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 128);
    mbar_init(bar_a_empty, 1);
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(bar_b_full + 8 * i, 128);
      mbar_init(bar_b_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_tiles = (a.m + kTileM - 1) / kTileM;
  const int64_t ldw = (int64_t)a.n_src * LGCN_C + a.ks;
  const int nob = a.n_out_blocks;

  if (warp < 4) {
    // =========================================================== producers
    const int tid = threadIdx.x;  // 0..127
    uint32_t a_empty_phase = 0, b_phase = 0;
    int b_stage = 0;
    const int ch16 = lane & 7, kc_of_lane = lane >> 3;  // A rows: lane l owns float4 #l of the 512-byte row
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t m0 = t * kTileM;
      for (int s = 0; s < a.n_src; ++s) {
        // ---- A_s tile -> smem (hi/lo), once per (tile, source); waits until the MMAs that read the previous
        // contents have retired.
        mbar_wait(bar_a_empty, a_empty_phase ^ 1);
        a_empty_phase ^= 1;
        const float* __restrict__ src = a.a[s];
        const int32_t* __restrict__ idx = a.idx[s];
#pragma unroll 1
        for (int r0 = warp * 32; r0 < warp * 32 + 32; r0 += 8) {
          float4 x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t m = m0 + r0 + i;
            if (m < a.m) {
              const int64_t row = idx ? (int64_t)__ldg(idx + m) : m;
              x[i] = __ldg(reinterpret_cast<const float4*>(src + row * LGCN_C) + lane);
            } else {
              x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            split_store(smem + kSmemAHi + kc_of_lane * kChunkBytes, smem + kSmemALo + kc_of_lane * kChunkBytes,
                        r0 + i, ch16, x[i]);
        }
        fence_proxy_async();
        mbar_arrive(bar_a_full);
        // ---- weight tiles: for every output block, 4 K-chunk stages
        for (int ob = 0; ob < nob; ++ob) {
          const float* __restrict__ wbase = a.W + ((int64_t)ob * LGCN_C) * ldw + (int64_t)s * LGCN_C;
          for (int kc = 0; kc < 4; ++kc) {
            mbar_wait(bar_b_empty + 8 * b_stage, b_phase ^ 1);
            uint8_t* hi_blk = smem + kSmemB + b_stage * 2 * kChunkBytes;
            uint8_t* lo_blk = hi_blk + kChunkBytes;
            float4 x[8];
            // thread -> (row n = tid/8 + 16*i, 16-byte chunk tid%8): 8 lanes read one 128-byte row segment
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int n = (tid >> 3) + 16 * i;
              x[i] = __ldg(reinterpret_cast<const float4*>(wbase + (int64_t)n * ldw + kc * 32) + (tid & 7));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) split_store(hi_blk, lo_blk, (tid >> 3) + 16 * i, tid & 7, x[i]);
            fence_proxy_async();
            mbar_arrive(bar_b_full + 8 * b_stage);
            if (++b_stage == kBStages) {
              b_stage = 0;
              b_phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 8) {
    // =========================================================== MMA issuer
    uint32_t a_full_phase = 0, b_phase = 0, acc_phase = 0;
    int b_stage = 0, acc_stage = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int s = 0; s < a.n_src; ++s) {
        mbar_wait(bar_a_full, a_full_phase);
        a_full_phase ^= 1;
        for (int ob = 0; ob < nob; ++ob) {
          if (s == 0) {  // a fresh accumulator: wait until the epilogue has drained this TMEM stage
            mbar_wait(bar_acc_empty + 8 * acc_stage, acc_phase ^ 1);
          }
          const uint32_t d_main = tmem_base + acc_stage * 256, d_cross = d_main + 128;
          for (int kc = 0; kc < 4; ++kc) {
            mbar_wait(bar_b_full + 8 * b_stage, b_phase);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_hi = sbase + kSmemAHi + kc * kChunkBytes, a_lo = sbase + kSmemALo + kc * kChunkBytes;
              const uint32_t b_hi = sbase + kSmemB + b_stage * 2 * kChunkBytes, b_lo = b_hi + kChunkBytes;
#pragma unroll
              for (int j = 0; j < 4; ++j) {  // K = 8 floats = 32 bytes per instruction
                const uint32_t first = (s == 0 && kc == 0 && j == 0) ? 0u : 1u;
                umma_tf32(d_cross, umma_desc(a_lo + 32 * j), umma_desc(b_hi + 32 * j), first);
                umma_tf32(d_cross, umma_desc(a_hi + 32 * j), umma_desc(b_lo + 32 * j), 1u);
                umma_tf32(d_main, umma_desc(a_hi + 32 * j), umma_desc(b_hi + 32 * j), first);
              }
              umma_commit(bar_b_empty + 8 * b_stage);  // smem stage reusable once these MMAs retire
              if (kc == 3 && s == a.n_src - 1) umma_commit(bar_acc_full + 8 * acc_stage);
              if (kc == 3 && ob == nob - 1) umma_commit(bar_a_empty);
            }
            __syncwarp();
            if (++b_stage == kBStages) {
              b_stage = 0;
              b_phase ^= 1;
            }
          }
          if (s == a.n_src - 1) {  // accumulator published: move to the other TMEM stage
            if (++acc_stage == 2) {
              acc_stage = 0;
              acc_phase ^= 1;
            }
          }
        }
      }
    }
  } else {
    // =========================================================== epilogue (warps 4..7 -> TMEM lane quarters 0..3)
    const int q = warp - 4;
    uint8_t* stage_buf = smem + kSmemOut + q * 2 * 4096;
    uint32_t acc_phase = 0;
    int acc_stage = 0, out_buf = 0;
    const bool gn = a.flags & LGCN_EPI_GN;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t m0 = t * kTileM;
      const int64_t m = m0 + q * 32 + lane;  // this thread's output row
      for (int ob = 0; ob < nob; ++ob) {
        mbar_wait(bar_acc_full + 8 * acc_stage, acc_phase);
        tc_fence_after();
        uint32_t v[128];
        const uint32_t taddr = tmem_base + acc_stage * 256 + ((uint32_t)(q * 32) << 16);
        TMEM_LD32(v, 0, taddr);
        TMEM_LD32(v, 32, taddr + 32);
        TMEM_LD32(v, 64, taddr + 64);
        TMEM_LD32(v, 96, taddr + 96);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {  // + cross-term accumulator (columns 128..255 of the stage)
          uint32_t x[32];
          TMEM_LD32(x, 0, taddr + 128 + cb * 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < 32; ++c) v[cb * 32 + c] = __float_as_uint(__uint_as_float(v[cb * 32 + c]) + __uint_as_float(x[c]));
        }
        tc_fence_before();
        mbar_arrive(bar_acc_empty + 8 * acc_stage);  // TMEM stage may be overwritten by the next MMAs
        if (++acc_stage == 2) {
          acc_stage = 0;
          acc_phase ^= 1;
        }
        float* f = reinterpret_cast<float*>(v);
        if (a.ks > 0 && m < a.m) {  // rank-ks update: the 4 extra input columns of A2M.meta
          const float4 x = __ldg(reinterpret_cast<const float4*>(a.xs + m * 4));
          const float* wx = a.W + (int64_t)a.n_src * LGCN_C;
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(wx + (int64_t)c * ldw));
            f[c] = fmaf(x.w, w.w, fmaf(x.z, w.z, fmaf(x.y, w.y, fmaf(x.x, w.x, f[c]))));
          }
        }
        if (gn) {  // GroupNorm(1 group): the thread owns the whole 128-channel row
          float sum = 0.f;
#pragma unroll
          for (int c = 0; c < 128; ++c) sum += f[c];
          const float mean = sum * (1.0f / 128.0f);
          float sq = 0.f;
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            f[c] -= mean;
            sq = fmaf(f[c], f[c], sq);
          }
          const float rstd = 1.0f / sqrtf(sq * (1.0f / 128.0f) + LGCN_GN_EPS);
#pragma unroll
          for (int c = 0; c < 128; c += 4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta + c));
            f[c] = fmaf(f[c] * rstd, g.x, b.x);
            f[c + 1] = fmaf(f[c + 1] * rstd, g.y, b.y);
            f[c + 2] = fmaf(f[c + 2] * rstd, g.z, b.z);
            f[c + 3] = fmaf(f[c + 3] * rstd, g.w, b.w);
          }
        }
        if (a.flags & LGCN_EPI_RELU1) {
#pragma unroll
          for (int c = 0; c < 128; ++c) f[c] = fmaxf(f[c], 0.f);
        }
        if ((a.flags & LGCN_EPI_RES) && m < a.m) {
          const float4* r = reinterpret_cast<const float4*>(a.res + m * LGCN_C);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float4 x = __ldg(r + c);
            f[4 * c] += x.x; f[4 * c + 1] += x.y; f[4 * c + 2] += x.z; f[4 * c + 3] += x.w;
          }
        }
        if (a.flags & LGCN_EPI_RELU2) {
#pragma unroll
          for (int c = 0; c < 128; ++c) f[c] = fmaxf(f[c], 0.f);
        }
        // ---- 4 blocks of 32 rows x 32 columns: swizzled smem staging -> TMA store (rows >= M are clipped)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint8_t* buf = stage_buf + out_buf * 4096;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // buffer free again
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                make_float4(f[cb * 32 + 4 * c], f[cb * 32 + 4 * c + 1], f[cb * 32 + 4 * c + 2], f[cb * 32 + 4 * c + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&out_map)),
                         "r"(smem_u32(buf)), "r"(ob * 128 + cb * 32), "r"((int32_t)(m0 + q * 32))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          out_buf ^= 1;
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int g_num_sms = 0;
bool g_attr_set = false;

}  // namespace

int lgcn_launch_linear_tc(const LinearArgs& a, cudaStream_t st) {
  if (a.m <= 0) return 0;
  EncodeTiledFn encode = get_encode();
  LGCN_CHECK_ARG(encode != nullptr, "linear128(tcgen05): cuTensorMapEncodeTiled not available from the driver");
  LGCN_CHECK_ARG(a.n_src == 1 || a.n_out_blocks == 1, "linear128(tcgen05): several sources need n_out_blocks == 1");
  LGCN_CHECK_ARG((reinterpret_cast<uintptr_t>(a.out) & 15) == 0, "linear128(tcgen05): out must be 16-byte aligned");
  if (!g_attr_set) {
    int dev = 0;
    LGCN_CUDA_OK(cudaGetDevice(&dev));
    LGCN_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    g_attr_set = true;
  }
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)a.n_out_blocks * LGCN_C, (cuuint64_t)a.m};
  const cuuint64_t strides[1] = {(cuuint64_t)a.ldo * sizeof(float)};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.out, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LGCN_CHECK_ARG(r == CUDA_SUCCESS, "linear128(tcgen05): cuTensorMapEncodeTiled failed (%d)", (int)r);
  const int64_t n_tiles = (a.m + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(n_tiles < g_num_sms ? n_tiles : g_num_sms);
  k_linear_tc<<<grid, kNumThreads, kSmemAlloc, st>>>(a, map);
  LGCN_LAUNCH_OK();
  return 0;
}
