// laneconv_v2.cu — second generation of the aggregate-first LaneConv block (lanegcn.py:331-362 = :448-479):
//
//     out[n] = relu( GN( relu( GN( sum_k W_k . agg_k[n] ) ) . Wctr2^T ) + X[n] )
//
// Same arithmetic, plan, weight images and tensor-memory layout as k_laneconv_fused<false> (laneconv_fused.cu); what
// changed is WHO does the A feed.  A per-stage clock trace of the first generation (tools/timeline2_fused.py,
// profiles/r2_fused_producer_chain.md) showed that its eight producer warps, which fetched their own source rows
// (table entry -> shuffle -> 64-bit address -> 4 cp.async), converted them and published the stage, ran a SERIAL chain
// of ~890 cycles per 32-float stage against 768 tensor cycles: the feed, not the tensor pipe or L2 (which streams
// > 50 B/clk/SM for this pattern, tools/micro/gather3.cu), bounded the kernel, and every epilogue cycle became pipe
// idle time because the producers had no slack to catch up.  The second epilogue read its residual rows one row per
// lane (32 lines per request: ~4 k cycles of L1 tag time per tile).
//
// Warp roles (512 threads, one persistent CTA per SM, 128 destination rows per tile; setmaxnreg moves registers from
// warps 0-7 (88 each) to the convert warps (168 each)):
//   warp 0      TMA producer of the pre-split weight chunks (hi | lo, 2 x 16 KB boxes, 4-stage ring)
//   warp 1      MMA issuer (one elected thread): A from TENSOR MEMORY, B from the ring; per stage 4 k-steps x 3 products
//               (M=128, N=128, K=8); hi.hi -> main, lo.hi + hi.lo -> cross accumulator
//   warps 4-7   GATHER warps: read the plan's table entries (one coalesced load per key, a key ahead), turn them into
//               row pointers ONCE per key and stream the source rows into a 3-slot X ring with cp.async (8 lanes per
//               128 B row chunk: 4 lines per request); completion is tracked by an mbarrier
//               (cp.async.mbarrier.arrive.noinc), so the consumers never execute cp.async.wait_group.  The residual
//               rows of a tile travel through the same ring as four extra stages.
//   warps 8-15  CONVERT + epilogue warps: wait X-full, read their 64 B, release the slot, split hi / lo, tcgen05.st into
//               the 4-stage A ring, publish.  Thread (q, h, lane) owns row 32q + lane and, of every 32-float K-chunk
//               kc, floats [16h, 16h + 16): the SAME columns 32 kc + 16 h + i in the feed, in both epilogues and in the
//               residual stages, so ctr2's A operand and the residual add need no data movement between threads.
//   TMEM        A ring 4 x (hi 32 | lo 32) = [0,256) | main [256,384) | cross [384,512)
//   smem        W ring 4 x 32 KB | X ring 3 x 16 KB | store staging 8 x 4 KB | norm vectors | GN statistics | barriers
#include <cstring>

#include "tc_common.cuh"

using namespace tc;

long long* lgcn_timeline_buffer();

namespace {

#ifdef LGCN_TIMELINE2
#define V2_TK(i) asm volatile("mov.u32 %0, %%clock;" : "=r"(tk[i]) :: "memory")
#else
#define V2_TK(i) (void)0
#endif

constexpr int kTileM = 128;
constexpr int kStages = 4;                                   // weight ring == A ring depth (shared barriers)
constexpr int kXStages = 3;
constexpr int kWStageBytes = 2 * 128 * 128;                  // hi 16 KB | lo 16 KB
constexpr int kXStageBytes = 128 * 128;                      // 128 rows x 32 floats
constexpr int kSmemX = kStages * kWStageBytes;               // 128 KB
constexpr int kSmemOut = kSmemX + kXStages * kXStageBytes;   // 8 x 4 KB store staging (two 2 KB halves per warp)
constexpr int kSmemGam = kSmemOut + 8 * 4096;                // gamma1 | beta1 | gamma2 | beta2 (4 x 512 B)
constexpr int kSmemStat = kSmemGam + 4 * 512;                // float2 [2][128]
constexpr int kSmemBar = kSmemStat + 2 * 128 * 8;
constexpr int kSmemTotal = kSmemBar + 256;
constexpr int kNumThreads = 512;
constexpr int kGatherWarps = 4;                                 // a warp issues one cp.async per ~55 cycles (tools/micro/gather4.cu):
                                                             // 32 per stage need >= 4 warps to stay below the 768 tensor cycles
constexpr int kRowsPerGW = kTileM / kGatherWarps, kInstrPerGW = kRowsPerGW / 4;
constexpr uint32_t kIdesc = idesc_tf32(128, 128);
constexpr uint32_t kColMain = 256, kColCross = 384;
constexpr int kFlushKeys = 3;   // see laneconv_fused.cu: the main accumulator is flushed into registers every 3 keys

__device__ __forceinline__ float rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

struct V2Args {
  const float* X;        // [M,128] input features (also the residual)
  const float* XA;       // auxiliary rows (multi-source sums)
  const int32_t* tab;    // [n_tiles][n_keys][128]
  const float* gn;       // gamma1 | beta1 | gamma2 | beta2
  int64_t M;             // row capacity ...
  const int32_t* m_dev;  // ... and, when not NULL, the live row count in device memory
  int n_keys;
  int dbg;               // 1: no stores, 4: no MMAs (ablation)
  uint32_t* tl;          // LGCN_TIMELINE2: clock stamps of CTA 0, uint32 [2][1024][8] (MMA thread | convert warp 8)
};

__global__ void __launch_bounds__(kNumThreads, 1)
k_laneconv_v2(const V2Args a, const __grid_constant__ CUtensorMap out_map, const __grid_constant__ CUtensorMap whi_map,
              const __grid_constant__ CUtensorMap wlo_map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  lgcn_pdl_trigger();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  if (sbase & 1023u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_full = sbase + kSmemBar;               // [4]: 8 convert-warp arrives + the TMA expect_tx arrive
  const uint32_t bar_empty = bar_full + 8 * kStages;        // [4]: tcgen05.commit
  const uint32_t bar_xfull = bar_empty + 8 * kStages;       // [3]: 64 cp.async-completion arrives (gather lanes)
  const uint32_t bar_xempty = bar_xfull + 8 * kXStages;     // [3]: 8 convert-warp arrives
  const uint32_t bar_acc_full = bar_xempty + 8 * kXStages;
  const uint32_t bar_acc_empty = bar_acc_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + 192);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, 9);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < kXStages; ++i) {
      mbar_init(bar_xfull + 8 * i, kGatherWarps * 32);
      mbar_init(bar_xempty + 8 * i, 8);
    }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_acc_empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    float* g = reinterpret_cast<float*>(smem + kSmemGam);
    for (int i = threadIdx.x; i < 4 * 128; i += kNumThreads) g[i] = a.gn[i];
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  lgcn_pdl_wait();   // everything above touched only this CTA's shared / tensor memory and the (static) norm vectors

  const int64_t M = lgcn_devn(a.m_dev, a.M);
  const float* __restrict__ X = a.X;
  const float* __restrict__ XA = a.XA;
  const int32_t* __restrict__ tab = a.tab;
  const int n_keys = a.n_keys, nk = a.n_keys + 1;
  const int dbg = a.dbg;
  const int n_tiles = (int)((M + kTileM - 1) / kTileM);
  const int grid = (int)gridDim.x;
  const int keys_per_tile = nk + 1;   // key nk = ctr2 (its weights follow the projections in w_hi / w_lo)

  // 512 threads start with 128 registers each; warp groups 0-1 (TMA, MMA, gather) hand theirs to the convert warps
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");   // the MMA thread's unrolled flush group needs ~100
  if (warp == 0) {
    // =========================================================== TMA producer: weight chunks
    uint32_t phase = 0;
    int ws = 0;
    for (int t = blockIdx.x; t < n_tiles; t += grid) {
      for (int kk = 0; kk < keys_per_tile; ++kk) {
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait(bar_empty + 8 * ws, phase ^ 1);
          if (elect_one()) {
            const uint32_t bar = bar_full + 8 * ws, dst = sbase + ws * kWStageBytes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kWStageBytes) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&whi_map)), "r"(bar), "r"(kc * 32), "r"(kk * 128)
                : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(dst + kWStageBytes / 2), "l"(reinterpret_cast<uint64_t>(&wlo_map)), "r"(bar), "r"(kc * 32), "r"(kk * 128)
                : "memory");
          }
          __syncwarp();
          if (++ws == kStages) {
            ws = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer (one elected thread runs the whole loop)
    // The loop body is one FLUSH GROUP (kFlushKeys keys = 12 stages, fully unrolled): tcgen05.mma takes its descriptors
    // from uniform registers, and re-writing a uniform register that a queued MMA still names stalls the issuing thread
    // until the tensor pipe has drained to it — measured ~450 cycles per loop back-edge whatever the body size (rolled
    // per-stage loop: 1,200 cycles per stage; four unrolled stages: +450 per key; tools/timeline_v2.py).  With the
    // back-edge at the flush boundary the drain coincides with the accumulator hand-over.  The ring slot of a stage is
    // its K-chunk index (every key has 4 stages).  The full barrier of stage s+1 is TESTED (non-blocking) before the
    // MMAs of stage s are issued and the predicate read after them.
    if (elect_one()) {
      uint32_t phase = 0, acc_uses = 0;
#ifdef LGCN_TIMELINE2
      int tls = 0;
#endif
      const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + grid - 1) / grid : 0;
      int64_t stages_left = (int64_t)my_tiles * keys_per_tile * 4;
      asm volatile(".reg .pred lgv2_peek;");
      if (stages_left > 0) {
        mbar_wait(bar_full, 0);
        tc_fence_after();
      }
      const uint32_t d_main = tmem_base + kColMain, d_cross = tmem_base + kColCross;
      // one stage: K-chunk kc (= ring slot) of a key.  restart_all / restart_main: the key's first MMAs overwrite both
      // accumulators / main only (the latter after the convert warps have flushed main: `late` hand-over)
      auto stage = [&](const int kc, const bool restart_all, const bool restart_main, const bool publish, int kk_dbg) {
        const uint32_t a_hi = tmem_base + kc * 64, a_lo = a_hi + 32;
        const uint32_t w_hi = sbase + kc * kWStageBytes, w_lo = w_hi + kWStageBytes / 2;
        const int st2 = (kc + 1) & 3;
        const uint32_t ph2 = kc == 3 ? phase ^ 1 : phase;
        --stages_left;
#ifdef LGCN_TIMELINE2
        if (a.tl && blockIdx.x == 0 && tls < 1024) {
          uint32_t c;
          asm volatile("mov.u32 %0, %%clock;" : "=r"(c) :: "memory");
          a.tl[tls * 8 + 0] = c; a.tl[tls * 8 + 3] = kk_dbg * 4 + kc;
        }
        ++tls;
#endif
        asm volatile("mbarrier.test_wait.parity.shared::cta.b64 lgv2_peek, [%0], %1;" ::"r"(bar_full + 8 * st2), "r"(ph2) : "memory");
        if (kc == 0 && restart_main && !restart_all) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            umma_tf32_ts(d_cross, a_lo + 8 * j, umma_desc(w_hi + j * 32), kIdesc, 1u);
            umma_tf32_ts(d_cross, a_hi + 8 * j, umma_desc(w_lo + j * 32), kIdesc, 1u);
          }
          mbar_wait(bar_acc_empty, (acc_uses & 1) ^ 1);   // the convert warps have added main to their sums
          ++acc_uses;
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_tf32_ts(d_main, a_hi + 8 * j, umma_desc(w_hi + j * 32), kIdesc, j ? 1u : 0u);
        } else if (!(dbg & 4)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool k0 = kc == 0 && j == 0;
            umma_tf32_ts(d_cross, a_lo + 8 * j, umma_desc(w_hi + j * 32), kIdesc, (restart_all && k0) ? 0u : 1u);
            umma_tf32_ts(d_cross, a_hi + 8 * j, umma_desc(w_lo + j * 32), kIdesc, 1u);
            umma_tf32_ts(d_main, a_hi + 8 * j, umma_desc(w_hi + j * 32), kIdesc, (restart_main && k0) ? 0u : 1u);
          }
        }
        umma_commit(bar_empty + 8 * kc);
        if (publish && kc == 3) umma_commit(bar_acc_full);
        uint32_t ready;
        asm volatile("selp.u32 %0, 1, 0, lgv2_peek;" : "=r"(ready));
        if (stages_left > 0) {
          if (!ready) mbar_wait(bar_full + 8 * st2, ph2);
          tc_fence_after();
        }
        if (kc == 3) phase ^= 1;
      };
      auto acc_acquire = [&]() {   // both accumulators have been read by the convert warps
        mbar_wait(bar_acc_empty, (acc_uses & 1) ^ 1);
        ++acc_uses;
        tc_fence_after();
      };
      for (int t = blockIdx.x; t < n_tiles; t += grid) {
#pragma unroll 1
        for (int k0 = 0; k0 < nk; k0 += kFlushKeys) {   // one flush group per iteration
          if (k0 == 0) acc_acquire();
#pragma unroll
          for (int k3 = 0; k3 < kFlushKeys; ++k3) {
            const int kk = k0 + k3;
            if (kk < nk) {
              const bool publish = kk == nk - 1 || k3 == kFlushKeys - 1;
#pragma unroll
              for (int kc = 0; kc < 4; ++kc) stage(kc, k3 == 0 && k0 == 0, k3 == 0, publish, kk);
            }
          }
        }
        acc_acquire();   // ctr2 (key nk): both accumulators restart
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) stage(kc, true, true, true, nk);
      }
    }
    __syncwarp();
  }   // warps 2, 3: no role
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    // =========================================================== gather warps 4..7: source rows -> X ring
    // Lane (g, lane) fetches, per stage, the 16-byte piece (lane & 7) of rows 32 g + 4j + (lane >> 3), j = 0..7.
    const int g = warp - 4, piece = lane & 7, rsub = lane >> 3;
    const uint32_t xbase = sbase + kSmemX;
    uint32_t xphase = 0;
    int xs = 0;
    // source of (tile t, key kk, local row) — key 0 is the row itself
    auto load_idx = [&](int t, int kk, int (&v)[kInstrPerGW]) {
      if (kk >= nk) {
        kk = 0;
        t += grid;
      }
      // row 4j + rsub of this warp's block is live iff 4j < rows_left; one base pointer per key, constant offsets per j
      const int64_t m_first = (int64_t)t * kTileM + g * kRowsPerGW + rsub;
      const int64_t rows_left = t < n_tiles ? M - m_first : 0;
      const int32_t* p = tab + (((int64_t)t * n_keys + (kk - 1)) << 7) + g * kRowsPerGW + rsub;
#pragma unroll
      for (int j = 0; j < kInstrPerGW; ++j) {
        int s = -1;
        if (4 * j < rows_left) s = kk == 0 ? (int)m_first + 4 * j : __ldg(p + 4 * j);
        v[j] = s;
      }
    };
    auto emit = [&](const float* const (&ptr)[kInstrPerGW], uint32_t mask, int kc) {
      mbar_wait(bar_xempty + 8 * xs, xphase ^ 1);
      const uint32_t slot = xbase + xs * kXStageBytes;
#pragma unroll
      for (int j = 0; j < kInstrPerGW; ++j) {
        const int row = g * kRowsPerGW + 4 * j + rsub;
        const uint32_t dst = slot + row * 128 + ((piece ^ (row & 7)) << 4);
        const uint32_t n = (mask >> j) & 1u ? 16u : 0u;   // 0 source bytes: the 16 B are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ptr[j] + kc * 32 + piece * 4), "r"(n) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_xfull + 8 * xs) : "memory");
      if (++xs == kXStages) {
        xs = 0;
        xphase ^= 1;
      }
    };
    auto emit_residual = [&](int t) {   // the tile's own rows, chunk by chunk
      const float* ptr[kInstrPerGW];
      uint32_t mask = 0;
#pragma unroll
      for (int j = 0; j < kInstrPerGW; ++j) {
        const int64_t m = (int64_t)t * kTileM + g * kRowsPerGW + 4 * j + rsub;
        const bool live = m < M;
        ptr[j] = X + (live ? m : 0) * LGCN_C;
        mask |= live ? (1u << j) : 0u;
      }
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) emit(ptr, mask, kc);
    };
    int vnext[kInstrPerGW];
    load_idx((int)blockIdx.x, 0, vnext);
    int prev_t = -1;
    for (int t = blockIdx.x; t < n_tiles; t += grid) {
      for (int kk = 0; kk < nk; ++kk) {
        const float* ptr[kInstrPerGW];
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < kInstrPerGW; ++j) {
          const int v = vnext[j];
          ptr[j] = v >= 0 ? X + (int64_t)v * LGCN_C : v < -1 ? XA + (int64_t)(-2 - v) * LGCN_C : X;
          mask |= v != -1 ? (1u << j) : 0u;
        }
        load_idx(t, kk + 1, vnext);   // a key ahead: the loads are in flight during this key's four stages
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          if (kk == 0 && kc == 3 && prev_t >= 0) emit_residual(prev_t);
          emit(ptr, mask, kc);
        }
      }
      prev_t = t;
    }
    if (prev_t >= 0) emit_residual(prev_t);
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    // =========================================================== convert + epilogue warps 8..15
    const int e = warp - 8, q = e & 3, h = e >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t my_buf = sbase + kSmemOut + e * 4096;
    const uint32_t stat_mine = sbase + kSmemStat + (h * 128 + r) * 8, stat_other = sbase + kSmemStat + ((h ^ 1) * 128 + r) * 8;
    const uint32_t gam = sbase + kSmemGam + h * 64;   // + 0: gamma1, + 512: beta1, + 1024: gamma2, + 1536: beta2
    const uint32_t xrow = sbase + kSmemX + r * 128;   // + slot * kXStageBytes + swizzled piece

    uint32_t a_phase = 0, x_phase = 0, acc_uses = 0;
    int as = 0, xs = 0;
#ifdef LGCN_TIMELINE2
    int tls = 0;
#endif
    // f[16 kc + i] = column 32 kc + 16 h + i of this thread's row: running fp32 sum of the flushed main accumulator
    float f[64];

    // one X-ring stage: this thread's 64 B (floats [16h, 16h+16) of the chunk); the slot is released once all lanes read
    auto xstage = [&](float4(&cur)[4]) {
      mbar_wait(bar_xfull + 8 * xs, x_phase);
      const uint32_t src = xrow + xs * kXStageBytes;
#pragma unroll
      for (int c = 0; c < 4; ++c) cur[c] = ld_shared_f4(src + (((4 * h + c) ^ (r & 7)) << 4));
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_xempty + 8 * xs);
      if (++xs == kXStages) {
        xs = 0;
        x_phase ^= 1;
      }
    };
    // 16 floats -> hi / lo -> TMEM columns [c0, c0+16) (hi) and [c0+32, c0+48) (lo) of A stage `as`
    auto put16 = [&](const float4(&x)[4], int c0) {
#pragma unroll
      for (int gq = 0; gq < 2; ++gq) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 y = x[2 * gq + c];
          const float h0 = rna(y.x), h1 = rna(y.y), h2 = rna(y.z), h3 = rna(y.w);
          hi[4 * c] = __float_as_uint(h0); lo[4 * c] = __float_as_uint(rna(y.x - h0));
          hi[4 * c + 1] = __float_as_uint(h1); lo[4 * c + 1] = __float_as_uint(rna(y.y - h1));
          hi[4 * c + 2] = __float_as_uint(h2); lo[4 * c + 2] = __float_as_uint(rna(y.z - h2));
          hi[4 * c + 3] = __float_as_uint(h3); lo[4 * c + 3] = __float_as_uint(rna(y.w - h3));
        }
        TMEM_ST8(t_lane + as * 64 + c0 + 8 * gq, hi, 0);
        TMEM_ST8(t_lane + as * 64 + 32 + c0 + 8 * gq, lo, 0);
      }
    };
    auto stage_begin = [&]() {
      mbar_wait(bar_empty + 8 * as, a_phase ^ 1);
      tc_fence_after();
    };
    auto stage_end = [&]() {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * as);
      if (++as == kStages) {
        as = 0;
        a_phase ^= 1;
      }
    };
    auto acc_wait = [&]() {
      mbar_wait(bar_acc_full, acc_uses & 1);
      ++acc_uses;
      tc_fence_after();
    };
    auto acc_release = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty);
    };
    auto flush_main = [&]() {   // f += main; the MMA warp restarts main with accumulate = 0
      acc_wait();
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t v[16];
        TMEM_LD16(v, 0, t_lane + kColMain + 32 * kc + 16 * h);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) f[kc * 16 + c] += __uint_as_float(v[c]);
      }
      acc_release();
    };
    auto drain = [&](bool add) {   // f (+)= main + cross, then the accumulators are released to the MMA warp
      acc_wait();
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t v[16], x[16];
        TMEM_LD16(v, 0, t_lane + kColMain + 32 * kc + 16 * h);
        TMEM_LD16(x, 0, t_lane + kColCross + 32 * kc + 16 * h);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float y = __uint_as_float(v[c]) + __uint_as_float(x[c]);
          f[kc * 16 + c] = add ? f[kc * 16 + c] + y : y;
        }
      }
      acc_release();
    };
    auto gn = [&](uint32_t gb) {   // GroupNorm(1) of the row; gb: shared address of gamma (+ 16h floats; beta 512 B behind it)
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        s4[0] += f[c];
        s4[1] += f[c + 1];
        s4[2] += f[c + 2];
        s4[3] += f[c + 3];
      }
      const float mean_h = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / 64.0f);
      float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float d0 = f[c] - mean_h, d1 = f[c + 1] - mean_h, d2 = f[c + 2] - mean_h, d3 = f[c + 3] - mean_h;
        q4[0] = fmaf(d0, d0, q4[0]);
        q4[1] = fmaf(d1, d1, q4[1]);
        q4[2] = fmaf(d2, d2, q4[2]);
        q4[3] = fmaf(d3, d3, q4[3]);
      }
      const float m2_h = (q4[0] + q4[1]) + (q4[2] + q4[3]);
      st_shared_f2(stat_mine, mean_h, m2_h);
      named_bar_sync(1 + q, 64);
      const float2 o = ld_shared_f2(stat_other);
      named_bar_sync(1 + q, 64);   // both halves have read: the slots may be rewritten by the next norm
      const float mean = 0.5f * (mean_h + o.x);
      const float dm = mean_h - o.x;
      const float var = (m2_h + o.y + dm * dm * 32.0f) * (1.0f / 128.0f);
      const float rstd = 1.0f / sqrtf(var + LGCN_GN_EPS);
#pragma unroll
      for (int c = 0; c < 64; c += 4) {   // f[c..c+3] = columns 32 (c >> 4) + 16 h + (c & 15) ...
        const uint32_t at = gb + (32 * (c >> 4) + (c & 15)) * 4;
        const float4 gg = ld_shared_f4(at), b = ld_shared_f4(at + 512);
        f[c] = fmaf((f[c] - mean) * rstd, gg.x, b.x);
        f[c + 1] = fmaf((f[c + 1] - mean) * rstd, gg.y, b.y);
        f[c + 2] = fmaf((f[c + 2] - mean) * rstd, gg.z, b.z);
        f[c + 3] = fmaf((f[c + 3] - mean) * rstd, gg.w, b.w);
      }
    };
    // f[16 kc .. 16 kc + 16) -> one 32-row x 16-column box (2 KB half of the staging buffer, SWIZZLE_64B: rows packed at
    // 64 B, 16-byte chunk index c ^ ((row >> 1) & 3): the eight lanes of a quarter warp hit eight distinct bank groups)
    auto store16 = [&](int kc, int64_t m0) {
      if (dbg & 1) return;
      const uint32_t buf = my_buf + (kc & 1) * 2048;
      if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c)
        st_shared_f4(buf + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4),
                     make_float4(f[16 * kc + 4 * c], f[16 * kc + 4 * c + 1], f[16 * kc + 4 * c + 2], f[16 * kc + 4 * c + 3]));
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&out_map)),
                     "r"(buf), "r"(32 * kc + 16 * h), "r"((int32_t)(m0 + q * 32))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    // Second epilogue of a tile (ctr2 accumulators -> GroupNorm + residual + ReLU -> store).  It runs three stages into
    // the NEXT tile (those stages are produced while the ctr2 MMAs still execute); the residual arrives through the X
    // ring as four stages holding exactly this thread's columns.
    auto finish_tile = [&](int64_t pm0) {
      drain(false);
      gn(gam + 1024);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        float4 res[4];
        xstage(res);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          f[16 * kc + 4 * c] = fmaxf(f[16 * kc + 4 * c] + res[c].x, 0.f);
          f[16 * kc + 4 * c + 1] = fmaxf(f[16 * kc + 4 * c + 1] + res[c].y, 0.f);
          f[16 * kc + 4 * c + 2] = fmaxf(f[16 * kc + 4 * c + 2] + res[c].z, 0.f);
          f[16 * kc + 4 * c + 3] = fmaxf(f[16 * kc + 4 * c + 3] + res[c].w, 0.f);
        }
        store16(kc, pm0);
      }
    };

    bool pending = false;
    int64_t pending_m0 = 0;
    asm volatile(".reg .pred lgv2_px;\n\t.reg .pred lgv2_pa;");
    for (int t = blockIdx.x; t < n_tiles; t += grid) {
      const int64_t m0 = (int64_t)t * kTileM;
      if (!pending) {
#pragma unroll
        for (int c = 0; c < 64; ++c) f[c] = 0.f;
      }
      for (int kk = 0; kk < nk; ++kk) {
        // flush of the key group that ended at key kk-1: at K-chunk 2 of this key, i.e. three stages late, so the A
        // ring is full again when the MMA warp resumes
        const bool flush_here = kk > 0 && kk % kFlushKeys == 0;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          if (kc == 3 && kk == 0 && pending) {
            finish_tile(pending_m0);
            pending = false;
#pragma unroll
            for (int c = 0; c < 64; ++c) f[c] = 0.f;
          }
#ifdef LGCN_TIMELINE2
          uint32_t tk0, tk1, tk2, tk3, tk4;
#define V2_TS(v) asm volatile("mov.u32 %0, %%clock;" : "=r"(v) :: "memory")
#else
#define V2_TS(v) (void)0
#endif
          V2_TS(tk0);
          // both barriers are TESTED first (non-blocking), the predicates consumed where the data is needed
          asm volatile("mbarrier.test_wait.parity.shared::cta.b64 lgv2_px, [%0], %1;" ::"r"(bar_xfull + 8 * xs), "r"(x_phase) : "memory");
          asm volatile("mbarrier.test_wait.parity.shared::cta.b64 lgv2_pa, [%0], %1;" ::"r"(bar_empty + 8 * as), "r"(a_phase ^ 1) : "memory");
          uint32_t okx, oka;
          asm volatile("selp.u32 %0, 1, 0, lgv2_px;" : "=r"(okx));
          if (!okx) mbar_wait(bar_xfull + 8 * xs, x_phase);
          float4 cur[4];
          {
            const uint32_t src = xrow + xs * kXStageBytes;
#pragma unroll
            for (int c = 0; c < 4; ++c) cur[c] = ld_shared_f4(src + (((4 * h + c) ^ (r & 7)) << 4));
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xempty + 8 * xs);
            if (++xs == kXStages) {
              xs = 0;
              x_phase ^= 1;
            }
          }
          V2_TS(tk1);
          asm volatile("selp.u32 %0, 1, 0, lgv2_pa;" : "=r"(oka));
          if (!oka) mbar_wait(bar_empty + 8 * as, a_phase ^ 1);
          tc_fence_after();
          V2_TS(tk2);
          put16(cur, h * 16);
          V2_TS(tk3);
          stage_end();
          V2_TS(tk4);
#ifdef LGCN_TIMELINE2
          if (a.tl && blockIdx.x == 0 && e == 0 && lane == 0 && tls < 1024) {
            uint32_t* t2 = a.tl + 8192 + tls * 8;
            t2[0] = tk0; t2[1] = tk1; t2[2] = tk2; t2[3] = tk3; t2[4] = tk4; t2[5] = kk * 4 + kc;
          }
          ++tls;
#endif
          if (kc == 2 && flush_here) flush_main();
        }
      }
      drain(true);
      gn(gam);
#pragma unroll
      for (int c = 0; c < 64; ++c) f[c] = fmaxf(f[c], 0.f);
      // ---- ctr2: h = f is the A operand; every warp holds 16 floats of each K-chunk, exactly as in the feed
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        float4 x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = make_float4(f[16 * kc + 4 * c], f[16 * kc + 4 * c + 1], f[16 * kc + 4 * c + 2], f[16 * kc + 4 * c + 3]);
        stage_begin();
        put16(x, h * 16);
        stage_end();
      }
      pending = true;
      pending_m0 = m0;
    }
    if (pending) finish_tile(pending_m0);
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

}  // namespace

// One LaneConv block (chain form) on the second-generation kernel.  The caller has run k_multi_sum (xa) already.
int lgcn_launch_laneconv_v2(const float* x, const float* xa, const int32_t* tab, float* out, int64_t n_nodes,
                            const int32_t* n_dev, int n_keys, const float* w_hi, const float* w_lo, const float* gn,
                            cudaStream_t st) {
  if (n_nodes <= 0) return 0;
  if (first_use(kFamFusedV2))
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_laneconv_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  const int nkw = n_keys + 2;
  CUtensorMap map, mhi, mlo;
  if (int rc = make_map_2d_sw64(&map, out, LGCN_C, n_nodes, LGCN_C, 32)) return rc;
  if (int rc = make_map_2d(&mhi, w_hi, LGCN_C, (int64_t)nkw * LGCN_C, LGCN_C, 32, 128)) return rc;
  if (int rc = make_map_2d(&mlo, w_lo, LGCN_C, (int64_t)nkw * LGCN_C, LGCN_C, 32, 128)) return rc;
  V2Args a;
  memset(&a, 0, sizeof(a));
  a.X = x; a.XA = xa; a.tab = tab; a.gn = gn; a.M = n_nodes; a.m_dev = n_dev; a.n_keys = n_keys; a.dbg = lgcn_debug_get();
  a.tl = (a.dbg & 256) ? reinterpret_cast<uint32_t*>(lgcn_timeline_buffer()) : nullptr;
  const int64_t n_tiles = (n_nodes + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(n_tiles < num_sms() ? n_tiles : num_sms());
  LGCN_CUDA_OK(lgcn_launch_pdl(k_laneconv_v2, grid, kNumThreads, kSmemTotal, st, a, map, mhi, mlo));
  LGCN_LAUNCH_OK();
  return 0;
}
