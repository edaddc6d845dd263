// laneconv_fused.cu — one LaneConv block (lanegcn.py:331-362 = :448-479) as ONE tcgen05 kernel, aggregate-first:
//
//     out[n] = relu( GN( relu( GN( sum_k W_k . agg_k[n] ) ) . Wctr2^T ) + X[n] ),
//     agg_0[n] = X[n] (ctr),   agg_k[n] = sum over edges e of key k with destination n of X[src(e)].
//
// The split path (gemm_tc_wide.cu + laneconv.cu) materialises Y = X . Wcat^T ([N, 1920] fp32: 1.49 GB written and
// 1.31 GB gathered back per block at batch 128); here the neighbour rows of X are gathered straight into the A operand
// and Y never exists: HBM traffic per block drops from ~3.2 GB to ~0.3 GB, the arithmetic (K = 15 x 128 per output
// row, 3xTF32) is unchanged.  Sums are linear, so summing the source rows first and multiplying once per key is the
// same map; the fp32 rounding differs from the reference's per-edge order at the 1e-7 level (parity tests hold both
// paths to the same 1e-4 / 1e-5 tolerance against the oracle).
//
// Plan (static per graph, lgcn_laneconv_plan_build): tab[tile][key][128] = the ONE source row of (destination row,
// key) (-1: none).  (row, key) pairs with several sources (3 % on lane graphs: merges, dilated scales) get an
// auxiliary row: k_multi_sum writes XA[i] = sum of their sources (CSR order) before each block and tab holds -2-i,
// so the hot loop has exactly one 64-byte load per thread and stage, with no data-dependent trip counts.
//
// Kernel: persistent, one CTA per SM, 128 destination rows per tile, 4 x (n_keys+1) [+4 for ctr2] stages per tile;
// one stage = one 32-float K-chunk of one key.  One full / one empty mbarrier per stage serve both rings:
//   warp 0      TMA producer: W_k[:, 32 kc .. +32) hi | lo (pre-split tf32, 2 x 16 KB boxes) into a 4-stage ring
//   warp 1      MMA issuer (ONE elected thread runs the whole loop): A from TENSOR MEMORY (row -> lane, k -> column), B
//               from the ring; per stage 4 k-steps x 3 products (M=128, N=128, K=8): lo.hi + hi.lo -> cross accumulator,
//               hi.hi -> main accumulator; one tcgen05.commit per stage.  The next stage's full barrier is tested
//               (non-blocking) before the 12 MMAs and the predicate read after them.
//   warps 4-11  A producers + epilogues.  Warp (q = w&3, h = w>>2 & 1) owns rows 32q..32q+31 (TMEM lane quarter) and
//               floats [16h, 16h+16) of every chunk.  The source rows arrive by cp.async (4 lanes per 64 B slice, three
//               stages in flight, XOR-swizzled so that writes and reads are conflict-free), the owner lane reads its
//               64 B back, splits hi/lo in registers and stores them with tcgen05.st into a 4-stage A ring in TMEM.
//               The main accumulator is flushed into registers every kFlushKeys keys (see below).  After the last key
//               the warps drain the accumulators (columns [64h, 64h+64)), GroupNorm (Chan-combined with the partner
//               warp) + ReLU, and feed the result back as the A operand of ctr2 (4 more stages, no trip through
//               memory); GroupNorm + residual + ReLU and the TMA stores of that second result run three stages into
//               the next tile.
//   TMEM        A ring 4 x (hi 32 | lo 32) = [0,256) | main [256,384) | cross [384,512)
//   smem        W ring 4 x 32 KB | X ring 3 x 16 KB | store staging 8 x 4 KB | norm vectors, stats, table slots
//
// Measured (batch 128: 193,536 rows, 2.4 M edges): 0.48 ms per block against 0.81 ms for the three split kernels;
// DESIGN.md section 3 lists what the tuning found and what was tried and rejected.
#include <cstring>
#include <mutex>

#include "tc_common.cuh"

using namespace tc;

int lgcn_launch_laneconv_v2(const float* x, const float* xa, const int32_t* tab, float* out, int64_t n_nodes,
                            const int32_t* n_dev, int n_keys, const float* w_hi, const float* w_lo, const float* gn,
                            cudaStream_t st);

namespace {

#ifndef LGCN_FUSED_WAIT
#define LGCN_FUSED_WAIT mbar_spin
#endif

constexpr int kTileM = 128;
constexpr int kWStages = 4;
constexpr int kXStages = 3;                                  // raw fp32 A chunks in flight (cp.async ring)
constexpr int kWStageBytes = 2 * 128 * 128;                  // hi 16 KB | lo 16 KB
constexpr int kAStages = 4;
constexpr int kSmemX = kWStages * kWStageBytes;              // 128 KB: X ring, kXStages x (8 warps x 32 rows x 64 B)
constexpr int kSmemOut = kSmemX + kXStages * 16384;          // 8 x 4 KB store staging
constexpr int kSmemGam = kSmemOut + 8 * 4096;                // gamma1 | beta1 | gamma2 | beta2 (4 x 512 B)
constexpr int kSmemStat = kSmemGam + 4 * 512;                // float2 [2][128]
constexpr int kSmemTab = kSmemStat + 2 * 128 * 8;            // int32 [4][256]: table entries in flight (cp.async)
constexpr int kSmemBar = kSmemTab + 4 * 256 * 4;
constexpr int kSmemHead = kSmemBar + 256;                    // 2 KB: float4 [128] = (w0, w1, b, 0) of a K=2 head (MLP0), or the
                                                             //       [128][4] weights of the 4 extra columns (KS4)
constexpr int kSmemTotal = kSmemHead + 2048;
constexpr int kNumThreads = 384;
constexpr int kMmaWarp = 1;
constexpr uint32_t kIdesc = idesc_tf32(128, 128);
// Cross terms on the 16-bit tensor path (2x the tf32 rate): D_cross += [a_lo | a] . [w | w_lo]^T over a K = 64 chunk per
// stage (4 MMAs of K = 16 instead of 8 tf32 MMAs; a stage = 512 tensor cycles instead of 768); main (tf32(a) . tf32(w)) is
// unchanged.  LGCN_BF16_CROSS selects the 16-bit format:
//   0  tf32 cross terms (the 3xTF32 form of round 1)
//   1  bf16: fp32 range, but 8-bit mantissas put 2^-20 |a w| per term into the sum — measured 1.4x the north-star
//      tolerance on the batch-32 forward (profiles/r2_cross_terms.md), so NOT used
//   2  fp16 with the lo operands scaled by 2^11 (a_lo 2^11 ~ |a|, w_lo 2^11 ~ |w|: well inside fp16's range; the cross
//      accumulator then holds 2^11 x the cross sum and the drain multiplies by 2^-11, exactly): 10-bit mantissas, i.e.
//      the SAME rounding as tf32 cross terms.  Operands beyond fp16's range saturate at +-65504 (cvt.satfinite: the cross
//      term of such an element degrades to plain-TF32 accuracy instead of producing inf).
// Measured on the batch-128 graph (tools/ablate_fused.py, us per block): 519 (mode 0), 484 (mode 1), 497 (mode 2): the
// kernel is bound by the producer / barrier hand-off (355 us with every MMA, load and conversion switched off), not by
// the tensor pipe, so a third fewer tensor cycles buy 4-7 %.  Mode 0 stays the default: identical arithmetic to the
// split path and no range caveat; modes 1 / 2 are kept for the day the A feed is faster (-DLGCN_BF16_CROSS=2).
#ifndef LGCN_BF16_CROSS
#define LGCN_BF16_CROSS 0
#endif
#if LGCN_BF16_CROSS == 2
constexpr uint32_t kIdescX = idesc_f16(128, 128);
constexpr float kLoScale = 2048.0f, kCrossUnscale = 1.0f / 2048.0f;
#define LGCN_PACK16(hi, lo) pack_f16(hi, lo)
#else
constexpr uint32_t kIdescX = idesc_bf16(128, 128);
constexpr float kLoScale = 1.0f, kCrossUnscale = 1.0f;
#define LGCN_PACK16(hi, lo) pack_bf16(hi, lo)
#endif
constexpr uint32_t kColMain = 256, kColCross = 384;
// The tensor core truncates the fp32 accumulator toward zero once per MMA instruction (-1.3e-8 relative each, see
// gemm_tc.cu), so a 240-instruction hi.hi chain (15 keys x 16 k-steps) would carry a 3e-6 bias.  The main accumulator
// is therefore flushed into registers (round-to-nearest fp32 adds) every kFlushKeys keys; the cross accumulator
// holds values 2^-11 smaller and runs through.
constexpr int kFlushKeys = 3;

// cvt.rna.tf32.f32 for finite inputs (round to nearest, ties away from zero, on the sign-magnitude bit pattern):
// two integer instructions instead of the multi-instruction sequence ptxas emits for the cvt (which also handles
// Inf / NaN; features never are)
__device__ __forceinline__ float rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

struct FusedArgs {
  const float* X;        // [M,128] input features (also the residual)
  const float* XA;       // auxiliary rows (multi-source sums)
  const int32_t* tab;    // [n_tiles][n_keys][128]
  const float* gn;       // gamma1 | beta1 | gamma2 | beta2
  int64_t M;             // row capacity (grid, tensor-map extent) ...
  const int32_t* m_dev;  // ... and, when not NULL, the live row count in device memory (min(*m_dev, M) rows are processed)
  int n_keys;
  int chain;             // 1: ctr2 + GN + residual + ReLU inside the kernel
  int flags2;            // linear mode with LCHAIN: LGCN_EPI_* of the second Linear (no GN)
  // linear mode (tab == nullptr): out = epilogue( sum_k W_k . src[k][ idx[k] ? idx[k][row] : row ] ), the generic
  // bias-free Linear (+GroupNorm, +ReLU, +residual, +ReLU) of lgcn_linear128 with up to three K=128 sources
  const float* src[3];
  const int32_t* idx[3];
  const float* res;
  const float* beta;     // linear mode: gn = gamma, beta separate
  const float* xs;       // KS4: [M,4] extra input columns ...
  const float* wx;       // ... and their weights: wx[n * ldw + j] = W[n, n_src*128 + j]
  int64_t ldw;
  int flags;             // LGCN_EPI_* (linear mode)
  // MLP0: source 0 is not read but COMPUTED per row: relu(W1 . (p[ip ? ip[m] : m] - q[iq ? iq[m] : m]) + b1), the
  // nn.Linear(2, 128) + ReLU heads in front of MapNet.input / seg and Att.dist (lanegcn.py:277-286, 644-648, 693)
  const float* head_w;   // W1 [128][2] | b1 [128]
  const float2* head_p;
  const int32_t* head_ip;
  const float2* head_q;  // may be NULL (no subtraction)
  const int32_t* head_iq;
  long long* tl;         // timeline buffer [1024][8] (dbg & 256, CTA 0 only)
  int dbg;               // lgcn_debug_flags (ablation: 1 no stores, 4 no MMAs, 8 no loads, 32 no A conversion, 64 no flushes, 128 no weight loads)
};

// LINEAR = false: one LaneConv block (plan / table driven);  LINEAR = true: the generic Linear of lgcn_linear128.
// Two instantiations so that neither hot loop carries the other mode's branches and registers.
// KS4 (linear mode only): rank-4 update of the accumulators by four extra input columns (A2M.meta, lanegcn.py:387-395).
// LCHAIN (linear mode only): a second Linear (weights = the block after the sources' blocks) is applied to the first one's
// result without leaving the kernel, like ctr2: y = epi2(W2 . epi1(sum_s W_s x_s))  (Att: ctx.0 -> ctx.1, lanegcn.py:698-700).
template <bool LINEAR, bool KS4 = false, bool LCHAIN = false, bool MLP0 = false>
__global__ void __launch_bounds__(kNumThreads, 1)
k_laneconv_fused(const FusedArgs a, const __grid_constant__ CUtensorMap out_map,
                 const __grid_constant__ CUtensorMap whi_map, const __grid_constant__ CUtensorMap wlo_map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  lgcn_pdl_trigger();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  if (sbase & 1023u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // One full / empty barrier pair per stage serves both rings (weights in shared memory, A in tensor memory): the MMA
  // warp pays one try_wait and one tcgen05.commit per stage instead of two (each costs the issuing thread ~200 cycles
  // during which the tensor pipe, whose queue is only a couple of instructions deep, runs dry).
  static_assert(kWStages == kAStages, "the weight ring and the A ring share their barriers");
  const uint32_t bar_full = sbase + kSmemBar;               // [4]: 8 producer-warp arrives + the TMA expect_tx arrive
  const uint32_t bar_empty = bar_full + 8 * kAStages;       // [4]: tcgen05.commit
  const uint32_t bar_acc_full = bar_empty + 8 * kAStages;
  const uint32_t bar_acc_empty = bar_acc_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + 192);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kAStages; ++i) {
      mbar_init(bar_full + 8 * i, 9);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_acc_empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    float* g = reinterpret_cast<float*>(smem + kSmemGam);
    if (!LINEAR) {
      for (int i = threadIdx.x; i < 4 * 128; i += kNumThreads) g[i] = a.gn[i];
    } else if (a.gn && threadIdx.x < 256) {
      g[threadIdx.x] = threadIdx.x < 128 ? a.gn[threadIdx.x] : a.beta[threadIdx.x - 128];
    }
  }
  if constexpr (MLP0) {   // (w0, w1, b, 0) per output channel: one broadcast 16-byte shared load per computed value
    if (threadIdx.x < 128) {
      const float2 w = __ldg(reinterpret_cast<const float2*>(a.head_w) + threadIdx.x);
      reinterpret_cast<float4*>(smem + kSmemHead)[threadIdx.x] = make_float4(w.x, w.y, __ldg(a.head_w + 256 + threadIdx.x), 0.f);
    }
  }
  if constexpr (KS4) {    // the [128][4] weights of the extra columns (64 strided global loads per row otherwise)
    if (threadIdx.x < 128)
      reinterpret_cast<float4*>(smem + kSmemHead)[threadIdx.x] =
          __ldg(reinterpret_cast<const float4*>(a.wx + (int64_t)threadIdx.x * a.ldw));
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  lgcn_pdl_wait();   // everything above touched only this CTA's shared / tensor memory and the (static) weights

  // scalars only below (a by-value struct captured by reference in a lambda ends up in local memory)
  const int64_t M = lgcn_devn(a.m_dev, a.M);
  const float* __restrict__ X = a.X;
  const float* __restrict__ XA = a.XA;
  const int32_t* __restrict__ tab = a.tab;
  constexpr bool linear = LINEAR;
  const float* __restrict__ src0 = a.src[0];
  const float* __restrict__ src1 = a.src[1];
  const float* __restrict__ src2 = a.src[2];
  const int32_t* __restrict__ idx0 = a.idx[0];
  const int32_t* __restrict__ idx1 = a.idx[1];
  const int32_t* __restrict__ idx2 = a.idx[2];
  const float* __restrict__ lin_res = a.res;
  const float* __restrict__ lin_xs = a.xs;
  const int lin_flags = a.flags;
  const int n_keys = a.n_keys, nk = a.n_keys + 1;
  const bool chain = LINEAR ? LCHAIN : a.chain != 0;
  const int lin_flags2 = a.flags2;
  const int dbg = a.dbg;
#ifdef LGCN_TIMELINE
  long long* tl = (blockIdx.x == 0 && (dbg & 256)) ? a.tl : nullptr;
#define LGCN_TL_MMA(c) if (tl && tls < 1024) tl[tls * 8 + (c)] = clock64()
#define LGCN_TL_PROD(c) if (tl && e == 0 && lane == 0 && tls < 1024) tl[tls * 8 + (c)] = clock64()
#else
#define LGCN_TL_MMA(c) (void)0
#define LGCN_TL_PROD(c) (void)0
#endif
#ifdef LGCN_TIMELINE2   // fine-grained producer stamps (warp 4, lane 0): 16 x 32-bit clocks per stage into a.tl viewed as int32 [1024][16]
  uint32_t tk[16];
#define LGCN_TK(i) asm volatile("mov.u32 %0, %%clock;" : "=r"(tk[i]) :: "memory")
#else
#define LGCN_TK(i) (void)0
#endif
  const int flush_keys = (dbg & 64) ? (1 << 20) : (dbg & 512) ? 5 : kFlushKeys;
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int64_t grid = gridDim.x;
  const int keys_per_tile = nk + (chain ? 1 : 0);   // key nk = ctr2 (weights follow the projections in w_hi / w_lo)

  if (warp == 0) {
    // =========================================================== TMA producer: weight chunks
    uint32_t phase = 0;
    int ws = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += grid) {
      for (int kk = 0; kk < keys_per_tile; ++kk) {
        for (int kc = 0; kc < 4; ++kc) {
          LGCN_FUSED_WAIT(bar_empty + 8 * ws, phase ^ 1);
          if (elect_one()) {
            const uint32_t bar = bar_full + 8 * ws, dst = sbase + ws * kWStageBytes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kWStageBytes) : "memory");
            if (dbg & 128) {   // ablation: no weight loads
              asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"((uint32_t)kWStageBytes) : "memory");
            } else {
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&whi_map)), "r"(bar), "r"(kc * 32), "r"(kk * 128)
                : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(dst + kWStageBytes / 2), "l"(reinterpret_cast<uint64_t>(&wlo_map)), "r"(bar),
                  "r"(kc * (LGCN_BF16_CROSS ? 64 : 32)), "r"(kk * 128)
                : "memory");
            }
          }
          __syncwarp();
          if (++ws == kWStages) {
            ws = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================================================== MMA issuer
    // The full barrier of stage s+1 is TESTED (non-blocking) before the MMAs of stage s are issued and the predicate is
    // read after them, so the ~200-cycle round trip of the test overlaps the issue; the blocking wait runs only when
    // the producers are not ahead (tools/timeline_fused.py).
    // ONE elected thread runs the whole loop (waits included): no per-stage elect / __syncwarp.
    if (elect_one()) {
      uint32_t phase = 0, acc_uses = 0;
      int st = 0, tls = 0;
      const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + grid - 1) / grid : 0;
      int64_t stages_left = my_tiles * keys_per_tile * 4;
      asm volatile(".reg .pred lgcn_peek;");
      if (stages_left > 0) {
        LGCN_FUSED_WAIT(bar_full, 0);
        tc_fence_after();
      }
      const uint32_t d_main = tmem_base + kColMain, d_cross = tmem_base + kColCross;
      for (int64_t t = blockIdx.x; t < n_tiles; t += grid) {
        int since = 0;   // keys accumulated in main since its last restart
        for (int kk = 0; kk < keys_per_tile; ++kk) {
          const bool fresh_all = kk == 0 || kk == nk;                     // projections / ctr2 start
          const bool fresh_main = fresh_all || since == flush_keys;       // main restarts after a flush
          // after a flush only MAIN restarts: the cross-term MMAs of the first stage are issued before waiting for the
          // producers to have read main, so the tensor pipe has 8 instructions of work during the flush
          const bool late_wait = fresh_main && !fresh_all && !(dbg & 4);
          if (fresh_main) {
            since = 0;
            if (!late_wait) {
              LGCN_FUSED_WAIT(bar_acc_empty, (acc_uses & 1) ^ 1);  // the producers have read the accumulators
              ++acc_uses;
              tc_fence_after();
            }
          }
          ++since;
          const bool publish = kk >= nk - 1 || since == flush_keys;       // a flush / drain follows this key
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            const uint32_t a_hi = tmem_base + st * 64, a_lo = a_hi + 32;
            const uint32_t w_hi = sbase + st * kWStageBytes, w_lo = w_hi + kWStageBytes / 2;
            int st2 = st + 1;
            uint32_t ph2 = phase;
            if (st2 == kAStages) {
              st2 = 0;
              ph2 ^= 1;
            }
            --stages_left;
            LGCN_TL_MMA(0);
            asm volatile("mbarrier.test_wait.parity.shared::cta.b64 lgcn_peek, [%0], %1;" ::"r"(bar_full + 8 * st2), "r"(ph2)
                         : "memory");
            if (kc == 0 && late_wait) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
#if LGCN_BF16_CROSS
                umma_bf16_ts(d_cross, a_lo + 8 * j, umma_desc(w_lo + j * 32), kIdescX, 1u);
#else
                umma_tf32_ts(d_cross, a_lo + 8 * j, umma_desc(w_hi + j * 32), kIdesc, 1u);
                umma_tf32_ts(d_cross, a_hi + 8 * j, umma_desc(w_lo + j * 32), kIdesc, 1u);
#endif
              }
              LGCN_FUSED_WAIT(bar_acc_empty, (acc_uses & 1) ^ 1);  // the producers have added main to their sums
              ++acc_uses;
              tc_fence_after();
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_tf32_ts(d_main, a_hi + 8 * j, umma_desc(w_hi + j * 32), kIdesc, j ? 1u : 0u);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (dbg & 4) break;
                const bool k0 = kc == 0 && j == 0;
#if LGCN_BF16_CROSS
                // a_lo / w_lo name the CROSS operands here: [a_lo | a] (32 TMEM columns of bf16 pairs), [w | w_lo] (bf16 tile)
                umma_bf16_ts(d_cross, a_lo + 8 * j, umma_desc(w_lo + j * 32), kIdescX, (fresh_all && k0) ? 0u : 1u);
#else
                umma_tf32_ts(d_cross, a_lo + 8 * j, umma_desc(w_hi + j * 32), kIdesc, (fresh_all && k0) ? 0u : 1u);
                umma_tf32_ts(d_cross, a_hi + 8 * j, umma_desc(w_lo + j * 32), kIdesc, 1u);
#endif
                umma_tf32_ts(d_main, a_hi + 8 * j, umma_desc(w_hi + j * 32), kIdesc, (fresh_main && k0) ? 0u : 1u);
              }
            }
            umma_commit(bar_empty + 8 * st);
            if (publish && kc == 3) umma_commit(bar_acc_full);
            LGCN_TL_MMA(1);
            uint32_t ready;
            asm volatile("selp.u32 %0, 1, 0, lgcn_peek;" : "=r"(ready));
            if (stages_left > 0) {
              if (!ready) LGCN_FUSED_WAIT(bar_full + 8 * st2, ph2);
              tc_fence_after();
            }
            LGCN_TL_MMA(2);
            ++tls;
            st = st2;
            phase = ph2;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // =========================================================== A producers + epilogue: warps 4..11
    const int e = warp - 4, q = e & 3, h = e >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t my_buf = sbase + kSmemOut + e * 4096;
    const uint32_t stat_mine = sbase + kSmemStat + (h * 128 + r) * 8, stat_other = sbase + kSmemStat + ((h ^ 1) * 128 + r) * 8;
    const uint32_t gam = sbase + kSmemGam + h * 256;   // + 0: gamma1, + 512: beta1, + 1024: gamma2, + 1536: beta2

    // ---- A feed.  Rows are gathered with cp.async (global -> shared, no registers, three stages in flight): four
    // lanes fetch the 64 B slice of one source row, so a warp-wide 16 B request touches 8 rows instead of 32 (with one
    // row per lane the L1 tag stage, one lookup per distinct line, was the limiter).  The owner lane of a row then
    // reads its 64 B back (XOR-swizzled 16 B pieces: conflict-free both ways), splits hi/lo and stores to TMEM.
    // Table entries (source row of (r, key)) travel the same way, two keys ahead.
    const int n_tiles_i = (int)n_tiles, grid_i = (int)grid;
    const uint32_t slot0 = sbase + kSmemTab + (e * 32 + lane) * 4;   // + 1024 * (key index & 3)
    const uint32_t xblk = sbase + kSmemX + e * 2048;                 // + 16384 * slot: this warp's 32 rows x 64 B
    auto key_norm = [&](int& t, int& kk) {
      while (kk >= nk) {
        kk -= nk;
        t += grid_i;
      }
    };
    auto self_src = [&](int t) -> int {
      const int64_t m = (int64_t)t * kTileM + r;
      return (t < n_tiles_i && m < M) ? (int)m : -1;
    };
    auto key_idx = [&](int kk) -> const int32_t* { return kk == 0 ? idx0 : kk == 1 ? idx1 : idx2; };
    auto request = [&](int t, int kk, int ring) {   // joins the cp.async group of the current stage
      key_norm(t, kk);
      if (t >= n_tiles_i) return;
      const int32_t* src = nullptr;
      if (linear) {
        const int32_t* ix = key_idx(kk);
        const int64_t m = (int64_t)t * kTileM + r;
        if (ix && m < M) src = ix + m;
      } else if (kk > 0) {
        src = tab + (((int64_t)t * n_keys + (kk - 1)) << 7) + r;
      }
      if (src) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(slot0 + 1024u * ring), "l"(src) : "memory");
    };
    // sources of the four rows this lane fetches for (tile t, key kk): rows 8i + (lane >> 2) of the warp's block
    auto sources = [&](int t, int kk, int ring, int (&vr)[4]) {
      key_norm(t, kk);
      int v = -1;
      if (t < n_tiles_i) {
        const bool own = linear ? key_idx(kk) == nullptr : kk == 0;
        v = self_src(t);
        if (!own && v >= 0) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(slot0 + 1024u * ring) : "memory");
        if (MLP0 && kk == 0) v = -1;   // computed source: the ring slot is only zero-filled
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) vr[i] = __shfl_sync(0xffffffffu, v, 8 * i + (lane >> 2));
    };
    const uint32_t piece = lane & 3;
    // kk: the key the rows belong to (selects the source matrix in linear mode; may be nk = key 0 of the next tile)
    auto issue = [&](const int (&vr)[4], int kc, int slot, int kk) {   // 4 x 16 B per lane; one commit group per stage
      if (kk >= nk) kk -= nk;
      const float* base = !linear ? X : kk == 0 ? src0 : kk == 1 ? src1 : src2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (dbg & 1024) break;   // ablation: no cp.async instructions at all
        const int v = vr[i];
        const int row = 8 * i + (lane >> 2);
        const float* p = v >= 0 ? base + (int64_t)v * LGCN_C : XA + (int64_t)(v < -1 ? -2 - v : 0) * LGCN_C;
        const uint32_t dst = xblk + slot * 16384 + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4);
        const uint32_t n = (v != -1 && !(dbg & 8)) ? 16u : 0u;   // 0 source bytes: the 16 B are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(p + kc * 32 + h * 16 + piece * 4), "r"(n)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto take = [&](float4(&cur)[4], int slot) {   // this lane's row: 64 B of the slot (stage s; s+1, s+2 may be pending)
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      LGCN_TK(2);
      __syncwarp();
      const uint32_t src = xblk + slot * 16384 + lane * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) cur[c] = ld_shared_f4(src + ((c ^ ((lane >> 1) & 3)) << 4));
      __syncwarp();   // every lane has read: the slot may be refilled
      LGCN_TK(3);
    };
    uint32_t a_phase = 0, acc_uses = 0;
    int as = 0;
    // 16 floats -> hi/lo -> TMEM columns [c0, c0+16) (hi) and [c0+32, c0+48) (lo) of A stage `as`; eight columns per
    // tcgen05.st keeps the live temporaries at 16 registers
    auto put16 = [&](const float4(&x)[4], int c0) {
#if LGCN_BF16_CROSS
      // hi = tf32(x) -> columns [c0, c0+16); cross operand (two bf16 per column, k even in the low half):
      // x - hi -> columns 32 + [c0/2, c0/2 + 8),  x itself -> columns 48 + [c0/2, c0/2 + 8)
      uint32_t lp[8], xp[8];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t hi[8];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 y = x[2 * g + c];
          const float h0 = rna(y.x), h1 = rna(y.y), h2 = rna(y.z), h3 = rna(y.w);
          hi[4 * c] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1);
          hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
          lp[4 * g + 2 * c] = LGCN_PACK16((y.y - h1) * kLoScale, (y.x - h0) * kLoScale);
          lp[4 * g + 2 * c + 1] = LGCN_PACK16((y.w - h3) * kLoScale, (y.z - h2) * kLoScale);
          xp[4 * g + 2 * c] = LGCN_PACK16(y.y, y.x);
          xp[4 * g + 2 * c + 1] = LGCN_PACK16(y.w, y.z);
        }
        TMEM_ST8(t_lane + as * 64 + c0 + 8 * g, hi, 0);
      }
      TMEM_ST8(t_lane + as * 64 + 32 + (c0 >> 1), lp, 0);
      TMEM_ST8(t_lane + as * 64 + 48 + (c0 >> 1), xp, 0);
#else
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 y = x[2 * g + c];
          const float h0 = rna(y.x), h1 = rna(y.y), h2 = rna(y.z), h3 = rna(y.w);
          hi[4 * c] = __float_as_uint(h0); lo[4 * c] = __float_as_uint(rna(y.x - h0));
          hi[4 * c + 1] = __float_as_uint(h1); lo[4 * c + 1] = __float_as_uint(rna(y.y - h1));
          hi[4 * c + 2] = __float_as_uint(h2); lo[4 * c + 2] = __float_as_uint(rna(y.z - h2));
          hi[4 * c + 3] = __float_as_uint(h3); lo[4 * c + 3] = __float_as_uint(rna(y.w - h3));
        }
        TMEM_ST8(t_lane + as * 64 + c0 + 8 * g, hi, 0);
        TMEM_ST8(t_lane + as * 64 + 32 + c0 + 8 * g, lo, 0);
      }
#endif
    };
    auto stage_begin = [&]() {
      LGCN_FUSED_WAIT(bar_empty + 8 * as, a_phase ^ 1);
      tc_fence_after();
    };
    // the same wait split in two: the (non-blocking) test is issued before the chunk is read from shared memory and
    // its predicate consumed afterwards, so the barrier round trip overlaps the shared-memory latency
    asm volatile(".reg .pred lgcn_peek_a;");
    auto stage_peek = [&]() {
      asm volatile("mbarrier.test_wait.parity.shared::cta.b64 lgcn_peek_a, [%0], %1;" ::"r"(bar_empty + 8 * as), "r"(a_phase ^ 1)
                   : "memory");
    };
    auto stage_begin_peeked = [&]() {
      uint32_t ready;
      asm volatile("selp.u32 %0, 1, 0, lgcn_peek_a;" : "=r"(ready));
      if (!ready) LGCN_FUSED_WAIT(bar_empty + 8 * as, a_phase ^ 1);
      tc_fence_after();
    };
    auto stage_end = [&]() {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      LGCN_TK(7);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * as);
      LGCN_TK(8);
      if (++as == kAStages) {
        as = 0;
        a_phase ^= 1;
      }
    };
    // f[64] = columns [64h, 64h+64) of this thread's row: running fp32 sum of the flushed main accumulator
    float f[64];
    auto acc_wait = [&]() {
      LGCN_FUSED_WAIT(bar_acc_full, acc_uses & 1);
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
      for (int g = 0; g < 4; ++g) {
        uint32_t v[16];
        TMEM_LD16(v, 0, t_lane + kColMain + h * 64 + g * 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) f[g * 16 + c] += __uint_as_float(v[c]);
      }
      acc_release();
    };
    auto drain = [&](bool add) {   // f (+)= main + cross, then the accumulators are released to the MMA warp
      acc_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t v[16], x[16];
        TMEM_LD16(v, 0, t_lane + kColMain + h * 64 + g * 16);
        TMEM_LD16(x, 0, t_lane + kColCross + h * 64 + g * 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float y = LGCN_BF16_CROSS == 2 ? fmaf(__uint_as_float(x[c]), kCrossUnscale, __uint_as_float(v[c]))
                                               : __uint_as_float(v[c]) + __uint_as_float(x[c]);
          f[g * 16 + c] = add ? f[g * 16 + c] + y : y;
        }
      }
      acc_release();
    };
    auto gn = [&](uint32_t gb) {   // GroupNorm(1) of the row; gb: shared address of gamma (beta 512 B behind it)
      // four independent partial sums: a single 64-long dependent chain costs ~4 cycles per element
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
      named_bar_sync(1 + q, 64);   // both halves have read: the slots may be rewritten by the next drain
      const float mean = 0.5f * (mean_h + o.x);
      const float dm = mean_h - o.x;
      const float var = (m2_h + o.y + dm * dm * 32.0f) * (1.0f / 128.0f);
      const float rstd = 1.0f / sqrtf(var + LGCN_GN_EPS);
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float4 g = ld_shared_f4(gb + c * 4), b = ld_shared_f4(gb + 512 + c * 4);
        f[c] = fmaf((f[c] - mean) * rstd, g.x, b.x);
        f[c + 1] = fmaf((f[c + 1] - mean) * rstd, g.y, b.y);
        f[c + 2] = fmaf((f[c + 2] - mean) * rstd, g.z, b.z);
        f[c + 3] = fmaf((f[c + 3] - mean) * rstd, g.w, b.w);
      }
    };
    auto store_out = [&](int64_t m0) {   // f -> two 32 x 32 boxes through the 4 KB staging buffer
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        if (dbg & 1) break;
        if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_shared_f4(my_buf + lane * 128 + ((c ^ (lane & 7)) << 4),
                       make_float4(f[cb * 32 + 4 * c], f[cb * 32 + 4 * c + 1], f[cb * 32 + 4 * c + 2], f[cb * 32 + 4 * c + 3]));
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&out_map)),
                       "r"(my_buf), "r"(h * 64 + cb * 32), "r"((int32_t)(m0 + q * 32))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    };

    // prologue: table entries of keys 1 and 2, chunks of stages 0..2
    int vr[4], vrn[4];
    int kseq = 0;   // running key count & 3: ring slot of a table entry = (kseq + distance) & 3
    request((int)blockIdx.x, 0, 0);   // (key 0 has an entry only in linear mode with a gathered first source)
    request((int)blockIdx.x, 1, 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    sources((int)blockIdx.x, 0, 0, vr);
#pragma unroll
    for (int s0 = 0; s0 < kXStages; ++s0) issue(vr, s0, s0, 0);   // stages 0..2 = chunks 0..2 of key 0 (kXStages <= 4)
    int xs = 0;   // ring slot of the stage being converted
    int tls = 0;
    // Second epilogue of a tile (ctr2 accumulators -> GroupNorm + residual + ReLU -> store).  It runs three stages
    // into the NEXT tile: those stages are produced while the ctr2 MMAs still execute, the accumulators are drained
    // the moment ctr2 retires, and the tensor pipe then has three stages of queued work while the norm, the residual
    // and the stores are done.  (Finishing the tile first idled the pipe ~7 k cycles per tile; spreading the epilogue
    // over four later stages was slower still: the producers are not far enough ahead; direct 16-byte global stores
    // instead of the staged TMA store took 2.0 k cycles instead of 1.6 k.  tools/timeline_fused.py.)
    auto finish_tile = [&](int64_t pm0) {
      const int64_t m = pm0 + r;
      const bool live = m < M;
      const float4* resp = reinterpret_cast<const float4*>(X + (live ? m : 0) * LGCN_C + h * 64);
      float4 ra[8], rb[8];   // residual: first half requested before the drain, second half before the norm
#pragma unroll
      for (int c = 0; c < 8; ++c) ra[c] = live ? __ldg(resp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      drain(false);
#pragma unroll
      for (int c = 0; c < 8; ++c) rb[c] = live ? __ldg(resp + 8 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      gn(gam + 1024);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        f[4 * c] = fmaxf(f[4 * c] + ra[c].x, 0.f);
        f[4 * c + 1] = fmaxf(f[4 * c + 1] + ra[c].y, 0.f);
        f[4 * c + 2] = fmaxf(f[4 * c + 2] + ra[c].z, 0.f);
        f[4 * c + 3] = fmaxf(f[4 * c + 3] + ra[c].w, 0.f);
        f[32 + 4 * c] = fmaxf(f[32 + 4 * c] + rb[c].x, 0.f);
        f[33 + 4 * c] = fmaxf(f[33 + 4 * c] + rb[c].y, 0.f);
        f[34 + 4 * c] = fmaxf(f[34 + 4 * c] + rb[c].z, 0.f);
        f[35 + 4 * c] = fmaxf(f[35 + 4 * c] + rb[c].w, 0.f);
      }
      store_out(pm0);
    };
    // linear mode: the generic Linear epilogue [GroupNorm] [ReLU] [+ residual] [ReLU], deferred like finish_tile
    auto finish_linear = [&](int64_t pm0) {
      const int64_t m = pm0 + r;
      const int fl = LCHAIN ? lin_flags2 : lin_flags;   // with a chain this is the SECOND Linear's epilogue
      const bool live = m < M && (fl & LGCN_EPI_RES);
      const float4* resp = reinterpret_cast<const float4*>(lin_res + (live ? m : 0) * LGCN_C + h * 64);
      float4 ra[8], rb[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) ra[c] = live ? __ldg(resp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      drain(false);
#pragma unroll
      for (int c = 0; c < 8; ++c) rb[c] = live ? __ldg(resp + 8 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (KS4) {   // the weights of the 4 extra columns are broadcast loads from L2
        if (m < M) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(lin_xs + m * 4));
          const uint32_t wr = sbase + kSmemHead + (h * 64) * 16;
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const float4 w = ld_shared_f4(wr + c * 16);
            f[c] = fmaf(x.w, w.w, fmaf(x.z, w.z, fmaf(x.y, w.y, fmaf(x.x, w.x, f[c]))));
          }
        }
      }
      if (!LCHAIN && (fl & LGCN_EPI_GN)) gn(gam);
      if (fl & LGCN_EPI_RELU1) {
#pragma unroll
        for (int c = 0; c < 64; ++c) f[c] = fmaxf(f[c], 0.f);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        f[4 * c] += ra[c].x;
        f[4 * c + 1] += ra[c].y;
        f[4 * c + 2] += ra[c].z;
        f[4 * c + 3] += ra[c].w;
        f[32 + 4 * c] += rb[c].x;
        f[33 + 4 * c] += rb[c].y;
        f[34 + 4 * c] += rb[c].z;
        f[35 + 4 * c] += rb[c].w;
      }
      if (fl & LGCN_EPI_RELU2) {
#pragma unroll
        for (int c = 0; c < 64; ++c) f[c] = fmaxf(f[c], 0.f);
      }
      store_out(pm0);
    };
    // MLP0: this row's input of the K=2 head, read a tile ahead (two dependent loads: index, then centre)
    const float2* __restrict__ head_p = a.head_p;
    const float2* __restrict__ head_q = a.head_q;
    const int32_t* __restrict__ head_ip = a.head_ip;
    const int32_t* __restrict__ head_iq = a.head_iq;
    auto head_xy = [&](int64_t t) -> float2 {
      const int64_t m = t * kTileM + r;
      float2 x = make_float2(0.f, 0.f);
      if (MLP0 && t < n_tiles && m < M) {
        x = __ldg(head_p + (head_ip ? __ldg(head_ip + m) : m));
        if (head_q) {
          const float2 y = __ldg(head_q + (head_iq ? __ldg(head_iq + m) : m));
          x.x -= y.x;
          x.y -= y.y;
        }
      }
      return x;
    };
    float2 xy_next = head_xy(blockIdx.x);
    bool pending = false;
    int64_t pending_m0 = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += grid) {
      const int64_t m0 = t * kTileM;
      const float2 xy = xy_next;
      const bool head_live = m0 + r < M;
      if (MLP0) xy_next = head_xy(t + grid);
      if (!pending) {
#pragma unroll
        for (int c = 0; c < 64; ++c) f[c] = 0.f;
      }
      int since = 0;   // keys accumulated in the main accumulator since its last restart
      for (int kk = 0; kk < nk; ++kk) {
        const bool flush_here = kk > 0 && since == flush_keys;
        if (flush_here) since = 0;
        ++since;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          if (kc == 3 && kk == 0 && pending) {   // before stage 3 of the new tile (before stage 4 measured slower)
            if (linear) finish_linear(pending_m0);
            else finish_tile(pending_m0);
            pending = false;
#pragma unroll
            for (int c = 0; c < 64; ++c) f[c] = 0.f;
          }
          float4 cur[4];
          LGCN_TL_PROD(3);
          LGCN_TK(0);
          stage_peek();
          LGCN_TK(1);
          take(cur, xs);
          if (MLP0 && kk == 0) {   // source 0 = relu(W1 . xy + b1): same association as addmm(bias, x, W^T) in fp32 (k_mlp2_in)
            const uint32_t wq = sbase + kSmemHead + (32 * kc + 16 * h) * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 w0 = ld_shared_f4(wq + (4 * c) * 16), w1 = ld_shared_f4(wq + (4 * c + 1) * 16),
                           w2 = ld_shared_f4(wq + (4 * c + 2) * 16), w3 = ld_shared_f4(wq + (4 * c + 3) * 16);
              cur[c] = make_float4(fmaxf(fmaf(xy.y, w0.y, xy.x * w0.x) + w0.z, 0.f), fmaxf(fmaf(xy.y, w1.y, xy.x * w1.x) + w1.z, 0.f),
                                   fmaxf(fmaf(xy.y, w2.y, xy.x * w2.x) + w2.z, 0.f), fmaxf(fmaf(xy.y, w3.y, xy.x * w3.x) + w3.z, 0.f));
              if (!head_live) cur[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          LGCN_TL_PROD(4);
          stage_begin_peeked();
          LGCN_TK(4);
          LGCN_TL_PROD(5);
          if (!(dbg & 32)) put16(cur, h * 16);
          LGCN_TK(5);
          // refill the slot with stage + 3 (chunk kc-1 of the next key) while the tcgen05.st complete; at kc == 0 the
          // sources of the next key come first, and that longer sequence runs after the publish instead
          if (kc != 0) issue(vrn, kc - 1, xs, kk + 1);
          LGCN_TK(6);
          stage_end();
          LGCN_TL_PROD(6);
          if (kc == 0) {
            // all groups but the two newest have landed: the entry of key kk+1 (requested a key ago) is readable
            sources((int)t, kk + 1, (kseq + 1) & 3, vrn);
            request((int)t, kk + 2, (kseq + 2) & 3);   // joins this stage's group
            issue(vr, 3, xs, kk);                      // chunk 3 of this key
          }
          if (++xs == kXStages) xs = 0;
          // flush of the key group that ended at kk-1: three stages late, so the A ring is full again when the MMA
          // warp resumes.  (Handling two stages per iteration -- one cp.async wait, one tcgen05.wait::st and two
          // independent conversion chains per pair -- was slower, 548 vs 507 us per block: the coarser hand-off costs
          // more pipelining than the shared fixed costs save.  Two producer TEAMS -- warps 4-7 / 8-11 producing the
          // stages of one parity each, whole 32-float chunks, 5 arrivals per stage -- was slower too, 560 us, and so
          // was its barrier skeleton: the per-warp instruction stream is not what bounds the feed.)
          LGCN_TK(9);
          if (kc == 2 && flush_here) flush_main();
          LGCN_TK(10);
#ifdef LGCN_TIMELINE2
          if (a.tl && (dbg & 256) && blockIdx.x == 0 && e == 0 && lane == 0 && tls < 1024) {
            uint32_t* t2 = reinterpret_cast<uint32_t*>(a.tl) + tls * 16;
#pragma unroll
            for (int i = 0; i < 11; ++i) t2[i] = tk[i];
            t2[11] = (uint32_t)(kk * 4 + kc);
          }
#endif
          ++tls;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) vr[i] = vrn[i];
        kseq = (kseq + 1) & 3;
      }
      if (linear && !LCHAIN) {   // the epilogue runs three stages into the next tile (finish_linear below)
        pending = true;
        pending_m0 = m0;
        continue;
      }
      drain(true);
      if (!linear || (lin_flags & LGCN_EPI_GN)) gn(gam);
      if (!linear || (lin_flags & LGCN_EPI_RELU1)) {
#pragma unroll
        for (int c = 0; c < 64; ++c) f[c] = fmaxf(f[c], 0.f);
      }
      if (!chain) {
        store_out(m0);
        continue;
      }
      // ---- ctr2: h = f is the A operand; K-chunk kc = columns [32 kc, 32 kc + 32) lives in the warps with h == kc>>1
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        stage_begin();
        if ((kc >> 1) == h) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float4 x[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int o = (kc & 1) * 32 + g * 16 + 4 * c;
              x[c] = make_float4(f[o], f[o + 1], f[o + 2], f[o + 3]);
            }
            put16(x, g * 16);
          }
        }
        stage_end();
      }
      pending = true;
      pending_m0 = m0;
    }
    if (pending) {
      if (linear) finish_linear(pending_m0);
      else finish_tile(pending_m0);
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");   // prefetches past the last tile (zero-filled)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ------------------------------------------------------------------ plan: one source row per (row, key)
struct PlanHdr {
  int32_t n_multi, n_mcol, pad[62];
};

// One thread per destination row, two passes over the row's CSR entries (sorted by key; col = v * nb + k + 1):
//   1. every key's run -> the table entry of the single-source keys, and the row's number of multi-source keys / of their
//      sources;  a block-wide scan turns those into offsets, ONE pair of atomics per CTA reserves the CTA's descriptors
//      (the first version paid two same-address atomics per multi-source key — 160 k at batch 128 — and walked the
//      row through a data-dependent while loop: 70 us stand-alone);
//   2. rows that have multi-source keys write their descriptors, source lists and table entries.
// The descriptor order depends on which CTA reserves first; what a table entry points at does not.
constexpr int kPlanThreads = 512;
__global__ void __launch_bounds__(kPlanThreads)
k_plan_build(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n_keys,
             int64_t n_cap, const int32_t* __restrict__ n_dev, int32_t* __restrict__ hdr,
             int32_t* __restrict__ tab, int2* __restrict__ mdesc, int32_t* __restrict__ mcol,
             int64_t max_multi) {
  __shared__ int32_t wsum[2][kPlanThreads / 32];
  __shared__ int32_t base[2];
  const int64_t n_nodes = lgcn_devn(n_dev, n_cap);
  const int64_t n_rows_padded = (n_nodes + kTileM - 1) / kTileM * kTileM;
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = m < n_rows_padded;
  int32_t* my = tab + ((m >> 7) * n_keys << 7) + (m & 127);
  const uint32_t nb = n_keys + 1;
  const uint64_t magic = ((1ull << 40) + nb - 1) / nb;   // c / nb == (c * magic) >> 40 for c < 2^31, nb <= 17
  int32_t beg = 0, end = 0;
  if (m < n_nodes) {
    beg = rowptr[m];
    end = rowptr[m + 1];
  }
  int32_t n_multi = 0, n_src = 0;
  if (in_range) {
    for (int k = 0; k < n_keys; ++k) my[(int64_t)k << 7] = -1;
    uint32_t run_key = 0, run_first = 0;
    int32_t run_len = 0;
    auto close_run = [&]() {
      if (run_len == 1) my[(int64_t)(run_key - 1) << 7] = (int32_t)run_first;
      else if (run_len > 1) { ++n_multi; n_src += run_len; }
    };
#pragma unroll 4
    for (int32_t j = beg; j < end; ++j) {
      const uint32_t c = (uint32_t)col[j];
      const uint32_t src = (uint32_t)((c * magic) >> 40), key = c - src * nb;
      if (key != run_key) {
        close_run();
        run_key = key; run_first = src; run_len = 0;
      }
      ++run_len;
    }
    close_run();
  }
  // block-wide exclusive scan of (n_multi, n_src)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t inc[2] = {n_multi, n_src};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, inc[q], o);
      if (lane >= o) inc[q] += t;
    }
    if (lane == 31) wsum[q][warp] = inc[q];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int32_t v = lane < kPlanThreads / 32 ? wsum[q][lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (lane < kPlanThreads / 32) wsum[q][lane] = v;                     // inclusive over warps
      if (lane == kPlanThreads / 32 - 1) base[q] = v ? atomicAdd(hdr + q, v) : 0;   // hdr[0]: descriptors, hdr[1]: sources
    }
  }
  __syncthreads();
  if (!n_multi) return;
  int32_t di = base[0] + (warp ? wsum[0][warp - 1] : 0) + inc[0] - n_multi;   // this row's first descriptor ...
  int32_t si = base[1] + (warp ? wsum[1][warp - 1] : 0) + inc[1] - n_src;     // ... and first source slot
  int32_t j = beg;
  while (j < end) {
    const uint32_t c = (uint32_t)col[j];
    const uint32_t key = c - (uint32_t)((c * magic) >> 40) * nb;
    int32_t j1 = j + 1;
    while (j1 < end) {
      const uint32_t c1 = (uint32_t)col[j1];
      if (c1 - (uint32_t)((c1 * magic) >> 40) * nb != key) break;
      ++j1;
    }
    const int32_t cnt = j1 - j;
    if (cnt > 1) {
      if (di < max_multi) {
        mdesc[di] = make_int2(si, cnt);
        for (int32_t q = 0; q < cnt; ++q) mcol[si + q] = (int32_t)(((uint32_t)col[j + q] * magic) >> 40);
        my[(int64_t)(key - 1) << 7] = -2 - di;
      }
      ++di;
      si += cnt;
    }
    j = j1;
  }
}

// XA[i] = sum of the source rows of multi-source entry i, in CSR (= edge-list) order; one warp per entry
__global__ void __launch_bounds__(256)
k_multi_sum(const float* __restrict__ X, const int32_t* __restrict__ hdr, const int2* __restrict__ mdesc,
            const int32_t* __restrict__ mcol, float* __restrict__ XA, int64_t max_multi) {
  const int lane = threadIdx.x & 31;
  lgcn_pdl_trigger();
  int64_t n = hdr[0];
  if (n > max_multi) n = max_multi;
  for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += (int64_t)gridDim.x * 8) {
    const int2 d = mdesc[i];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int32_t j0 = 0; j0 < d.y; j0 += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u < d.y) v[u] = __ldg(reinterpret_cast<const float4*>(X + (int64_t)mcol[d.x + j0 + u] * LGCN_C) + lane);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u < d.y) {
          acc.x += v[u].x;
          acc.y += v[u].y;
          acc.z += v[u].z;
          acc.w += v[u].w;
        }
    }
    reinterpret_cast<float4*>(XA + i * LGCN_C)[lane] = acc;
  }
}

long long* g_timeline = nullptr;

struct PlanView {
  int32_t* hdr;
  int32_t* tab;
  int2* mdesc;
  int32_t* mcol;
  int64_t max_multi;
};
PlanView plan_view(void* plan, int64_t n_nodes, int64_t n_edges, int n_keys) {
  PlanView v;
  const int64_t n_tiles = (n_nodes + kTileM - 1) / kTileM;
  char* p = (char*)plan;
  v.hdr = (int32_t*)p;
  p += 256;
  v.tab = (int32_t*)p;
  p += lgcn_align_up(n_tiles * n_keys * 128 * 4, 256);
  v.max_multi = n_edges / 2;
  v.mdesc = (int2*)p;
  p += lgcn_align_up((v.max_multi + 1) * 8, 256);
  v.mcol = (int32_t*)p;
  return v;
}

}  // namespace

extern "C" int64_t lgcn_laneconv_plan_bytes(int64_t n_nodes, int64_t n_edges, int n_keys) {
  const int64_t n_tiles = (n_nodes + kTileM - 1) / kTileM;
  return 256 + lgcn_align_up(n_tiles * n_keys * 128 * 4, 256) + lgcn_align_up((n_edges / 2 + 1) * 8, 256) +
         lgcn_align_up(n_edges * 4, 256) + 256;
}

// n_nodes / n_edges are capacities when n_dev (live node count in device memory) is given
int lgcn_launch_plan_build(const int32_t* rowptr, const int32_t* col, int n_keys, int64_t n_nodes, const int32_t* n_dev,
                           int64_t n_edges, void* plan, cudaStream_t st) {
  LGCN_CHECK_ARG(n_keys >= 0 && n_keys <= LGCN_MAX_KEYS, "plan_build: n_keys %d", n_keys);
  LGCN_CHECK_ARG(plan && (n_nodes == 0 || rowptr), "plan_build: NULL argument");
  LGCN_CHECK_ARG(n_nodes >= 0 && n_edges >= 0 && n_nodes < (1ll << 31) / 16, "plan_build: sizes");
  PlanView v = plan_view(plan, n_nodes, n_edges, n_keys);
  if (int rc = lgcn_zero_async(v.hdr, 256, st)) return rc;
  const int64_t rows = (n_nodes + kTileM - 1) / kTileM * kTileM;
  if (rows == 0 || n_keys == 0) return 0;
  k_plan_build<<<lgcn_cdiv(rows, kPlanThreads), kPlanThreads, 0, st>>>(rowptr, col, n_keys, n_nodes, n_dev, v.hdr, v.tab, v.mdesc, v.mcol,
                                                    v.max_multi);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_laneconv_plan_build(const int32_t* rowptr, const int32_t* col, int n_keys, int64_t n_nodes,
                                        int64_t n_edges, void* plan, void* stream) {
  return lgcn_launch_plan_build(rowptr, col, n_keys, n_nodes, nullptr, n_edges, plan, (cudaStream_t)stream);
}

// profiling aid (tools/timeline_fused.py): device buffer [1024][8] of clock64 stamps written by CTA 0 when debug flag
// 256 is set.  Columns: MMA warp {weights ready, A ready, issued}, producer warp 4 {stage start, chunk in registers,
// A slot free, stored + published}.
extern "C" int lgcn_debug_timeline(long long* device_buffer) {
  g_timeline = device_buffer;
  return 0;
}
long long* lgcn_timeline_buffer() { return g_timeline; }

// ------------------------------------------------------------------ linear mode: lgcn_linear128 on this kernel
namespace {
// W [128, n_src*128] (row stride ldw) -> hi / lo blocks [n_src*128, 128]: block k, row n = W[n, 128k .. 128k+127]
// Weight images of the kernel: hi = tf32(w) as fp32 [row][128], and the CROSS image x (bf16, [row][4 chunks][64]):
// per 32-float K-chunk kc of a row, 32 x bf16(w) followed by 32 x bf16(w - tf32(w)) — the B operand of the K = 64 bf16
// cross MMA chain.  (With LGCN_BF16_CROSS=0: x is the fp32 lo = tf32(w - hi) image, [row][128].)  Both images of a row
// have the same byte size (512 B), so `lo` buffers keep their sizes.
__device__ __forceinline__ void split_store4(float4 x, int row, int c4, float* __restrict__ hi, float* __restrict__ lo) {
  float4 a;
  a.x = rna(x.x); a.y = rna(x.y); a.z = rna(x.z); a.w = rna(x.w);
  reinterpret_cast<float4*>(hi)[(int64_t)row * 32 + c4] = a;
#if LGCN_BF16_CROSS
  const int kc = c4 >> 3, j4 = c4 & 7;   // chunk of 32 floats, float4 inside the chunk
  uint2* xr = reinterpret_cast<uint2*>(lo) + (int64_t)row * 64 + kc * 16;   // 16 uint2 (= 64 bf16) per chunk
  xr[j4] = make_uint2(LGCN_PACK16(x.y, x.x), LGCN_PACK16(x.w, x.z));
  xr[8 + j4] = make_uint2(LGCN_PACK16((x.y - a.y) * kLoScale, (x.x - a.x) * kLoScale),
                          LGCN_PACK16((x.w - a.w) * kLoScale, (x.z - a.z) * kLoScale));
#else
  float4 b;
  b.x = rna(x.x - a.x); b.y = rna(x.y - a.y); b.z = rna(x.z - a.z); b.w = rna(x.w - a.w);
  reinterpret_cast<float4*>(lo)[(int64_t)row * 32 + c4] = b;
#endif
}

// W [128, n_src*128] (row stride ldw) -> images of blocks [n_src*128, 128]: block k, row n = W[n, 128k .. 128k+127]
__global__ void k_split_blocks(const float* __restrict__ w, int64_t ldw, int n_src, float* __restrict__ hi,
                               float* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of the output
  if (i >= n_src * 128 * 32) return;
  const int c4 = i & 31, n = (i >> 5) & 127, k = i >> 12;
  split_store4(*reinterpret_cast<const float4*>(w + (int64_t)n * ldw + k * 128 + c4 * 4), i >> 5, c4, hi, lo);
}

__global__ void k_split_many(const LgcnSplitList l, float* __restrict__ hi, float* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of the output
  if (i >= l.n_blocks * 128 * 32) return;
  const int c4 = i & 31, n = (i >> 5) & 127, b = i >> 12;
  split_store4(*reinterpret_cast<const float4*>(l.p[b] + (int64_t)n * l.ldw[b] + c4 * 4), i >> 5, c4, hi, lo);
}

// a flat run of 128-float rows (a whole LaneConv wpack: the norm vectors between the matrices are split harmlessly)
__global__ void k_split_rows(const float4* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) split_store4(w[i], (int)(i >> 5), (int)(i & 31), hi, lo);
}

// Scratch for the split weights of a launch whose caller did not pre-split them (the stand-alone lgcn_linear128 entry
// point; lgcn_forward / lgcn_att_forward / the LaneConv stacks pass pre-split weights and never come here): a ring of
// device slots PER DEVICE, each guarded by an event recorded after the kernel that read it, so a slot is never
// rewritten (on any stream) before its last reader has finished.  The cursor is taken under a mutex (ctypes releases
// the GIL, so two host threads may call concurrently).
constexpr int kRing = 16;
constexpr int64_t kSlotFloats = 2 * 3 * 128 * 128;
struct SplitRing {
  float* buf = nullptr;
  cudaEvent_t ev[kRing];
  bool used[kRing];
  int next = 0;
};
SplitRing g_rings[kMaxDevices];
std::mutex g_ring_mu;

int make_cross_map(CUtensorMap* m, const float* w_x, int64_t rows) {
#if LGCN_BF16_CROSS
  return make_map_2d_bf16(m, w_x, 256, rows, 256, 64, 128);    // [rows][4 chunks x 64 bf16], one chunk per box
#else
  return make_map_2d(m, w_x, LGCN_C, rows, LGCN_C, 32, 128);
#endif
}

int set_fused_attrs() {
  if (!first_use(kFamFused)) return 0;
  LGCN_CUDA_OK(cudaFuncSetAttribute(k_laneconv_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  LGCN_CUDA_OK(cudaFuncSetAttribute(k_laneconv_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  LGCN_CUDA_OK(cudaFuncSetAttribute(k_laneconv_fused<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  LGCN_CUDA_OK(cudaFuncSetAttribute(k_laneconv_fused<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  LGCN_CUDA_OK(cudaFuncSetAttribute(k_laneconv_fused<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  return 0;
}
}  // namespace

// images of n floats (a multiple of 128) for the aggregate-first kernel (LaneConv stacks)
int lgcn_split_fused(const float* w, float* hi, float* lo, int64_t n, cudaStream_t st) {
  LGCN_CHECK_ARG(n % 128 == 0, "split_fused: n %% 128 != 0");
  if (n == 0) return 0;
  k_split_rows<<<lgcn_cdiv(n / 4, 256), 256, 0, st>>>((const float4*)w, hi, lo, n / 4);
  LGCN_LAUNCH_OK();
  return 0;
}

int64_t lgcn_linear_split_bytes() { return kSlotFloats * (int64_t)sizeof(float); }

int lgcn_split_blocks_many(const LgcnSplitList& l, float* hi, float* lo, cudaStream_t st) {
  LGCN_CHECK_ARG(l.n_blocks >= 1 && l.n_blocks <= 8, "split_blocks_many: n_blocks %d", l.n_blocks);
  k_split_many<<<lgcn_cdiv(l.n_blocks * 128 * 32, 256), 256, 0, st>>>(l, hi, lo);
  LGCN_LAUNCH_OK();
  return 0;
}

int lgcn_launch_linear_fused(const LinearArgs& la, cudaStream_t st) {
  if (la.m <= 0) return 0;
  LGCN_CHECK_ARG(la.n_out_blocks == 1 && (la.ks == 0 || la.ks == 4) && la.n_src >= 1 && la.n_src <= 3,
                 "linear_fused: unsupported shape");
  LGCN_CHECK_ARG(!la.head_w || (la.ks == 0 && !la.chain && la.w_hi && la.w_lo && la.head_p),
                 "linear_fused: a computed first source needs pre-split weights and excludes ks / chain");
  LGCN_CHECK_ARG(!la.chain || (la.ks == 0 && la.w_hi && la.w_lo && !(la.flags2 & LGCN_EPI_GN) &&
                               !(la.flags & (LGCN_EPI_RES | LGCN_EPI_RELU2))),
                 "linear_fused: a chained second Linear needs pre-split weights, no GroupNorm of its own and no residual on the first");
  if (int rc = set_fused_attrs()) return rc;
  const float *w_hi = la.w_hi, *w_lo = la.w_lo;
  int slot = -1;
  SplitRing& ring = g_rings[current_device()];
  if ((!w_hi || !w_lo) && la.split_ws) {   // split W into the caller's workspace (stream-ordered reuse is the caller's)
    float* s_hi = (float*)la.split_ws;
    float* s_lo = s_hi + kSlotFloats / 2;
    const int64_t ldw = (int64_t)la.n_src * LGCN_C + la.ks;
    k_split_blocks<<<lgcn_cdiv(la.n_src * 128 * 32, 256), 256, 0, st>>>(la.W, ldw, la.n_src, s_hi, s_lo);
    LGCN_LAUNCH_OK();
    w_hi = s_hi;
    w_lo = s_lo;
  }
  if (!w_hi || !w_lo) {   // split W into a slot of the library's own scratch ring
    std::lock_guard<std::mutex> lock(g_ring_mu);
    if (!ring.buf) {
      LGCN_CUDA_OK(cudaMalloc(&ring.buf, kRing * kSlotFloats * sizeof(float)));
      for (int i = 0; i < kRing; ++i) {
        LGCN_CUDA_OK(cudaEventCreateWithFlags(&ring.ev[i], cudaEventDisableTiming));
        ring.used[i] = false;
      }
    }
    slot = ring.next;
    ring.next = (ring.next + 1) % kRing;
    if (ring.used[slot]) LGCN_CUDA_OK(cudaStreamWaitEvent(st, ring.ev[slot], 0));
    float* s_hi = ring.buf + slot * kSlotFloats;
    float* s_lo = s_hi + kSlotFloats / 2;
    const int64_t ldw = (int64_t)la.n_src * LGCN_C + la.ks;
    k_split_blocks<<<lgcn_cdiv(la.n_src * 128 * 32, 256), 256, 0, st>>>(la.W, ldw, la.n_src, s_hi, s_lo);
    LGCN_LAUNCH_OK();
    w_hi = s_hi;
    w_lo = s_lo;
  }
  CUtensorMap map, mhi, mlo;
  if (int rc = make_out_map(&map, la.out, LGCN_C, la.m, la.ldo)) return rc;
  if (int rc = make_map_2d(&mhi, w_hi, LGCN_C, (int64_t)(la.n_src + (la.chain ? 1 : 0)) * LGCN_C, LGCN_C, 32, 128)) return rc;
  if (int rc = make_cross_map(&mlo, w_lo, (int64_t)(la.n_src + (la.chain ? 1 : 0)) * LGCN_C)) return rc;
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  for (int k = 0; k < la.n_src; ++k) {
    a.src[k] = la.a[k];
    a.idx[k] = la.idx[k];
  }
  if (la.head_w) {   // source 0 is computed in the kernel: nothing is read through src[0] (any valid address will do)
    a.src[0] = la.out;
    a.idx[0] = nullptr;
    a.head_w = la.head_w; a.head_p = (const float2*)la.head_p; a.head_ip = la.head_ip;
    a.head_q = (const float2*)la.head_q; a.head_iq = la.head_iq;
  }
  a.X = a.src[0]; a.XA = a.src[0]; a.tab = nullptr;
  a.gn = (la.flags & LGCN_EPI_GN) ? la.gamma : nullptr;
  a.beta = la.beta;
  a.res = la.res ? la.res : a.src[0];
  if (la.ks == 4) {
    a.xs = la.xs;
    a.wx = la.W + (int64_t)la.n_src * LGCN_C;
    a.ldw = (int64_t)la.n_src * LGCN_C + 4;
  }
  a.flags = la.flags; a.M = la.m; a.m_dev = la.m_dev; a.n_keys = la.n_src - 1; a.chain = la.chain ? 1 : 0; a.flags2 = la.flags2; a.dbg = lgcn_debug_get(); a.tl = g_timeline;
  const int64_t n_tiles = (la.m + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(n_tiles < num_sms() ? n_tiles : num_sms());
  if (la.head_w) LGCN_CUDA_OK(lgcn_launch_pdl(k_laneconv_fused<true, false, false, true>, grid, kNumThreads, kSmemTotal, st, a, map, mhi, mlo));
  else if (la.chain) LGCN_CUDA_OK(lgcn_launch_pdl(k_laneconv_fused<true, false, true>, grid, kNumThreads, kSmemTotal, st, a, map, mhi, mlo));
  else if (la.ks == 4) LGCN_CUDA_OK(lgcn_launch_pdl(k_laneconv_fused<true, true>, grid, kNumThreads, kSmemTotal, st, a, map, mhi, mlo));
  else LGCN_CUDA_OK(lgcn_launch_pdl(k_laneconv_fused<true, false>, grid, kNumThreads, kSmemTotal, st, a, map, mhi, mlo));
  LGCN_LAUNCH_OK();
  if (slot >= 0) {
    std::lock_guard<std::mutex> lock(g_ring_mu);
    LGCN_CUDA_OK(cudaEventRecord(ring.ev[slot], st));
    ring.used[slot] = true;
  }
  return 0;
}

int64_t lgcn_laneconv_fused_aux_bytes(int64_t n_edges) { return lgcn_align_up((n_edges / 2 + 1) * LGCN_C * 4, 1024); }

// out[m] = one LaneConv block of x (chain = 1) or relu(GN(sum_k W_k agg_k)) (chain = 0).  w_hi / w_lo: pre-split
// [(n_keys + 1 (+1 with chain: ctr2)) * 128, 128]; gn: gamma1 | beta1 | gamma2 | beta2; xa: aux rows workspace.
int lgcn_launch_laneconv_fused(const float* x, float* out, void* plan, int64_t n_nodes, const int32_t* n_dev,
                               int64_t n_edges, int n_keys, const float* w_hi, const float* w_lo, const float* gn,
                               float* xa, int chain, cudaStream_t st) {
  if (n_nodes <= 0) return 0;
  LGCN_CHECK_ARG(x != out, "laneconv_fused: in-place is not possible (neighbour rows are read by other tiles)");
  if (int rc = set_fused_attrs()) return rc;
  PlanView v = plan_view(plan, n_nodes, n_edges, n_keys);
  if (n_keys > 0 && n_edges > 1) {
    k_multi_sum<<<num_sms() * 4, 256, 0, st>>>(x, v.hdr, v.mdesc, v.mcol, xa, v.max_multi);
    LGCN_LAUNCH_OK();
  }
  // the chain form runs on the second-generation kernel (laneconv_v2.cu); debug flag 2048 keeps this one
  if (chain && !LGCN_BF16_CROSS && !(lgcn_debug_get() & 2048)) {
    LgcnProfScope ps(LGCN_PROF_BLOCK_KERNEL, st);   // the dominant kernel on its own (bench.py's roofline)
    return lgcn_launch_laneconv_v2(x, xa, v.tab, out, n_nodes, n_dev, n_keys, w_hi, w_lo, gn, st);
  }
  const int nkw = n_keys + 1 + (chain ? 1 : 0);
  CUtensorMap map, mhi, mlo;
  if (int rc = make_out_map(&map, out, LGCN_C, n_nodes, LGCN_C)) return rc;
  if (int rc = make_map_2d(&mhi, w_hi, LGCN_C, (int64_t)nkw * LGCN_C, LGCN_C, 32, 128)) return rc;
  if (int rc = make_cross_map(&mlo, w_lo, (int64_t)nkw * LGCN_C)) return rc;
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.X = x; a.XA = xa; a.tab = v.tab; a.gn = gn; a.M = n_nodes; a.m_dev = n_dev; a.n_keys = n_keys; a.chain = chain; a.dbg = lgcn_debug_get();
  a.tl = g_timeline;
  const int64_t n_tiles = (n_nodes + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(n_tiles < num_sms() ? n_tiles : num_sms());
  LgcnProfScope ps(LGCN_PROF_BLOCK_KERNEL, st);
  k_laneconv_fused<false><<<grid, kNumThreads, kSmemTotal, st>>>(a, map, mhi, mlo);
  LGCN_LAUNCH_OK();
  return 0;
}
