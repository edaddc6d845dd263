// gemm_tc.cu — tcgen05 (5th-gen tensor core) implementation of lgcn_linear128 (engine 1), sm_100a only.
//
//   out[m, ob*128 + c] = epi( sum_s A_s[idx_s[m], :] . W[ob*128 + c, s*128:(s+1)*128] )      fp32 in, fp32 out
//
// Precision: 3xTF32.  Every fp32 operand x is split into hi = tf32(x) and lo = tf32(x - hi) (both rounded to
// nearest), and   D = sum A_hi.B_hi  +  sum (A_lo.B_hi + A_hi.B_lo)   is accumulated in fp32 in TMEM — in TWO
// accumulators.  Measured on B200 (tools/error_budget.py): the tensor core truncates toward zero once per MMA
// instruction (about -1.3e-8 relative per instruction), so folding the 32 cross-term instructions into the
// main accumulator triples its bias; a separate cross accumulator (2^-11 smaller, so its own truncation is
// negligible) is added in the epilogue with one IEEE fp32 add.  Result: rms error 3.4e-7 of the output rms
// (true fp32: 2.0e-7), inside the 1e-4 / 1e-5 budget where plain TF32 is not.
//
// Structure (one persistent CTA per SM, 416 threads, warp-specialised, mbarrier pipelines):
//   warps 0-3   producers.  The RESIDENT operand (128 rows x 128 floats) is loaded with coalesced 128-bit
//               loads, split into hi/lo and stored in the K-major SWIZZLE_128B layout the UMMA shared-memory
//               descriptor expects (4 K-chunks of 32 floats; chunk = [128 rows][128 B]).  The STREAMED operand
//               goes through a 2-stage ring of single K-chunks (hi+lo = 32 KB/stage); the global loads of
//               stage q+1 are issued into registers before stage q is converted, so L2/HBM latency overlaps
//               the MMAs.  Two modes:
//                 weight-resident (one source, one output block: ctr2, Att/MLP linears): W is loaded once per
//                   CTA and the (optionally row-gathered) rows of every M-tile stream from HBM;
//                 tile-resident (the 15-block wide LaneConv projection; the 3-source Att ctx.0): the 128-row A
//                   tile is resident and the weight tiles stream from L2, each CTA starting at a different
//                   output block so 148 SMs do not hit the same L2 lines at once.
//   warp 12     MMA issuer (one elected lane): per stage 4 k-steps x 3 tcgen05.mma.kind::tf32 (M=128, N=128,
//               K=8); accumulators in TMEM, 2 stages x [main 128 | cross 128] columns = all 512 columns, so
//               the epilogue of tile i overlaps the MMAs of tile i+1.  tcgen05.commit releases smem stages and
//               publishes accumulators.
//   warps 4-11  epilogue: warp e owns TMEM lane quarter e&3 (32 rows) and column half e>>2 (64 columns).
//               With 225 KB of shared memory carved out the SM has almost no L1, so every global load in the
//               epilogue is an L2 round trip: two warps per scheduler hide each other's latency, the residual
//               is prefetched before the GroupNorm exchange, gamma/beta live in shared memory.  GroupNorm(1)
//               statistics: each half computes (mean, M2) of its 64 values, the two halves are combined with
//               Chan's formula through shared memory (two 64-thread named barriers).  (Tried and rejected: two
//               teams of four warps alternating accumulator stages with each thread owning a full row re-read
//               from TMEM in two passes — correct, but 127 us instead of 109 us per ctr2 launch: TMEM->register
//               reads are bandwidth-limited, so every accumulator value should be read exactly once.)  Results are staged as
//               32x32 fp32 blocks in swizzled shared memory and written with TMA bulk tensor stores (coalesced
//               128 B rows; rows >= M clipped by the tensor map).
#include <atomic>

#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kTileM = 128;
constexpr int kChunkBytes = kTileM * 128;  // one K-chunk (32 floats) of a 128-row operand: 16 KB
constexpr int kSmemResHi = 0;
constexpr int kSmemResLo = 4 * kChunkBytes;  // 64 KB
constexpr int kSmemStream = 8 * kChunkBytes;  // 128 KB: stages of [hi 16 KB | lo 16 KB]
constexpr int kStages = 2;
constexpr int kSmemOut = kSmemStream + kStages * 2 * kChunkBytes;  // 192 KB: 8 epilogue warps x 4 KB
constexpr int kSmemBar = kSmemOut + 8 * 4096;                       // 224 KB
constexpr int kSmemGamma = kSmemBar + 128;                          // gamma[128] | beta[128]
constexpr int kSmemTotal = kSmemGamma + 1024;
constexpr int kNumThreads = 416;
constexpr int kMmaWarp = 12;
constexpr uint32_t kTmemCols = 512;  // 2 stages x (main + cross accumulator) x 128 columns
constexpr uint32_t kIdesc = idesc_tf32(128, 128);

__global__ void __launch_bounds__(kNumThreads, 1)
k_linear_tc(const LinearArgs a, const __grid_constant__ CUtensorMap out_map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  if (sbase & 1023u) __trap();  // SWIZZLE_128B operands need 1024-byte aligned blocks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_a_full = sbase + kSmemBar + 0, bar_a_empty = sbase + kSmemBar + 8;
  const uint32_t bar_b_full = sbase + kSmemBar + 16;     // [kStages]
  const uint32_t bar_b_empty = sbase + kSmemBar + 32;    // [kStages]
  const uint32_t bar_acc_full = sbase + kSmemBar + 48;   // [2]
  const uint32_t bar_acc_empty = sbase + kSmemBar + 64;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + 96);
  const bool gn = a.flags & LGCN_EPI_GN;

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 4);       // one elected arrive per producer warp
    mbar_init(bar_a_empty, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_b_full + 8 * i, 4);
      mbar_init(bar_b_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 8);  // one elected arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (gn && threadIdx.x < 64) {  // gamma | beta -> shared memory (broadcast reads in the epilogue)
    const float* src = threadIdx.x < 32 ? a.gamma : a.beta;
    reinterpret_cast<float4*>(smem + kSmemGamma)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(src) + (threadIdx.x & 31));
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(tmem_slot)),
                 "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_tiles = (a.m + kTileM - 1) / kTileM;
  const int64_t ldw = (int64_t)a.n_src * LGCN_C + a.ks;
  const int nob = a.n_out_blocks;
  const bool w_resident = (a.n_src == 1 && nob == 1);  // see the header comment

  if (warp < 4) {
    // =========================================================== producers
    const int tid = threadIdx.x;                         // 0..127
    const int ch16 = lane & 7, kc_of_lane = lane >> 3;   // resident rows: lane l owns float4 #l of the 512 B row
    const int srow = tid >> 3, sch = tid & 7;            // streamed chunk: rows srow + 16 i, 16-byte chunk sch
    uint32_t a_empty_phase = 0, b_phase = 0;
    int b_stage = 0;

    auto load_resident = [&](auto&& row_ptr /* (row) -> const float* or null */) {
#pragma unroll 1
      for (int r0 = warp * 32; r0 < warp * 32 + 32; r0 += 8) {
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float* p = row_ptr(r0 + i);
          x[i] = p ? __ldg(reinterpret_cast<const float4*>(p) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          split_store(sbase + kSmemResHi + kc_of_lane * kChunkBytes, sbase + kSmemResLo + kc_of_lane * kChunkBytes,
                      r0 + i, ch16, x[i]);
      }
      // every thread fences its own generic-proxy stores, the warp converges, ONE lane arrives: 128 arrives on
      // one mbarrier serialise in shared memory (measured: ~1500 cycles per stage round trip with them)
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a_full);
    };
    auto stream = [&](int64_t n_tasks, auto&& chunk_ptr /* (task, row) -> const float* or null */) {
      float4 nxt[8];
      auto issue = [&](int64_t q) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float* p = (a.dbg & 8) ? nullptr : chunk_ptr(q, srow + 16 * i);
          nxt[i] = p ? __ldg(reinterpret_cast<const float4*>(p) + sch) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if (n_tasks > 0) issue(0);
      for (int64_t q = 0; q < n_tasks; ++q) {
        float4 cur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
        if (q + 1 < n_tasks) issue(q + 1);  // next stage's loads fly while this one is converted / waited for
        mbar_wait(bar_b_empty + 8 * b_stage, b_phase ^ 1);
        const uint32_t hi_blk = sbase + kSmemStream + b_stage * 2 * kChunkBytes;
        if (!(a.dbg & 8)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) split_store(hi_blk, hi_blk + kChunkBytes, srow + 16 * i, sch, cur[i]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_b_full + 8 * b_stage);
        if (++b_stage == kStages) {
          b_stage = 0;
          b_phase ^= 1;
        }
      }
    };

    if (w_resident) {
      load_resident([&](int r) { return a.W + (int64_t)r * ldw; });
      const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const float* __restrict__ src = a.a[0];
      const int32_t* __restrict__ idx = a.idx[0];
      stream(my_tiles * 4, [&](int64_t q, int r) -> const float* {
        const int64_t m = (blockIdx.x + (q >> 2) * gridDim.x) * kTileM + r;
        if (m >= a.m) return nullptr;
        const int64_t row = idx ? (int64_t)__ldg(idx + m) : m;
        return src + row * LGCN_C + (q & 3) * 32;
      });
    } else {
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t m0 = t * kTileM;
        for (int s = 0; s < a.n_src; ++s) {
          mbar_wait(bar_a_empty, a_empty_phase ^ 1);  // MMAs reading the previous resident tile have retired
          a_empty_phase ^= 1;
          const float* __restrict__ src = a.a[s];
          const int32_t* __restrict__ idx = a.idx[s];
          load_resident([&](int r) -> const float* {
            const int64_t m = m0 + r;
            if (m >= a.m) return nullptr;
            return src + (idx ? (int64_t)__ldg(idx + m) : m) * LGCN_C;
          });
          const float* __restrict__ wsrc = a.W + (int64_t)s * LGCN_C;
          const int ob0 = blockIdx.x % nob;
          stream((int64_t)nob * 4, [&](int64_t q, int r) -> const float* {
            int ob = ob0 + (int)(q >> 2);
            if (ob >= nob) ob -= nob;
            return wsrc + ((int64_t)ob * LGCN_C + r) * ldw + (q & 3) * 32;
          });
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================================================== MMA issuer
    uint32_t a_full_phase = 0, b_phase = 0, acc_phase = 0;
    int b_stage = 0, acc_stage = 0;
    // one K-chunk: 4 k-steps (K = 8 floats = 32 B) x 3 products; M operand = node rows, N operand = weight rows
    auto issue_stage = [&](uint32_t d_main, uint32_t d_cross, uint32_t x_hi, uint32_t x_lo, uint32_t w_hi, uint32_t w_lo,
                           bool fresh) {
      if (a.dbg & 4) return;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t acc = (fresh && j == 0) ? 0u : 1u;
        umma_tf32(d_cross, umma_desc(x_lo + 32 * j), umma_desc(w_hi + 32 * j), kIdesc, acc);
        umma_tf32(d_cross, umma_desc(x_hi + 32 * j), umma_desc(w_lo + 32 * j), kIdesc, 1u);
        umma_tf32(d_main, umma_desc(x_hi + 32 * j), umma_desc(w_hi + 32 * j), kIdesc, acc);
      }
    };
    auto next_stage = [&]() {
      if (++b_stage == kStages) {
        b_stage = 0;
        b_phase ^= 1;
      }
    };
    auto next_acc = [&]() {
      if (++acc_stage == 2) {
        acc_stage = 0;
        acc_phase ^= 1;
      }
    };
    if (w_resident) {
      mbar_wait(bar_a_full, 0);  // the weight is resident for the whole kernel
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(bar_acc_empty + 8 * acc_stage, acc_phase ^ 1);
        const uint32_t d_main = tmem_base + acc_stage * 256, d_cross = d_main + 128;
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait(bar_b_full + 8 * b_stage, b_phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t w_hi = sbase + kSmemResHi + kc * kChunkBytes, w_lo = sbase + kSmemResLo + kc * kChunkBytes;
            const uint32_t x_hi = sbase + kSmemStream + b_stage * 2 * kChunkBytes, x_lo = x_hi + kChunkBytes;
            issue_stage(d_main, d_cross, x_hi, x_lo, w_hi, w_lo, kc == 0);
            umma_commit(bar_b_empty + 8 * b_stage);
            if (kc == 3) umma_commit(bar_acc_full + 8 * acc_stage);
          }
          __syncwarp();
          next_stage();
        }
        next_acc();
      }
    } else {
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int s = 0; s < a.n_src; ++s) {
          mbar_wait(bar_a_full, a_full_phase);
          a_full_phase ^= 1;
          for (int ob = 0; ob < nob; ++ob) {
            if (s == 0) mbar_wait(bar_acc_empty + 8 * acc_stage, acc_phase ^ 1);  // fresh accumulator stage
            const uint32_t d_main = tmem_base + acc_stage * 256, d_cross = d_main + 128;
            for (int kc = 0; kc < 4; ++kc) {
              mbar_wait(bar_b_full + 8 * b_stage, b_phase);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t x_hi = sbase + kSmemResHi + kc * kChunkBytes, x_lo = sbase + kSmemResLo + kc * kChunkBytes;
                const uint32_t w_hi = sbase + kSmemStream + b_stage * 2 * kChunkBytes, w_lo = w_hi + kChunkBytes;
                issue_stage(d_main, d_cross, x_hi, x_lo, w_hi, w_lo, s == 0 && kc == 0);
                umma_commit(bar_b_empty + 8 * b_stage);  // smem stage reusable once these MMAs retire
                if (kc == 3 && s == a.n_src - 1) umma_commit(bar_acc_full + 8 * acc_stage);
                if (kc == 3 && ob == nob - 1) umma_commit(bar_a_empty);
              }
              __syncwarp();
              next_stage();
            }
            if (s == a.n_src - 1) next_acc();  // accumulator published: move to the other TMEM stage
          }
        }
      }
    }
  } else {
    // =========================================================== epilogue: warps 4..11
    const int e = warp - 4, q = e & 3, h = e >> 2;  // TMEM lane quarter (== warp % 4), column half
    const uint32_t my_buf = sbase + kSmemOut + e * 4096, partner_buf = sbase + kSmemOut + (e ^ 4) * 4096;
    const uint32_t gam = sbase + kSmemGamma + h * 256, bet = gam + 512;
    uint32_t acc_phase = 0;
    int acc_stage = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t m0 = t * kTileM;
      const int64_t m = m0 + q * 32 + lane;  // this thread's output row
      const bool live = m < a.m;
      for (int obi = 0; obi < nob; ++obi) {
        int ob = (int)(blockIdx.x % nob) + obi;  // same (staggered) order as the producers stream the weights
        if (ob >= nob) ob -= nob;
        mbar_wait(bar_acc_full + 8 * acc_stage, acc_phase);
        tc_fence_after();
        uint32_t v[64];
        const uint32_t taddr = tmem_base + acc_stage * 256 + ((uint32_t)(q * 32) << 16) + h * 64;
        TMEM_LD32(v, 0, taddr);
        TMEM_LD32(v, 32, taddr + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {  // + cross-term accumulator (columns 128..255 of the stage)
          if (a.dbg & 2) break;
          uint32_t x[32];
          TMEM_LD32(x, 0, taddr + 128 + cb * 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < 32; ++c) v[cb * 32 + c] = __float_as_uint(__uint_as_float(v[cb * 32 + c]) + __uint_as_float(x[c]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc_stage);  // TMEM stage may be overwritten by the next MMAs
        if (++acc_stage == 2) {
          acc_stage = 0;
          acc_phase ^= 1;
        }
        float* f = reinterpret_cast<float*>(v);
        // residual of the first 32-column block: issued now (13 warps => 4 warps on one scheduler => 128
        // registers per thread, so not before the TMEM loads), consumed after the norm
        float4 r4[8];
        const float4* resp = reinterpret_cast<const float4*>(a.res + (live ? m : 0) * LGCN_C + h * 64);
        if (a.flags & LGCN_EPI_RES) {
#pragma unroll
          for (int c = 0; c < 8; ++c) r4[c] = live ? __ldg(resp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (a.ks > 0 && live) {  // rank-ks update: the 4 extra input columns of A2M.meta
          const float4 x = __ldg(reinterpret_cast<const float4*>(a.xs + m * 4));
          const float* wx = a.W + (int64_t)a.n_src * LGCN_C + (int64_t)(h * 64) * ldw;
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(wx + (int64_t)c * ldw));
            f[c] = fmaf(x.w, w.w, fmaf(x.z, w.z, fmaf(x.y, w.y, fmaf(x.x, w.x, f[c]))));
          }
        }
        // the staging buffer is reused every output block: the previous TMA store must have read it
        if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (gn) {
          // GroupNorm(1 group) over 128 channels = two 64-channel halves held by two warps: local (mean, M2),
          // exchanged through shared memory and combined with Chan's parallel-variance formula.
          float s1 = 0.f;
#pragma unroll
          for (int c = 0; c < 64; ++c) s1 += f[c];
          const float mean_h = s1 * (1.0f / 64.0f);
          float m2_h = 0.f;
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const float d = f[c] - mean_h;
            m2_h = fmaf(d, d, m2_h);
          }
          st_shared_f2(my_buf + lane * 8, mean_h, m2_h);
          named_bar_sync(1 + q, 64);
          const float2 o = ld_shared_f2(partner_buf + lane * 8);
          named_bar_sync(1 + q, 64);  // both halves have read: the buffers may now take output data
          const float mean = 0.5f * (mean_h + o.x);
          const float dm = mean_h - o.x;
          const float var = (m2_h + o.y + dm * dm * 32.0f) * (1.0f / 128.0f);
          const float rstd = 1.0f / sqrtf(var + LGCN_GN_EPS);
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            const float4 g = ld_shared_f4(gam + c * 4), b = ld_shared_f4(bet + c * 4);
            f[c] = fmaf((f[c] - mean) * rstd, g.x, b.x);
            f[c + 1] = fmaf((f[c + 1] - mean) * rstd, g.y, b.y);
            f[c + 2] = fmaf((f[c + 2] - mean) * rstd, g.z, b.z);
            f[c + 3] = fmaf((f[c + 3] - mean) * rstd, g.w, b.w);
          }
        }
        if (a.flags & LGCN_EPI_RELU1) {
#pragma unroll
          for (int c = 0; c < 64; ++c) f[c] = fmaxf(f[c], 0.f);
        }
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {  // two 32-row x 32-column blocks through the 4 KB staging buffer
          if (a.dbg & 1) break;
          if (a.flags & LGCN_EPI_RES) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              f[cb * 32 + 4 * c] += r4[c].x;
              f[cb * 32 + 4 * c + 1] += r4[c].y;
              f[cb * 32 + 4 * c + 2] += r4[c].z;
              f[cb * 32 + 4 * c + 3] += r4[c].w;
            }
            if (cb == 0) {  // residual of the second block: in flight while the first block is staged/stored
#pragma unroll
              for (int c = 0; c < 8; ++c) r4[c] = live ? __ldg(resp + 8 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (cb == 1) {
            if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 o = make_float4(f[cb * 32 + 4 * c], f[cb * 32 + 4 * c + 1], f[cb * 32 + 4 * c + 2], f[cb * 32 + 4 * c + 3]);
            if (a.flags & LGCN_EPI_RELU2) o = relu4(o);
            st_shared_f4(my_buf + lane * 128 + ((c ^ (lane & 7)) << 4), o);
          }
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&out_map)),
                         "r"(my_buf), "r"(ob * 128 + h * 64 + cb * 32), "r"((int32_t)(m0 + q * 32))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}


}  // namespace

namespace tc {
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
int make_map_2d(CUtensorMap* map, const float* base, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn encode = get_encode();
  LGCN_CHECK_ARG(encode != nullptr, "linear128(tcgen05): cuTensorMapEncodeTiled not available from the driver");
  LGCN_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "linear128(tcgen05): tensor must be 16-byte aligned");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LGCN_CHECK_ARG(r == CUDA_SUCCESS, "linear128(tcgen05): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}
// 2-D fp32 map with a 64-byte-wide box (16 floats) and SWIZZLE_64B: rows packed at 64 B, 16-byte chunk index XOR
// ((row >> 1) & 3)
int make_map_2d_sw64(CUtensorMap* map, const float* base, int64_t cols, int64_t rows, int64_t ld, int box_rows) {
  EncodeTiledFn encode = get_encode();
  LGCN_CHECK_ARG(encode != nullptr, "linear128(tcgen05): cuTensorMapEncodeTiled not available from the driver");
  LGCN_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "linear128(tcgen05): tensor must be 16-byte aligned");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LGCN_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (64 B box) failed (%d)", (int)r);
  return 0;
}
int make_map_2d_bf16(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn encode = get_encode();
  LGCN_CHECK_ARG(encode != nullptr, "linear128(tcgen05): cuTensorMapEncodeTiled not available from the driver");
  LGCN_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "linear128(tcgen05): tensor must be 16-byte aligned");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LGCN_CHECK_ARG(r == CUDA_SUCCESS, "linear128(tcgen05): cuTensorMapEncodeTiled (bf16) failed (%d)", (int)r);
  return 0;
}
int make_out_map(CUtensorMap* map, float* out, int64_t cols, int64_t rows, int64_t ldo) {
  return make_map_2d(map, out, cols, rows, ldo, 32, 32);
}
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}
int num_sms() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (!v) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
bool first_use(int family) {
  static std::atomic<int> seen[kMaxDevices][kFamCount];
  return seen[current_device()][family].exchange(1, std::memory_order_acq_rel) == 0;
}
}  // namespace tc

int lgcn_launch_linear_tc(const LinearArgs& a, cudaStream_t st) {
  if (a.m <= 0) return 0;
  // one output block: the aggregate-first kernel in linear mode (A operand through tensor memory,
  // TMA-fed pre-split weights) is 2-3x faster than k_linear_tc; debug flag 32768 keeps the old kernel
  if (a.n_out_blocks == 1 && (a.ks == 0 || a.ks == 4) && !(a.dbg & 32768)) return lgcn_launch_linear_fused(a, st);
  LGCN_CHECK_ARG(a.n_src == 1 || a.n_out_blocks == 1, "linear128(tcgen05): several sources need n_out_blocks == 1");
  // (the LaneConv stack routes its 15-block projection to gemm_tc_wide.cu, which needs pre-split weights)
  LGCN_CHECK_ARG(!a.m_dev, "linear128(tcgen05): a device-side row count needs n_out_blocks == 1");
  if (first_use(kFamLinearTc))
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  CUtensorMap map;
  if (int rc = make_out_map(&map, a.out, (int64_t)a.n_out_blocks * LGCN_C, a.m, a.ldo)) return rc;
  const int64_t n_tiles = (a.m + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(n_tiles < num_sms() ? n_tiles : num_sms());
  k_linear_tc<<<grid, kNumThreads, kSmemTotal, st>>>(a, map);
  LGCN_LAUNCH_OK();
  return 0;
}
