// gemm_tc_wide.cu — the LaneConv wide projection  Y[N, nob*128] = X[N,128] . Wcat[nob*128,128]^T  (nob = 15:
// ctr + 14 edge keys, lanegcn.py:332-354) on tcgen05, 3xTF32 (see gemm_tc.cu for the precision scheme).
//
// Why a second kernel: in gemm_tc.cu the resident 128-row A tile (hi+lo = 128 KB of shared memory) leaves room for
// only two 32 KB weight stages = 1,536 MMA cycles of buffered work, less than the ~1,600-cycle hand-off round trip
// (tcgen05.commit -> mbarrier -> producer wake -> convert/store -> fence -> MMA warp wake) measured with
// tools/ablate_gemm.py, so the tensor pipe idles (34 % active).  Here the A tile lives in TENSOR MEMORY instead
// (tcgen05.mma with the A operand from TMEM: row -> lane, k -> column, one 32-bit element per column), which frees
// shared memory for a 3-stage ring of WHOLE-K weight tiles:
//
//   TMEM (512 columns):  A_hi [0,128) | A_lo [128,256) | 2 accumulator stages x (main 64 | cross 64) [256,512)
//   smem (224 KB):       3 stages x 64 KB  (64 weight rows x K=128, hi | lo, K-major SWIZZLE_128B chunks of 8 KB)
//                        | 8 x 4 KB store staging | barriers
//
// One MMA stage = one 64-column output tile = 16 k-steps x 3 instructions (M=128, N=64, K=8) = 1,536 tensor
// cycles, so the ring buffers 4,608 cycles of work and one hand-off per 1,536 cycles instead of one per 768.
//
//   warp 0      TMA producer (one elected lane): the weights are static, so they are split into W_hi / W_lo ONCE
//               (k_split_tf32, a few microseconds per LaneConv block) and every stage is eight
//               cp.async.bulk.tensor loads (SWIZZLE_128B boxes of 64 rows x 32 floats) completing on the stage's
//               mbarrier: no conversion instructions, no proxy fences, no thread arrives on the weight path.  A
//               first version converted the weights in four producer warps: they, not the tensor pipe, bounded
//               the kernel (555 us; tools/ablate_gemm.py).
//   warp 1      MMA issuer (elect.sync).
//   warps 4-11  epilogue (warps 2-3 idle; warp ids keep warp % 4 == TMEM lane quarter): TMEM quarter e&3, 32-column half e>>2 of the 64-column tile: tcgen05.ld main + cross,
//               add, swizzled staging, TMA bulk tensor store (32x32 fp32 boxes).  These warps also own the A
//               tile: warp (q,h) prefetches columns [64h, 64h+64) of row 32q+lane of the NEXT row tile into
//               registers during the last output tile and stores them to TMEM (hi/lo) the moment the MMAs of the
//               current row tile have retired, so neither the weight ring nor the tensor pipe drains between
//               row tiles (a first version loaded A from the producers and took 727 us instead of 649).
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kTileM = 128;
constexpr int kTileN = 64;
constexpr int kStageBytes = 2 * kTileN * 512;     // hi 32 KB | lo 32 KB
constexpr int kChunk = kTileN * 128;              // one K-chunk (32 floats) of 64 rows: 8 KB
constexpr int kStages = 3;
constexpr int kSmemOut = kStages * kStageBytes;   // 192 KB
constexpr int kSmemBar = kSmemOut + 8 * 4096;     // 224 KB
constexpr int kSmemTotal = kSmemBar + 128;
constexpr int kNumThreads = 384;
constexpr int kMmaWarp = 1;
constexpr uint32_t kIdesc = idesc_tf32(128, kTileN);
constexpr uint32_t kColAHi = 0, kColALo = 128, kColAcc = 256;  // TMEM column map

__global__ void __launch_bounds__(kNumThreads, 1)
k_wide_tc(const LinearArgs a, const __grid_constant__ CUtensorMap out_map, const __grid_constant__ CUtensorMap whi_map,
          const __grid_constant__ CUtensorMap wlo_map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  if (sbase & 1023u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_a_full = sbase + kSmemBar + 0, bar_a_empty = sbase + kSmemBar + 8;
  const uint32_t bar_b_full = sbase + kSmemBar + 16;     // [3]
  const uint32_t bar_b_empty = sbase + kSmemBar + 40;    // [3]
  const uint32_t bar_acc_full = sbase + kSmemBar + 64;   // [2]
  const uint32_t bar_acc_empty = sbase + kSmemBar + 80;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + 96);

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 8);   // one arrive per epilogue warp
    mbar_init(bar_a_empty, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_b_full + 8 * i, 1);   // the producer's expect_tx arrive; TMA completes the bytes
      mbar_init(bar_b_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

  // scalars only below this point: lambdas that captured the by-value struct `a` by reference forced it into
  // local memory (and with ~3 KB of L1 left every access to it was an L2 round trip)
  const int64_t M = a.m;
  const int dbg = a.dbg;
  const float* __restrict__ X = a.a[0];
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int n_nt = a.n_out_blocks * 2;               // 64-column output tiles per row tile
  const int nt0 = (int)((blockIdx.x * 2) % n_nt);    // CTAs start at different weight tiles (spread L2 requests)

  if (warp == 0) {
    // =========================================================== TMA producer: the weight ring, never drained
    uint32_t b_phase = 0;
    int b_stage = 0, nt = nt0;
    const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (int64_t s = 0; s < my_tiles * n_nt; ++s) {
      mbar_wait(bar_b_empty + 8 * b_stage, b_phase ^ 1);
      if (elect_one()) {
        const uint32_t bar = bar_b_full + 8 * b_stage, dst = sbase + b_stage * kStageBytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kStageBytes) : "memory");
        if (!(dbg & 8)) {
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {  // box = 32 floats (one 128 B swizzle row) x 64 weight rows = 8 KB
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(dst + kc * kChunk), "l"(reinterpret_cast<uint64_t>(&whi_map)), "r"(bar), "r"(kc * 32), "r"(nt * kTileN)
                : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(dst + kTileN * 512 + kc * kChunk), "l"(reinterpret_cast<uint64_t>(&wlo_map)), "r"(bar), "r"(kc * 32),
                "r"(nt * kTileN)
                : "memory");
          }
        } else {  // ablation: no loads — complete the transaction count by hand
          asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"((uint32_t)kStageBytes) : "memory");
        }
      }
      __syncwarp();
      if (++nt == n_nt) nt = 0;
      if (++b_stage == kStages) {
        b_stage = 0;
        b_phase ^= 1;
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================================================== MMA issuer
    uint32_t a_full_phase = 0, b_phase = 0, acc_phase = 0;
    int b_stage = 0, acc_stage = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(bar_a_full, a_full_phase);
      a_full_phase ^= 1;
      for (int nt = 0; nt < n_nt; ++nt) {
        mbar_wait(bar_acc_empty + 8 * acc_stage, acc_phase ^ 1);
        mbar_wait(bar_b_full + 8 * b_stage, b_phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_main = tmem_base + kColAcc + acc_stage * 128, d_cross = d_main + 64;
          const uint32_t w_hi = sbase + b_stage * kStageBytes, w_lo = w_hi + kTileN * 512;
          if (!(dbg & 4)) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {  // k-step j: K-chunk j>>2, 32-byte slice j&3; A columns 8j..8j+7
              const uint32_t off = (j >> 2) * kChunk + (j & 3) * 32;
              const uint32_t acc = j == 0 ? 0u : 1u;
              umma_tf32_ts(d_cross, tmem_base + kColALo + 8 * j, umma_desc(w_hi + off), kIdesc, acc);
              umma_tf32_ts(d_cross, tmem_base + kColAHi + 8 * j, umma_desc(w_lo + off), kIdesc, 1u);
              umma_tf32_ts(d_main, tmem_base + kColAHi + 8 * j, umma_desc(w_hi + off), kIdesc, acc);
            }
          }
          umma_commit(bar_b_empty + 8 * b_stage);
          umma_commit(bar_acc_full + 8 * acc_stage);
          if (nt == n_nt - 1) umma_commit(bar_a_empty);
        }
        __syncwarp();
        if (++b_stage == kStages) {
          b_stage = 0;
          b_phase ^= 1;
        }
        if (++acc_stage == 2) {
          acc_stage = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // =========================================================== epilogue + A-tile owner: warps 4..11
    const int e = warp - 4, q = e & 3, h = e >> 2;
    const uint32_t my_buf = sbase + kSmemOut + e * 4096;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_phase = 0, a_empty_phase = 0;
    int acc_stage = 0;
    float4 ax[16];  // columns [64h, 64h+64) of row 32q+lane of the next row tile
    auto a_prefetch = [&](int64_t t) {
      const int64_t m = t * kTileM + q * 32 + lane;
      const bool ok = t < n_tiles && m < M && !(dbg & 8);
      const float4* src = reinterpret_cast<const float4*>(X + (ok ? m : 0) * LGCN_C + h * 64);
#pragma unroll
      for (int c = 0; c < 16; ++c) ax[c] = ok ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto a_store = [&]() {  // registers -> hi/lo -> TMEM (16 columns per tcgen05.st), then publish
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 x = ax[4 * g + c];
          const float h0 = tf32_rna(x.x), h1 = tf32_rna(x.y), h2 = tf32_rna(x.z), h3 = tf32_rna(x.w);
          hi[4 * c] = __float_as_uint(h0); lo[4 * c] = __float_as_uint(tf32_rna(x.x - h0));
          hi[4 * c + 1] = __float_as_uint(h1); lo[4 * c + 1] = __float_as_uint(tf32_rna(x.y - h1));
          hi[4 * c + 2] = __float_as_uint(h2); lo[4 * c + 2] = __float_as_uint(tf32_rna(x.z - h2));
          hi[4 * c + 3] = __float_as_uint(h3); lo[4 * c + 3] = __float_as_uint(tf32_rna(x.w - h3));
        }
        TMEM_ST16(t_lane + kColAHi + h * 64 + g * 16, hi, 0);
        TMEM_ST16(t_lane + kColALo + h * 64 + g * 16, lo, 0);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a_full);
    };
    a_prefetch(blockIdx.x);
    a_store();  // first row tile: nothing reads A yet
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t m0 = t * kTileM;
      for (int i = 0; i < n_nt; ++i) {
        int nt = nt0 + i;
        if (nt >= n_nt) nt -= n_nt;
        if (i == n_nt - 1) a_prefetch(t + gridDim.x);  // next row tile's A rows: in flight during this tile
        mbar_wait(bar_acc_full + 8 * acc_stage, acc_phase);
        tc_fence_after();
        // main + cross accumulators, 16 columns at a time (keeps the register peak under the 128 cap while the
        // 64 registers of the prefetched A rows are live)
        const uint32_t taddr = tmem_base + kColAcc + acc_stage * 128 + ((uint32_t)(q * 32) << 16) + h * 32;
        float4 o[8];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t v[16], x[16];
          TMEM_LD16(v, 0, taddr + hf * 16);
          if (!(dbg & 2)) TMEM_LD16(x, 0, taddr + 64 + hf * 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            o[hf * 4 + c].x = __uint_as_float(v[4 * c]) + __uint_as_float(x[4 * c]);
            o[hf * 4 + c].y = __uint_as_float(v[4 * c + 1]) + __uint_as_float(x[4 * c + 1]);
            o[hf * 4 + c].z = __uint_as_float(v[4 * c + 2]) + __uint_as_float(x[4 * c + 2]);
            o[hf * 4 + c].w = __uint_as_float(v[4 * c + 3]) + __uint_as_float(x[4 * c + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc_stage);
        if (++acc_stage == 2) {
          acc_stage = 0;
          acc_phase ^= 1;
        }
        if (i == n_nt - 1 && t + gridDim.x < n_tiles) {
          // the commit that published this last accumulator also released A: refill it for the next row tile
          mbar_wait(bar_a_empty, a_empty_phase);
          a_empty_phase ^= 1;
          tc_fence_after();
          a_store();
        }
        if (dbg & 1) continue;
        if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging buffer free again
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) st_shared_f4(my_buf + lane * 128 + ((c ^ (lane & 7)) << 4), o[c]);
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&out_map)),
                       "r"(my_buf), "r"(nt * kTileN + h * 32), "r"((int32_t)(m0 + q * 32))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}


// hi = tf32(w), lo = tf32(w - hi) for a static weight matrix (once per LaneConv block, ~1 MB)
__global__ void k_split_tf32(const float4* __restrict__ w, float4* __restrict__ hi, float4* __restrict__ lo, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = w[i];
  float4 h, l;
  h.x = tf32_rna(x.x); l.x = tf32_rna(x.x - h.x);
  h.y = tf32_rna(x.y); l.y = tf32_rna(x.y - h.y);
  h.z = tf32_rna(x.z); l.z = tf32_rna(x.z - h.z);
  h.w = tf32_rna(x.w); l.w = tf32_rna(x.w - h.w);
  hi[i] = h;
  lo[i] = l;
}

}  // namespace

int lgcn_split_tf32(const float* w, float* hi, float* lo, int64_t n, cudaStream_t st) {
  LGCN_CHECK_ARG(n % 4 == 0, "split_tf32: n %% 4 != 0");
  if (n == 0) return 0;
  k_split_tf32<<<lgcn_cdiv(n / 4, 256), 256, 0, st>>>((const float4*)w, (float4*)hi, (float4*)lo, n / 4);
  LGCN_LAUNCH_OK();
  return 0;
}

// Y[m, nob*128] = X . W^T with the weights pre-split (w_hi / w_lo: [nob*128, 128] fp32 holding tf32 values)
int lgcn_launch_wide_tc(const LinearArgs& a, const float* w_hi, const float* w_lo, cudaStream_t st) {
  if (a.m <= 0) return 0;
  LGCN_CHECK_ARG(a.n_src == 1 && a.idx[0] == nullptr && a.flags == 0 && a.ks == 0 && w_hi && w_lo,
                 "wide projection kernel: one un-gathered source, no epilogue, pre-split weights");
  if (first_use(kFamWideTc))
    LGCN_CUDA_OK(cudaFuncSetAttribute(k_wide_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  CUtensorMap map, mhi, mlo;
  if (int rc = make_out_map(&map, a.out, (int64_t)a.n_out_blocks * LGCN_C, a.m, a.ldo)) return rc;
  if (int rc = make_map_2d(&mhi, w_hi, LGCN_C, (int64_t)a.n_out_blocks * LGCN_C, LGCN_C, 32, kTileN)) return rc;
  if (int rc = make_map_2d(&mlo, w_lo, LGCN_C, (int64_t)a.n_out_blocks * LGCN_C, LGCN_C, 32, kTileN)) return rc;
  const int64_t n_tiles = (a.m + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(n_tiles < num_sms() ? n_tiles : num_sms());
  k_wide_tc<<<grid, kNumThreads, kSmemTotal, st>>>(a, map, mhi, mlo);
  LGCN_LAUNCH_OK();
  return 0;
}
