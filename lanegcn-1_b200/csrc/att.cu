// att.cu — the index side of Att (lanegcn.py:672-689: distance-thresholded pair list, batched, no per-scene
// host sync) and the K=2 "first layer" MLP heads (lanegcn.py:277-286, 644-648, 693).
#include "common.cuh"

// Bit-exact restatement of torch's  sqrt(((a - c) ** 2).sum(2)) <= th  in fp32: every operation rounded
// separately (nvcc would otherwise contract dx*dx + dy*dy into an FMA and flip borderline pairs).
__device__ __forceinline__ bool within(float ax, float ay, float cx, float cy, float th) {
  const float dx = __fsub_rn(ax, cx), dy = __fsub_rn(ay, cy);
  const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  return d <= th;
}

#define PAIR_WARPS 16   // kIndexThreads / 32

// FILL == false: cnt[r] = number of context rows of r's scene within th.
// FILL == true : write the pairs of row r at row_start[r].., hi = local row + hi_off[b], wi = j + wi_off[b]; entries at
//                positions >= p_cap are dropped (the caller learns about the overflow through p_total / status).
// The agent row count may live in device memory (n_agt_dev); the scene tables are padded to n_scenes entries with
// empty scenes (off[b] = total), so the scene count itself never has to.
template <bool FILL>
__global__ void __launch_bounds__(PAIR_WARPS * 32)
k_pairs(const float2* __restrict__ agt_ctrs, const float2* __restrict__ ctx_ctrs,
        const int32_t* __restrict__ agt_off, const int32_t* __restrict__ ctx_off, int n_scenes,
        int64_t n_agt_cap, const int32_t* __restrict__ n_agt_dev, float th, int32_t* __restrict__ cnt,
        const int32_t* __restrict__ row_start, const int32_t* __restrict__ hi_off, const int32_t* __restrict__ wi_off,
        int32_t* __restrict__ hi32, int32_t* __restrict__ wi32, int64_t* __restrict__ hi64, int64_t* __restrict__ wi64,
        int64_t p_cap) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * PAIR_WARPS + (threadIdx.x >> 5);
  if (r >= lgcn_devn(n_agt_dev, n_agt_cap)) return;
  const int b = scene_of(agt_off, n_scenes, (int32_t)r);
  const int32_t c0 = ctx_off[b], c1 = ctx_off[b + 1];
  const float2 a = agt_ctrs[r];
  int32_t run = 0;
  int32_t pos = 0, h = 0, w0 = 0;
  if (FILL) {
    pos = row_start[r];
    if (row_start[r + 1] == pos) return;  // nothing to write for this row
    h = (int32_t)r - agt_off[b] + hi_off[b];
    w0 = wi_off[b] - c0;
  }
  // four 32-wide chunks of context centres per iteration: the loads are independent, so a row with ~1.5 k context
  // rows (M2A) pays ~12 dependent memory round trips instead of ~48
  for (int32_t j0 = c0; j0 < c1; j0 += 128) {
    float2 c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int32_t j = j0 + 32 * q + lane;
      c[q] = j < c1 ? ctx_ctrs[j] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int32_t j = j0 + 32 * q + lane;
      const bool p = j < c1 && within(a.x, a.y, c[q].x, c[q].y, th);
      const unsigned m = __ballot_sync(0xffffffffu, p);
      if (FILL && p) {
        const int32_t o = pos + run + __popc(m & ((1u << lane) - 1u));
        if (o < p_cap) {
          if (hi32) hi32[o] = h;
          if (wi32) wi32[o] = j + w0;
          if (hi64) hi64[o] = h;
          if (wi64) wi64[o] = j + w0;
        }
      }
      run += __popc(m);
    }
  }
  if (!FILL && lane == 0) cnt[r] = run;
}

// Same contract, ONE THREAD per agent row: for lists whose scenes have few context rows (A2M: ~1.5 k lane nodes x 20
// actors per scene, 193 k agent rows per batch) a warp per row leaves most lanes idle and pays a dozen dependent
// loads per warp; here a warp covers 32 consecutive rows (nearly always of one scene, so the scene lookup and the
// context centres are warp-uniform, broadcast loads) and a thread writes its row's few pairs itself, in context order.
template <bool FILL>
__global__ void __launch_bounds__(kIndexThreads)
k_pairs_thin(const float2* __restrict__ agt_ctrs, const float2* __restrict__ ctx_ctrs,
             const int32_t* __restrict__ agt_off, const int32_t* __restrict__ ctx_off, int n_scenes,
             int64_t n_agt_cap, const int32_t* __restrict__ n_agt_dev, float th, int32_t* __restrict__ cnt,
             const int32_t* __restrict__ row_start, const int32_t* __restrict__ hi_off, const int32_t* __restrict__ wi_off,
             int32_t* __restrict__ hi32, int32_t* __restrict__ wi32, int64_t* __restrict__ hi64, int64_t* __restrict__ wi64,
             int64_t p_cap) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= lgcn_devn(n_agt_dev, n_agt_cap)) return;
  int32_t pos = 0;
  if (FILL) {
    pos = row_start[r];
    if (row_start[r + 1] == pos) return;  // nothing to write for this row
  }
  const int b = scene_of(agt_off, n_scenes, (int32_t)r);
  const int32_t c0 = ctx_off[b], c1 = ctx_off[b + 1];
  const float2 a = agt_ctrs[r];
  int32_t h = 0, w0 = 0;
  if (FILL) {
    h = (int32_t)r - agt_off[b] + hi_off[b];
    w0 = wi_off[b] - c0;
  }
  int32_t run = 0;
  for (int32_t j = c0; j < c1; ++j) {
    const float2 c = __ldg(ctx_ctrs + j);
    if (within(a.x, a.y, c.x, c.y, th)) {
      if (FILL) {
        const int32_t o = pos + run;
        if (o < p_cap) {
          if (hi32) hi32[o] = h;
          if (wi32) wi32[o] = j + w0;
          if (hi64) hi64[o] = h;
          if (wi64) wi64[o] = j + w0;
        }
      }
      ++run;
    }
  }
  if (!FILL) cnt[r] = run;
}

// Per-scene totals from the scanned counts, then the reference's offset bookkeeping (lanegcn.py:681-687: a scene
// without pairs `continue`s BEFORE hi_count/wi_count advance).  One CTA: thread b owns scene b, the two running
// offsets are block-wide exclusive scans of the sizes of the scenes that have pairs.
// ws layout (int32): row_start[n_agt+1] | hi_off[B] | wi_off[B] | cnt[n_agt] | scan scratch[1088]
__global__ void __launch_bounds__(1024)
k_pairs_scene_offsets(const int32_t* __restrict__ row_start, const int32_t* __restrict__ agt_off,
                      const int32_t* __restrict__ ctx_off, int n_scenes, int keep_quirk,
                      int32_t* __restrict__ hi_off, int32_t* __restrict__ wi_off, int32_t* __restrict__ used_rows) {
  __shared__ int32_t wt[32];
  __shared__ int32_t carry[2];
  if (threadIdx.x == 0) carry[0] = carry[1] = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_scenes; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    int32_t na = 0, nc = 0;
    if (b < n_scenes) {
      const int32_t tot = row_start[agt_off[b + 1]] - row_start[agt_off[b]];
      if (!keep_quirk || tot > 0) {
        na = agt_off[b + 1] - agt_off[b];
        nc = ctx_off[b + 1] - ctx_off[b];
      }
    }
    // two block-wide inclusive scans (warp shuffles + one smem hop)
    int32_t inc[2] = {na, nc};
    for (int k = 0; k < 2; ++k) {
      int32_t v = inc[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += u;
      }
      if ((threadIdx.x & 31) == 31) wt[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x < 32) {
        int32_t x = wt[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int32_t u = __shfl_up_sync(0xffffffffu, x, o);
          if (threadIdx.x >= o) x += u;
        }
        wt[threadIdx.x] = x;
      }
      __syncthreads();
      inc[k] = v + ((threadIdx.x >> 5) ? wt[(threadIdx.x >> 5) - 1] : 0) + carry[k];
      __syncthreads();
    }
    if (b < n_scenes) {
      if (keep_quirk) {
        hi_off[b] = inc[0] - na;
        wi_off[b] = inc[1] - nc;
      } else {
        hi_off[b] = agt_off[b];
        wi_off[b] = ctx_off[b];
      }
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) {
      carry[0] = inc[0];
      carry[1] = inc[1];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *used_rows = keep_quirk ? carry[0] : agt_off[n_scenes];
}

// destination-indexed rowptr: d = r - agt_off[b] + hi_off[b] for rows of scenes that have pairs (monotone,
// injective); rows past the last used destination point at P.  With a pair capacity (p_cap >= 0) every entry is
// clamped to it, *p_total receives min(P, p_cap), and status[0] gets `overflow_bit` if P > p_cap or `empty_bit` if
// P == 0 (the reference raises there, lanegcn.py:688) — the caller reads status when it next synchronises.
__global__ void k_pairs_rowptr_dst(const int32_t* __restrict__ row_start, const int32_t* __restrict__ agt_off,
                                   int n_scenes, int64_t n_agt_cap, const int32_t* __restrict__ n_agt_dev,
                                   int keep_quirk, const int32_t* __restrict__ hi_off,
                                   const int32_t* __restrict__ used_rows, int32_t* __restrict__ rowptr_dst,
                                   int64_t p_cap, int32_t* __restrict__ p_total, int32_t* __restrict__ status,
                                   int32_t* __restrict__ p_exact, int overflow_bit, int empty_bit) {
  const int64_t n_agt = lgcn_devn(n_agt_dev, n_agt_cap);
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_agt) return;
  const int32_t P = row_start[n_agt];
  const int32_t lim = (p_cap >= 0 && P > p_cap) ? (int32_t)p_cap : P;
  if (r >= *used_rows) rowptr_dst[r] = lim;  // includes r == n_agt
  if (r < n_agt) {
    const int b = scene_of(agt_off, n_scenes, (int32_t)r);
    const int32_t tot = row_start[agt_off[b + 1]] - row_start[agt_off[b]];
    if (!keep_quirk || tot > 0) rowptr_dst[r - agt_off[b] + hi_off[b]] = min(row_start[r], lim);
  }
  if (r == n_agt) {
    if (p_total) *p_total = lim;
    if (p_exact) *p_exact = P;
    if (status) {
      if (P > lim) atomicOr(status, overflow_bit);
      if (P == 0) atomicOr(status, empty_bit);
    }
  }
}

static inline int64_t pairs_ws_ints(int64_t n_agt, int n_scenes) {
  return lgcn_align_up(n_agt + 1, 64) + 2 * lgcn_align_up(n_scenes, 64) + lgcn_align_up(n_agt, 64) + 1088;
}

extern "C" int64_t lgcn_pairs_workspace_bytes(int64_t n_agt, int n_scenes) {
  return 4 * pairs_ws_ints(n_agt, n_scenes) + 256;
}

// count (+ fill when hi32 / wi32 / hi64 / wi64 are given and p_cap >= 0): the whole pair list without a host round
// trip.  *status: flag bits (OR-ed in); *p_total = min(P, p_cap) drives the consumers; *p_exact = P.
int lgcn_launch_pairs(const float* agt_ctrs, const float* ctx_ctrs, const int32_t* agt_off, const int32_t* ctx_off,
                      int n_scenes, int64_t n_agt_cap, const int32_t* n_agt_dev, float th, int keep_quirk,
                      int32_t* rowptr, void* workspace, int64_t p_cap, int32_t* hi32, int32_t* wi32, int32_t* p_total,
                      int32_t* status, int32_t* p_exact, int overflow_bit, int empty_bit, int64_t n_ctx_hint, cudaStream_t st) {
  LGCN_CHECK_ARG(n_scenes >= 1 && n_agt_cap >= 0, "pairs: n_scenes %d n_agt %lld", n_scenes, (long long)n_agt_cap);
  LGCN_CHECK_ARG(n_agt_cap < (int64_t)1 << 31, "pairs: n_agt exceeds int32");
  int32_t* row_start = (int32_t*)workspace;
  int32_t* hi_off = row_start + lgcn_align_up(n_agt_cap + 1, 64);
  int32_t* wi_off = hi_off + lgcn_align_up(n_scenes, 64);
  int32_t* cnt = wi_off + lgcn_align_up(n_scenes, 64);
  // scenes with few context rows (n_ctx_hint = their total or its capacity; 0 = unknown): one thread per agent row
  const bool thin = n_ctx_hint > 0 && n_ctx_hint <= (int64_t)48 * n_scenes;
  const unsigned grid = thin ? lgcn_cdiv(n_agt_cap, kIndexThreads) : lgcn_cdiv(n_agt_cap, PAIR_WARPS);
  if (n_agt_cap > 0) {
    if (thin)
      k_pairs_thin<false><<<grid, kIndexThreads, 0, st>>>((const float2*)agt_ctrs, (const float2*)ctx_ctrs, agt_off, ctx_off, n_scenes,
                                                n_agt_cap, n_agt_dev, th, cnt, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                nullptr, nullptr, 0);
    else
      k_pairs<false><<<grid, PAIR_WARPS * 32, 0, st>>>((const float2*)agt_ctrs, (const float2*)ctx_ctrs, agt_off, ctx_off,
                                                       n_scenes, n_agt_cap, n_agt_dev, th, cnt, nullptr, nullptr, nullptr,
                                                       nullptr, nullptr, nullptr, nullptr, 0);
    LGCN_LAUNCH_OK();
  }
  int32_t* scratch = cnt + lgcn_align_up(n_agt_cap, 64);  // [0..1024] scan scratch, [1056] used_rows
  if (lgcn_launch_exclusive_scan(cnt, row_start, n_agt_cap, n_agt_dev, scratch, st)) return -2;
  k_pairs_scene_offsets<<<1, 1024, 0, st>>>(row_start, agt_off, ctx_off, n_scenes, keep_quirk, hi_off, wi_off,
                                            scratch + 1056);
  LGCN_LAUNCH_OK();
  k_pairs_rowptr_dst<<<lgcn_cdiv(n_agt_cap + 1, 256), 256, 0, st>>>(row_start, agt_off, n_scenes, n_agt_cap, n_agt_dev,
                                                                   keep_quirk, hi_off, scratch + 1056, rowptr, p_cap,
                                                                   p_total, status, p_exact, overflow_bit, empty_bit);
  LGCN_LAUNCH_OK();
  if (p_cap >= 0 && n_agt_cap > 0 && (hi32 || wi32)) {
    if (thin)
      k_pairs_thin<true><<<grid, kIndexThreads, 0, st>>>((const float2*)agt_ctrs, (const float2*)ctx_ctrs, agt_off, ctx_off, n_scenes,
                                               n_agt_cap, n_agt_dev, th, nullptr, row_start, hi_off, wi_off, hi32, wi32,
                                               nullptr, nullptr, p_cap);
    else
      k_pairs<true><<<grid, PAIR_WARPS * 32, 0, st>>>((const float2*)agt_ctrs, (const float2*)ctx_ctrs, agt_off, ctx_off,
                                                      n_scenes, n_agt_cap, n_agt_dev, th, nullptr, row_start, hi_off, wi_off,
                                                      hi32, wi32, nullptr, nullptr, p_cap);
    LGCN_LAUNCH_OK();
  }
  return 0;
}

extern "C" int lgcn_pairs_count(const float* agt_ctrs, const float* ctx_ctrs, const int32_t* agt_off,
                                const int32_t* ctx_off, int n_scenes, int64_t n_agt, float th, int keep_quirk,
                                int32_t* rowptr, void* workspace, int64_t* h_total, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = lgcn_launch_pairs(agt_ctrs, ctx_ctrs, agt_off, ctx_off, n_scenes, n_agt, nullptr, th, keep_quirk, rowptr,
                                 workspace, -1, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, st))
    return rc;
  if (h_total) {
    int32_t p = 0;
    LGCN_CUDA_OK(cudaMemcpyAsync(&p, (int32_t*)workspace + n_agt, 4, cudaMemcpyDeviceToHost, st));
    LGCN_CUDA_OK(cudaStreamSynchronize(st));
    *h_total = p;
  }
  return 0;
}

extern "C" int lgcn_pairs_fill(const float* agt_ctrs, const float* ctx_ctrs, const int32_t* agt_off,
                               const int32_t* ctx_off, int n_scenes, int64_t n_agt, float th,
                               const void* workspace, int32_t* hi32, int32_t* wi32, int64_t* hi64,
                               int64_t* wi64, void* stream) {
  if (n_agt <= 0) return 0;
  const int32_t* row_start = (const int32_t*)workspace;
  const int32_t* hi_off = row_start + lgcn_align_up(n_agt + 1, 64);
  const int32_t* wi_off = hi_off + lgcn_align_up(n_scenes, 64);
  k_pairs<true><<<lgcn_cdiv(n_agt, PAIR_WARPS), PAIR_WARPS * 32, 0, (cudaStream_t)stream>>>(
      (const float2*)agt_ctrs, (const float2*)ctx_ctrs, agt_off, ctx_off, n_scenes, n_agt, nullptr, th, nullptr,
      row_start, hi_off, wi_off, hi32, wi32, hi64, wi64, (int64_t)1 << 40);
  LGCN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ nn.Linear(2,128) + bias + ReLU heads
// One warp per output row (512 B coalesced store); x = p[ip[m]] - q[iq[m]] for Att.dist (lanegcn.py:693).
__global__ void __launch_bounds__(256)
k_mlp2_in(const float2* __restrict__ p, const int32_t* __restrict__ ip, const float2* __restrict__ q,
          const int32_t* __restrict__ iq, const float* __restrict__ W1, const float* __restrict__ b1,
          float* __restrict__ h, int64_t m_cap, const int32_t* __restrict__ m_dev) {
  const int lane = threadIdx.x & 31;
  lgcn_pdl_trigger();
  const int64_t m = lgcn_devn(m_dev, m_cap);
  // W1 is [128,2] row-major: this lane's 4 output channels are rows lane*4..lane*4+3 = 8 contiguous floats
  const float4 w01 = reinterpret_cast<const float4*>(W1)[lane * 2];
  const float4 w23 = reinterpret_cast<const float4*>(W1)[lane * 2 + 1];
  const float4 bb = reinterpret_cast<const float4*>(b1)[lane];
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < m; r += (int64_t)gridDim.x * 8) {
    float2 x = p[ip ? ip[r] : r];
    if (q) {
      const float2 y = q[iq ? iq[r] : r];
      x.x -= y.x;
      x.y -= y.y;
    }
    float4 o;
    // same association as addmm(bias, x, W^T): (x0*w0 + x1*w1) + b  — fp32, no reordering across terms
    o.x = fmaxf(fmaf(x.y, w01.y, x.x * w01.x) + bb.x, 0.f);
    o.y = fmaxf(fmaf(x.y, w01.w, x.x * w01.z) + bb.y, 0.f);
    o.z = fmaxf(fmaf(x.y, w23.y, x.x * w23.x) + bb.z, 0.f);
    o.w = fmaxf(fmaf(x.y, w23.w, x.x * w23.z) + bb.w, 0.f);
    reinterpret_cast<float4*>(h + r * LGCN_C)[lane] = o;
  }
}

// nn.Linear(4,128) + bias + ReLU on x = p[ip[m]] - q[iq[m]] (4 floats per row): LanePooling.relpose, lanercnn.py:443-446
__global__ void __launch_bounds__(256)
k_mlp4_in(const float4* __restrict__ p, const int32_t* __restrict__ ip, const float4* __restrict__ q,
          const int32_t* __restrict__ iq, const float* __restrict__ W1, const float* __restrict__ b1,
          float* __restrict__ h, int64_t m) {
  const int lane = threadIdx.x & 31;
  float4 w[4];  // rows lane*4 .. lane*4+3 of W1 [128,4]
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = reinterpret_cast<const float4*>(W1)[lane * 4 + j];
  const float4 bb = reinterpret_cast<const float4*>(b1)[lane];
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < m; r += (int64_t)gridDim.x * 8) {
    float4 x = p[ip ? ip[r] : r];
    if (q) {
      const float4 y = q[iq ? iq[r] : r];
      x.x -= y.x; x.y -= y.y; x.z -= y.z; x.w -= y.w;
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = fmaf(x.w, w[j].w, fmaf(x.z, w[j].z, fmaf(x.y, w[j].y, x.x * w[j].x)));
    reinterpret_cast<float4*>(h + r * LGCN_C)[lane] =
        make_float4(fmaxf(o[0] + bb.x, 0.f), fmaxf(o[1] + bb.y, 0.f), fmaxf(o[2] + bb.z, 0.f), fmaxf(o[3] + bb.w, 0.f));
  }
}

extern "C" int lgcn_mlp4_in(const float* p, const int32_t* ip, const float* q, const int32_t* iq, const float* W1,
                            const float* b1, float* h, int64_t m, void* stream) {
  if (m <= 0) return 0;
  const unsigned grid = min(lgcn_cdiv(m, 8), 148u * 32u);
  k_mlp4_in<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)p, ip, (const float4*)q, iq, W1, b1, h, m);
  LGCN_LAUNCH_OK();
  return 0;
}

int lgcn_launch_mlp2_in(const float* p, const int32_t* ip, const float* q, const int32_t* iq, const float* W1,
                        const float* b1, float* h, int64_t m_cap, const int32_t* m_dev, cudaStream_t st) {
  if (m_cap <= 0) return 0;
  const unsigned grid = min(lgcn_cdiv(m_cap, 8), 148u * 32u);
  k_mlp2_in<<<grid, 256, 0, st>>>((const float2*)p, ip, (const float2*)q, iq, W1, b1, h, m_cap, m_dev);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_mlp2_in(const float* p, const int32_t* ip, const float* q, const int32_t* iq,
                            const float* W1, const float* b1, float* h, int64_t m, void* stream) {
  return lgcn_launch_mlp2_in(p, ip, q, iq, W1, b1, h, m, nullptr, (cudaStream_t)stream);
}
