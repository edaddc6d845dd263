// laneconv.cu — the HBM-bound half of a LaneConv block: the deterministic CSR gather-reduce that replaces
// the 14 index_add_ scatters of lanegcn.py:333-354, with GroupNorm(1)+ReLU (lanegcn.py:356-357) fused in
// the epilogue.  Also the contiguous-segment variant used by Att (lanegcn.py:702-705).
//
// Mapping: one warp per destination row.  A row of any 128-float block is 512 B = 32 lanes x float4, so
// every gathered row is ONE fully coalesced 128-bit load per lane.  The column indices of a row are read
// with one coalesced load (lane e reads col[beg+e]) and broadcast by shuffle; the row loads of a batch of
// up to UNROLL edges are issued before the first add so several 512 B requests per warp are in flight.
// Adds are applied strictly in CSR order (ctr, then keys in order, then edge-list order), which is the
// order CPU index_add_ uses, so the result does not depend on scheduling (deterministic, unlike the atomic
// index_add_ CUDA kernel the reference would run).
//
// Algorithmic bytes per launch: 4*128*(E + 2N) (gathered rows + ctr rows + output) + 4*E (col) + 4*(N+1)
// (rowptr).  HBM-bound; target >= 60 % of the measured copy bandwidth.
#include "common.cuh"

#define GATHER_WARPS 8
#define GATHER_UNROLL 4

template <bool CONTIG>
__global__ void __launch_bounds__(GATHER_WARPS * 32)
k_gather_gn_relu(const float* __restrict__ base, int64_t base_ld, const float* __restrict__ blocks,
                 const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                 int64_t n_rows, const int32_t* __restrict__ n_dev) {
  const int lane = threadIdx.x & 31;
  lgcn_pdl_trigger();
  const int64_t row = (int64_t)blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5);
  if (row >= lgcn_devn(n_dev, n_rows)) return;
  const int32_t beg = rowptr[row], end = rowptr[row + 1];
  float4 acc = ld_stream_f4(base + row * base_ld + lane * 4);
  for (int32_t e0 = beg; e0 < end; e0 += 32) {
    const int32_t n = min(32, end - e0);
    int32_t my = 0;
    if (!CONTIG && lane < n) my = col[e0 + lane];
    for (int32_t j = 0; j < n; j += GATHER_UNROLL) {
      float4 v[GATHER_UNROLL];
#pragma unroll
      for (int q = 0; q < GATHER_UNROLL; ++q) {
        // shuffles are executed by all lanes (n and j are warp-uniform)
        const int32_t c = CONTIG ? (e0 + j + q) : __shfl_sync(0xffffffffu, my, (j + q) & 31);
        if (j + q < n) v[q] = ld_stream_f4(blocks + (int64_t)c * LGCN_C + lane * 4);
      }
#pragma unroll
      for (int q = 0; q < GATHER_UNROLL; ++q) {
        if (j + q < n) {
          acc.x += v[q].x;
          acc.y += v[q].y;
          acc.z += v[q].z;
          acc.w += v[q].w;
        }
      }
    }
  }
  const float4 g = reinterpret_cast<const float4*>(gamma)[lane];
  const float4 b = reinterpret_cast<const float4*>(beta)[lane];
  reinterpret_cast<float4*>(out + row * LGCN_C)[lane] = relu4(warp_gn128(acc, g, b));
}

extern "C" int lgcn_laneconv_gather_gn_relu(const float* Y, int n_blocks, const int32_t* rowptr,
                                            const int32_t* col, const float* gamma, const float* beta,
                                            float* out, int64_t n_nodes, void* stream) {
  LGCN_CHECK_ARG(n_blocks >= 1, "gather: n_blocks %d", n_blocks);
  if (n_nodes <= 0) return 0;
  k_gather_gn_relu<false><<<lgcn_cdiv(n_nodes, GATHER_WARPS), GATHER_WARPS * 32, 0, (cudaStream_t)stream>>>(
      Y, (int64_t)n_blocks * LGCN_C, Y, rowptr, col, gamma, beta, out, n_nodes, nullptr);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_gather_rows_gn_relu(const float* base, int64_t base_ld, const float* blocks, const int32_t* rowptr,
                                        const int32_t* col, const float* gamma, const float* beta, float* out,
                                        int64_t n_rows, void* stream) {
  LGCN_CHECK_ARG(base_ld >= LGCN_C && base_ld % 4 == 0, "gather_rows: bad base_ld %lld", (long long)base_ld);
  if (n_rows <= 0) return 0;
  k_gather_gn_relu<false><<<lgcn_cdiv(n_rows, GATHER_WARPS), GATHER_WARPS * 32, 0, (cudaStream_t)stream>>>(
      base, base_ld, blocks, rowptr, col, gamma, beta, out, n_rows, nullptr);
  LGCN_LAUNCH_OK();
  return 0;
}

int lgcn_launch_segsum_gn_relu(const float* a, const float* c, const int32_t* rowptr, const float* gamma,
                               const float* beta, float* out, int64_t n_cap, const int32_t* n_dev, cudaStream_t st) {
  if (n_cap <= 0) return 0;
  k_gather_gn_relu<true><<<lgcn_cdiv(n_cap, GATHER_WARPS), GATHER_WARPS * 32, 0, st>>>(a, LGCN_C, c, rowptr, nullptr, gamma,
                                                                                      beta, out, n_cap, n_dev);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_segsum_gn_relu(const float* a, const float* c, const int32_t* rowptr, const float* gamma,
                                   const float* beta, float* out, int64_t n_rows, void* stream) {
  return lgcn_launch_segsum_gn_relu(a, c, rowptr, gamma, beta, out, n_rows, nullptr, (cudaStream_t)stream);
}
