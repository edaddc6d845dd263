// gemm_simt.cu — fp32 CUDA-core implementation of lgcn_linear128 (engine 0).
// It is the numerically plain restatement (true fp32 FMA accumulation) that the tcgen05 3xTF32 engine in
// gemm_tc.cu is validated against on the GPU, and the engine used for shapes the tensor path does not
// take.  128x128 output tile per CTA, BK = 16, 256 threads, 8x8 register micro-tile, double-buffered smem.
#include "common.cuh"

#define BM 128
#define BN 128
#define BK 16
#define LDS 132  // padded leading dimension of the k-major smem tiles

struct RowSrc {
  const float* p[2];  // source row pointers of the two rows this thread stages (nullptr => zero row)
};

__device__ __forceinline__ void stage_load(float4 (&ra)[2], float4 (&rb)[2], const float* (&arow)[2],
                                           const float* (&wrow)[2], int k) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    ra[i] = arow[i] ? *reinterpret_cast<const float4*>(arow[i] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    rb[i] = *reinterpret_cast<const float4*>(wrow[i] + k);
  }
}

__global__ void __launch_bounds__(256) k_linear_simt(LinearArgs a) {
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Bs[2][BK][LDS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t M = lgcn_devn(a.m_dev, a.m);   // live rows (the grid is sized by the capacity a.m)
  const int ob = blockIdx.y;
  const int64_t ldw = (int64_t)a.n_src * LGCN_C + a.ks;

  // staging assignment: rows lrow and lrow+64 of the tile, k offset lk
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const float* wrow[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) wrow[i] = a.W + ((int64_t)ob * BN + lrow + 64 * i) * ldw + lk;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int buf = 0;
  for (int s = 0; s < a.n_src; ++s) {
    const float* arow[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t m = m0 + lrow + 64 * i;
      if (m < M) {
        const int64_t r = a.idx[s] ? (int64_t)a.idx[s][m] : m;
        arow[i] = a.a[s] + r * LGCN_C + lk;
      } else {
        arow[i] = nullptr;
      }
    }
    const float* wr[2] = {wrow[0] + s * LGCN_C, wrow[1] + s * LGCN_C};
    float4 ra[2], rb[2];
    stage_load(ra, rb, arow, wr, 0);
    for (int k0 = 0; k0 < LGCN_C; k0 += BK) {
      // registers -> smem (transposed: k-major)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = lrow + 64 * i;
        As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y;
        As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
        Bs[buf][lk + 0][r] = rb[i].x; Bs[buf][lk + 1][r] = rb[i].y;
        Bs[buf][lk + 2][r] = rb[i].z; Bs[buf][lk + 3][r] = rb[i].w;
      }
      __syncthreads();
      if (k0 + BK < LGCN_C) stage_load(ra, rb, arow, wr, k0 + BK);  // prefetch next chunk during the math
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      buf ^= 1;  // the other buffer was last read before the previous __syncthreads
    }
  }

  // ---- epilogue.  Thread owns rows {ty*4+i, 64+ty*4+i} x cols {tx*4+j, 64+tx*4+j}; the 16 threads that
  // share a row are 16 consecutive lanes (half a warp), so row statistics are 4 xor-shuffles.
  const int cbase[2] = {tx * 4, 64 + tx * 4};
  float g[8], bt[8];
  if (a.flags & LGCN_EPI_GN) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        g[h * 4 + j] = a.gamma[cbase[h] + j];
        bt[h * 4 + j] = a.beta[cbase[h] + j];
      }
  }
  float wx[4][8];
  if (a.ks > 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          wx[q][h * 4 + j] =
              q < a.ks ? a.W[((int64_t)ob * BN + cbase[h] + j) * ldw + (int64_t)a.n_src * LGCN_C + q] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    const bool live = m < M;  // dead rows still take part in the shuffles
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[i][j];
    if (a.ks > 0 && live) {
      for (int q = 0; q < a.ks; ++q) {
        const float x = a.xs[m * a.ks + q];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(x, wx[q][j], v[j]);
      }
    }
    if (a.flags & LGCN_EPI_GN) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.0f / 128.0f);
      float q2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] -= mean;
        q2 += v[j] * v[j];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
      const float rstd = 1.0f / sqrtf(q2 * (1.0f / 128.0f) + LGCN_GN_EPS);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] * rstd * g[j] + bt[j];
    }
    if (!live) continue;
    if (a.flags & LGCN_EPI_RELU1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 o = make_float4(v[h * 4], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
      if (a.flags & LGCN_EPI_RES) {
        const float4 r = *reinterpret_cast<const float4*>(a.res + m * LGCN_C + cbase[h]);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (a.flags & LGCN_EPI_RELU2) o = relu4(o);
      *reinterpret_cast<float4*>(a.out + m * a.ldo + (int64_t)ob * BN + cbase[h]) = o;
    }
  }
}

int lgcn_launch_linear_simt(const LinearArgs& a, cudaStream_t st) {
  if (a.m <= 0) return 0;
  dim3 grid(lgcn_cdiv(a.m, BM), a.n_out_blocks);
  k_linear_simt<<<grid, 256, 0, st>>>(a);
  LGCN_LAUNCH_OK();
  return 0;
}
