// Micro-benchmark 2: other ways to gather 128 rows x `row_bytes` per stage into shared memory (all 148 SMs):
//   mode 0  cp.async.cg 16 B (reference: tools/micro/gather.cu)
//   mode 1  cp.async.ca 16 B (L1-allocating)
//   mode 2  ld.global.nc.v4 + st.shared (8 lanes per 128 B)
//   mode 3  cp.async.bulk (one bulk copy of row_bytes per row, one lane per row, completes on an mbarrier)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather2 gather2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

constexpr int DEPTH = 4;

__global__ void __launch_bounds__(384, 1) k_gather(const float* __restrict__ X, int64_t n_rows, int iters, int pattern, int mode, int row_bytes,
                                                    long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const int stage_bytes = 128 * row_bytes;
  const uint32_t bars = sbase + DEPTH * stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  uint32_t rng = blockIdx.x * 9781u + threadIdx.x * 7919u + 12345u;
  const int64_t base = (int64_t)blockIdx.x * (n_rows / gridDim.x);
  auto row_of = [&](int i, int rloc, int group_lane) -> int64_t {
    if (pattern == 0) {
      rng = rng * 1664525u + 1013904223u;
      const uint32_t rr = __shfl_sync(0xffffffffu, rng, group_lane);
      return (base + (rr >> 8) % 4096) % n_rows;
    }
    const int64_t tile = base + ((i >> 2) % 8) * 128;
    return (tile + rloc + ((i >> 2) & 1 ? 1 : -1) + n_rows) % n_rows;
  };
  if (mode <= 2 && warp >= 4) {
    const int e = warp - 4;
    const int pieces = row_bytes / 16;                 // 16 B pieces per row (8 for 128 B)
    const int rows_per_instr = 32 / pieces;
    const int instrs = 16 / rows_per_instr;            // this warp's 16 rows
    for (int i = 0; i < iters; ++i) {
      const int kc = i & 3;
      for (int j = 0; j < instrs; ++j) {
        const int rloc = e * 16 + rows_per_instr * j + lane / pieces, po = (lane % pieces) * 16;
        const int64_t row = row_of(i, rloc, lane - lane % pieces);
        const char* p = (const char*)(X + row * 128) + (row_bytes == 512 ? 0 : kc * row_bytes % 512) + po;
        const uint32_t dst = sbase + (i % DEPTH) * stage_bytes + rloc * row_bytes + po;
        if (mode == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
        else if (mode == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
        else {
          float4 v;
          asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
      if (mode < 2) {
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
      }
    }
    if (mode < 2) asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (mode == 3 && warp >= 4 && warp < 8) {
    // warps 4-7: lane = one row; stage i completes on bars[i % DEPTH] (expect_tx by warp 4 lane 0 up front)
    const int rloc = (warp - 4) * 32 + lane;
    for (int i = 0; i < iters; ++i) {
      const int s = i % DEPTH, kc = i & 3;
      if (i >= DEPTH) mbar_wait(bars + 8 * s, ((i / DEPTH) - 1) & 1);
      if (warp == 4 && lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * s), "r"((uint32_t)stage_bytes) : "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");   // expect_tx before any complete_tx of this phase
      const int64_t row = row_of(i, rloc, lane);
      const char* p = (const char*)(X + row * 128) + (row_bytes == 512 ? 0 : kc * row_bytes % 512);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + s * stage_bytes + rloc * row_bytes),
                   "l"(p), "r"((uint32_t)row_bytes), "r"(bars + 8 * s)
                   : "memory");
    }
    for (int i = iters; i < iters + DEPTH; ++i) mbar_wait(bars + 8 * (i % DEPTH), ((i / DEPTH) - 1) & 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const int64_t n_rows = 193536;
  float* X; long long* out;
  cudaMalloc(&X, n_rows * 512); cudaMalloc(&out, 148 * 8);
  cudaMemset(X, 0, n_rows * 512);
  cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"cp.async.cg 16 B", "cp.async.ca 16 B", "ld.global.nc.v4 + st.shared", "cp.async.bulk per row"};
  const int iters = 3000;
  for (int pattern = 0; pattern < 2; ++pattern)
    for (int mode = 0; mode < 4; ++mode)
      for (int row_bytes = 128; row_bytes <= (mode == 3 ? 512 : 128); row_bytes *= 2) {
        if (DEPTH * 128 * row_bytes + 64 > 200 * 1024) continue;
        for (int rep = 0; rep < 2; ++rep) {
          k_gather<<<148, 384, DEPTH * 128 * row_bytes + 64>>>(X, n_rows, iters, pattern, mode, row_bytes, out);
          cudaError_t err = cudaDeviceSynchronize();
          if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        }
        long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
        printf("pattern %s  %-28s row_bytes %3d: %7.0f cycles per 128-row stage  (%5.1f B/clk/SM)\n", pattern ? "consecutive" : "random     ", names[mode],
               row_bytes, cyc / iters, 128.0 * row_bytes * iters / cyc);
      }
  return 0;
}
