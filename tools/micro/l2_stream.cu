// Micro-benchmark: how fast can every SM stream a small, L2-resident buffer into shared memory with bulk copies
// (the weight stream of k_laneconv_fused: 32 KB per stage from a ~2 MB image), alone and next to a row gather?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream l2_stream.cu && ./l2_stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// mode bit 0: weight stream (one thread, bulk copies of `chunk` bytes, `depth`-deep ring)
// mode bit 1: gather: 8 warps cp.async 16 B pieces of random 128 B rows of X (n_rows rows), 16 KB per "stage"
__global__ void __launch_bounds__(384, 1) k_stream(const uint8_t* __restrict__ w, int64_t w_bytes, int chunk, int depth, int iters,
                                                    const float* __restrict__ X, int64_t n_rows, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bars = sbase + 200 * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0 && (mode & 1)) {
    if (lane == 0) {
      const int64_t n_chunks = w_bytes / chunk;
      int64_t c = (blockIdx.x * 7) % n_chunks;
      for (int i = 0; i < iters + depth; ++i) {
        const int s = i % depth;
        if (i >= depth) mbar_wait(bars + 8 * s, ((i / depth) - 1) & 1);
        if (i < iters) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * s), "r"((uint32_t)chunk) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + s * chunk),
                       "l"(w + c * chunk), "r"((uint32_t)chunk), "r"(bars + 8 * s)
                       : "memory");
          if (++c == n_chunks) c = 0;
        }
      }
    }
  } else if (warp >= 4 && (mode & 2)) {
    // 8 warps: per stage each lane fetches 4 x 16 B (4 lanes per 64 B slice of a row: the kernel's pattern), 3 stages in flight
    const int e = warp - 4;
    const uint32_t xblk = sbase + 128 * 1024 + e * 2048;
    uint32_t rng = blockIdx.x * 9781u + threadIdx.x * 7919u + 12345u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rng = rng * 1664525u + 1013904223u;
        // rows near each other (lane graphs: neighbours are close in memory), shared by the 4 lanes of a slice
        const uint32_t r4 = __shfl_sync(0xffffffffu, rng, lane & ~3);
        const int64_t row = ((int64_t)blockIdx.x * (n_rows / gridDim.x) + (r4 >> 8) % 4096) % n_rows;
        const float* p = X + row * 128 + (i & 3) * 32 + (e >> 2) * 16 + (lane & 3) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xblk + (i % 3) * 16384 + (8 * j + (lane >> 2)) * 64 + (lane & 3) * 16), "l"(p) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 2;" ::: "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const int64_t w_bytes = 2 << 20, n_rows = 193536;
  uint8_t* w;
  float* X;
  long long* out;
  cudaMalloc(&w, w_bytes);
  cudaMalloc(&X, n_rows * 512);
  cudaMalloc(&out, 148 * 8);
  cudaMemset(w, 1, w_bytes);
  cudaMemset(X, 0, n_rows * 512);
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  const int iters = 4000;
  struct Cfg { int mode, chunk, depth; const char* name; } cfgs[] = {
      {1, 32768, 4, "weights 32 KB x4"}, {1, 16384, 8, "weights 16 KB x8"}, {1, 32768, 3, "weights 32 KB x3"}, {1, 16384, 4, "weights 16 KB x4"},
      {2, 32768, 4, "gather only (16 KB/stage)"}, {3, 32768, 4, "weights 32 KB x4 + gather"}, {3, 16384, 4, "weights 16 KB x4 + gather"}};
  for (auto c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k_stream<<<148, 384, 201 * 1024>>>(w, w_bytes, c.chunk, c.depth, iters, X, n_rows, c.mode, out);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
      if (rep) {
        const double wb = (c.mode & 1) ? (double)iters * c.chunk : 0, gb = (c.mode & 2) ? (double)iters * 16384 : 0;
        printf("%-30s %8.3f ms  %9.0f cyc/CTA  weights %6.1f B/clk/SM  gather %6.1f B/clk/SM  total %6.2f TB/s  cycles per stage %.0f\n", c.name, ms, cyc,
               wb / cyc, gb / cyc, (wb + gb) * 148 / (ms * 1e-3) / 1e12, cyc / iters);
      }
    }
  }
  return 0;
}
