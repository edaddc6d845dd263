// Micro-benchmark: row-gather feed of k_laneconv_fused in isolation.  8 warps fetch 128 rows x 128 B (one 32-float
// K-chunk of 128 source rows = 16 KB) per stage into shared memory with cp.async, `depth` stages in flight.
//   pattern 0: random rows in a 4096-row window   1: consecutive rows (tile rows t*128 + r +- 1: the lane-graph case)
//   lanes   4: 4 lanes per 64 B slice, two warps (h = 0, 1) split each row's 128 B (the kernel's layout)
//           8: 8 lanes per 128 B, each warp fetches whole 128 B lines of its 16 rows
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int DEPTH>
__global__ void __launch_bounds__(384, 1) k_gather(const float* __restrict__ X, int64_t n_rows, int iters, int pattern, int lanes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  const long long t0 = clock64();
  if (warp >= 4) {
    const int e = warp - 4;
    uint32_t rng = blockIdx.x * 9781u + threadIdx.x * 7919u + 12345u;
    const int64_t base = (int64_t)blockIdx.x * (n_rows / gridDim.x);
    for (int i = 0; i < iters; ++i) {
      const int kc = i & 3;
      const int64_t tile = base + ((i >> 2) % 8) * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int rloc, piece_off;   // row of the 128-row tile, byte offset inside the 128 B chunk
        if (lanes == 4) { rloc = (e & 3) * 32 + 8 * j + (lane >> 2); piece_off = (e >> 2) * 64 + (lane & 3) * 16; }
        else { rloc = e * 16 + 4 * j + (lane >> 3); piece_off = (lane & 7) * 16; }
        int64_t row;
        if (pattern == 0) {
          rng = rng * 1664525u + 1013904223u;
          const uint32_t rr = __shfl_sync(0xffffffffu, rng, lanes == 4 ? (lane & ~3) : (lane & ~7));
          row = (base + (rr >> 8) % 4096) % n_rows;
        } else {
          row = (tile + rloc + ((i >> 2) & 1 ? 1 : -1) + n_rows) % n_rows;
        }
        const char* p = (const char*)(X + row * 128 + kc * 32) + piece_off;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (i % DEPTH) * 16384 + rloc * 128 + piece_off), "l"(p) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

template <int DEPTH>
void run(const float* X, int64_t n_rows, long long* out, int pattern, int lanes) {
  const int iters = 4000;
  cudaFuncSetAttribute(k_gather<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * 16384);
  for (int rep = 0; rep < 2; ++rep) {
    k_gather<DEPTH><<<148, 384, DEPTH * 16384>>>(X, n_rows, iters, pattern, lanes, out);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return; }
  }
  long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
  printf("pattern %s  %d lanes/row-piece  depth %2d: %7.0f cycles per 16 KB stage  (%5.1f B/clk/SM)\n", pattern ? "consecutive" : "random     ", lanes, DEPTH,
         cyc / iters, 16384.0 * iters / cyc);
}

int main() {
  const int64_t n_rows = 193536;
  float* X; long long* out;
  cudaMalloc(&X, n_rows * 512); cudaMalloc(&out, 148 * 8);
  cudaMemset(X, 0, n_rows * 512);
  for (int pattern = 0; pattern < 2; ++pattern)
    for (int lanes = 4; lanes <= 8; lanes += 4) {
      run<3>(X, n_rows, out, pattern, lanes);
      run<6>(X, n_rows, out, pattern, lanes);
      run<10>(X, n_rows, out, pattern, lanes);
    }
  return 0;
}
