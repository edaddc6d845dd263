// Micro-benchmark 4: how many warps does the cp.async row gather need?  W warps fetch 128 rows x 128 B (16 KB) per
// stage, 8 lanes per row, 128/W rows per warp (= 32/W cp.async per lane and stage), 3 stages in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4 gather4.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int W>
__global__ void __launch_bounds__(384, 1) k_gather(const float* __restrict__ X, int64_t n_rows, int units, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  constexpr int INSTR = 32 / W;   // instructions per warp and stage (4 rows each)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < W) {
    const int64_t base = (int64_t)blockIdx.x * (n_rows / gridDim.x);
    int slot = 0;
    for (int u = 0; u < units; ++u) {
      const float* ptr[INSTR];
#pragma unroll
      for (int j = 0; j < INSTR; ++j) {
        const int rloc = warp * (128 / W) + 4 * j + (lane >> 3);
        ptr[j] = X + ((base + (u % 8) * 128 + rloc + ((u & 1) ? 1 : -1) + n_rows) % n_rows) * 128;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int j = 0; j < INSTR; ++j) {
          const int rloc = warp * (128 / W) + 4 * j + (lane >> 3), po = (lane & 7) * 16;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + slot * 16384 + rloc * 128 + po), "l"((const char*)(ptr[j] + s * 32) + po) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        if (++slot == 3) slot = 0;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

template <int W>
void run(const float* X, int64_t n_rows, long long* out) {
  const int units = 1000;
  cudaFuncSetAttribute(k_gather<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  for (int rep = 0; rep < 2; ++rep) {
    k_gather<W><<<148, 384, 49152>>>(X, n_rows, units, out);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return; }
  }
  long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
  printf("%d gather warps: %6.0f cycles per 16 KB stage (%5.1f B/clk/SM), %5.1f cycles per cp.async instruction and warp\n", W, cyc / units / 4,
         65536.0 * units / cyc, cyc / units / 4 / (32 / W));
}

int main() {
  const int64_t n_rows = 193536;
  float* X; long long* out;
  cudaMalloc(&X, n_rows * 512); cudaMalloc(&out, 148 * 8);
  cudaMemset(X, 0, n_rows * 512);
  run<1>(X, n_rows, out);
  run<2>(X, n_rows, out);
  run<4>(X, n_rows, out);
  run<8>(X, n_rows, out);
  return 0;
}
