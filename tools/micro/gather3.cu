// Micro-benchmark 3: cp.async.cg row gather with wider contiguous pieces per row (compile-time shapes, 8 warps).
// Per "unit" 128 rows x 512 B = 64 KB are fetched (one key of a LaneConv tile), as 512 / ROW_BYTES stages of
// 128 rows x ROW_BYTES, ROW_BYTES / 16 lanes per row.  Reports cycles per 64 KB unit (MMA time per key: 3072 cycles).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather3 gather3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ROW_BYTES, int PATTERN, bool CA>
__global__ void __launch_bounds__(384, 1) k_gather(const float* __restrict__ X, int64_t n_rows, int units, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  constexpr int LPR = ROW_BYTES / 16;          // lanes per row
  constexpr int RPI = 32 / LPR;                // rows per warp instruction
  constexpr int INSTR = 16 / RPI;              // instructions per warp and stage (16 rows per warp)
  constexpr int STAGES = 512 / ROW_BYTES;      // stages per unit
  constexpr int SLOTS = 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  const long long t0 = clock64();
  if (warp >= 4) {
    const int e = warp - 4;
    uint32_t rng = blockIdx.x * 9781u + threadIdx.x * 7919u + 12345u;
    const int64_t base = (int64_t)blockIdx.x * (n_rows / gridDim.x);
    int slot = 0;
    for (int u = 0; u < units; ++u) {
      int64_t rows[INSTR];
#pragma unroll
      for (int j = 0; j < INSTR; ++j) {
        const int rloc = e * 16 + RPI * j + lane / LPR;
        if (PATTERN == 0) {
          rng = rng * 1664525u + 1013904223u;
          const uint32_t rr = __shfl_sync(0xffffffffu, rng, lane - lane % LPR);
          rows[j] = (base + (rr >> 8) % 4096) % n_rows;
        } else {
          rows[j] = (base + (u % 8) * 128 + rloc + ((u & 1) ? 1 : -1) + n_rows) % n_rows;
        }
      }
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
#pragma unroll
        for (int j = 0; j < INSTR; ++j) {
          const int rloc = e * 16 + RPI * j + lane / LPR, po = (lane % LPR) * 16;
          const char* p = (const char*)(X + rows[j] * 128) + s * ROW_BYTES + po;
          const uint32_t dst = sbase + slot * (128 * ROW_BYTES) + rloc * ROW_BYTES + po;
          if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
          else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(SLOTS - 1) : "memory");
        if (++slot == SLOTS) slot = 0;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

template <int ROW_BYTES, int PATTERN, bool CA>
void run(const float* X, int64_t n_rows, long long* out) {
  const int units = 1000;
  const int smem = 3 * 128 * ROW_BYTES;
  cudaFuncSetAttribute(k_gather<ROW_BYTES, PATTERN, CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    k_gather<ROW_BYTES, PATTERN, CA><<<148, 384, smem>>>(X, n_rows, units, out);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return; }
  }
  long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
  printf("pattern %s %s row piece %3d B: %7.0f cycles per 64 KB unit (one key of a tile)  %5.1f B/clk/SM\n", PATTERN ? "consecutive" : "random     ",
         CA ? "ca" : "cg", ROW_BYTES, cyc / units, 65536.0 * units / cyc);
}

int main() {
  const int64_t n_rows = 193536;
  float* X; long long* out;
  cudaMalloc(&X, n_rows * 512); cudaMalloc(&out, 148 * 8);
  cudaMemset(X, 0, n_rows * 512);
  run<64, 0, false>(X, n_rows, out);
  run<128, 0, false>(X, n_rows, out);
  run<256, 0, false>(X, n_rows, out);
  run<512, 0, false>(X, n_rows, out);
  run<64, 1, false>(X, n_rows, out);
  run<128, 1, false>(X, n_rows, out);
  run<256, 1, false>(X, n_rows, out);
  run<512, 1, false>(X, n_rows, out);
  run<128, 1, true>(X, n_rows, out);
  run<512, 1, true>(X, n_rows, out);
  return 0;
}
