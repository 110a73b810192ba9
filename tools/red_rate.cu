// red_rate.cu — development probe: ceiling of fp32 vector reductions into a [N, 256] table on B200
// (the feature-backward epilogue pattern: 1 KB rows, a row is hit by ~3 different CTAs).
//   mode 0: red.global.add.v4.f32, a warp covers 4 rows x 128 B per instruction
//   mode 1: cp.reduce.async.bulk .add.f32, one 1 KB op per row from shared memory
//   mode 2: same as 1 with 512 B ops
//   mode 3: scalar red.global.add.f32, a warp covers 128 contiguous bytes per instruction
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/red_rate tools/red_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void __launch_bounds__(256) k_red(float *table, int nrows, int rows_per_cta, int mode, int spread, int share) {
  extern __shared__ __align__(128) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 8 * 256; i += 256) sm[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // rows: a CTA hits rows_per_cta pseudo-random rows inside a window (locality like a screen tile)
  const uint32_t base = hash(blockIdx.x / 3) % (uint32_t)(nrows - spread);   // 3 CTAs share a window
  if (mode == 3) {
    // scalar reds, a warp covers 128 contiguous bytes of one row per instruction (8 per 1 KB row)
    for (int r = warp; r < rows_per_cta; r += 8) {
      const uint32_t row = base + hash((share ? blockIdx.x / 3 : blockIdx.x) * 7919u + r) % (uint32_t)spread;
      float *dst = table + (size_t)row * 256;
#pragma unroll
      for (int h = 0; h < 8; ++h)
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + h * 32 + lane), "f"(1.0f) : "memory");
    }
  } else if (mode == 0) {
    for (int r = warp; r < rows_per_cta; r += 8) {
      const uint32_t row = base + hash((share ? blockIdx.x / 3 : blockIdx.x) * 7919u + r) % (uint32_t)spread;
      float *dst = table + (size_t)row * 256;
      // 2 instructions x 32 lanes x 16 B = 1 KB row
#pragma unroll
      for (int h = 0; h < 2; ++h)
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(dst + h * 128 + lane * 4), "f"(1.0f) : "memory");
    }
  } else {
    const uint32_t bytes = mode == 1 ? 1024u : 512u;
    const int ops_per_row = mode == 1 ? 1 : 2;
    if (tid < 32) {
      for (int r = lane; r < rows_per_cta * ops_per_row; r += 32) {
        const int rr = r / ops_per_row, part = r % ops_per_row;
        const uint32_t row = base + hash((share ? blockIdx.x / 3 : blockIdx.x) * 7919u + rr) % (uint32_t)spread;
        float *dst = table + (size_t)row * 256 + part * 128;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                     ::"l"(dst), "r"(smem_u32(sm + (lane & 7) * 256 + part * 128)), "r"(bytes) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
}

int main() {
  const int nrows = 2000000;
  float *table;
  CK(cudaMalloc(&table, (size_t)nrows * 1024));
  CK(cudaMemset(table, 0, (size_t)nrows * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int ctas = 16200, rows = 192;           // 3.1 M row reductions = 3.2 GB, like one config-3 view
  for (int share = 0; share < 2; ++share)
  for (int mode = 0; mode < 4; ++mode)
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemsetAsync(table, 0, (size_t)nrows * 1024));
      CK(cudaEventRecord(e0));
      k_red<<<ctas, 256, 8 * 1024>>>(table, nrows, rows, mode, 4000, share);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double gb = (double)ctas * rows * 1024 / 1e9;
      printf("same-rows-in-3-CTAs %d mode %d rep %d: %.3f ms for %.2f GB of reductions -> %.0f GB/s (%.1f B/cycle/SM at 1.965 GHz)\n",
             share, mode, rep, ms, gb, gb / ms * 1e3, gb * 1e9 / (ms * 1e-3) / 148 / 1.965e9);
    }
  printf("red probe done\n");
  return 0;
}
