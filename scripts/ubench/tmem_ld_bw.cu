// Micro-benchmark: tcgen05.ld throughput per SM (how fast can warps pull fp32 accumulators out of TMEM?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_bw scripts/ubench/tmem_ld_bw.cu && /tmp/tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../paintmind_b200/csrc/pm_common.cuh"
using namespace pm;

template <int X>   // columns per load: 32 or 16
__global__ void k(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (X == 32) { uint32_t r[32]; tmem_ld_x32(base + ((c * 32 + (warp >> 2) * 128) & 511), r); tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j]; }
      else { uint32_t r[16]; tmem_ld_x16(base + ((c * 16 + (warp >> 2) * 128) & 511), r); tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= r[j]; }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* d; uint32_t* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    k<32><<<148, warps * 32>>>(iters, d, s); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double bytes = double(iters) * 4 * warps * 32 * 32 * 4;
    printf("x32 %2d warps: %.1f cycles per (4 loads/warp) round, %.1f B/clk/SM\n", warps, double(h) / iters, bytes / h);
    k<16><<<148, warps * 32>>>(iters, d, s); cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("x16 %2d warps: %.1f cycles per round, %.1f B/clk/SM\n", warps, double(h) / iters, double(iters) * 4 * warps * 32 * 16 * 4 / h);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
