// TMEM -> register read bandwidth microbenchmark (sm_100a): how many bytes/clk/SM does tcgen05.ld deliver?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../paintmind_b200/csrc/pm_common.cuh"
using namespace pm;

template <int WARPS, int MODE>   // MODE 0: x32 + wait each; 1: 4 x x32 then wait; 2: x16 + wait each
__global__ void __launch_bounds__(WARPS * 32) tmem_read(int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(base + ((it & 3) * 128 + c * 32) % 512, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
      }
    } else if (MODE == 1) {
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(base + ((it & 3) * 128 + c * 32) % 512, r[c]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[c][i]);
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t r[16];
        tmem_ld_x16(base + ((it & 3) * 128 + c * 16) % 512, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += __uint_as_float(r[i]);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  if (warp == 0) { __syncwarp(); tmem_dealloc(slot, 512); }
}

template <int WARPS, int MODE>
void run(const char* name) {
  long long* d; float* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  tmem_read<WARPS, MODE><<<148, WARPS * 32>>>(iters, d, s);
  tmem_read<WARPS, MODE><<<148, WARPS * 32>>>(iters, d, s);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
  const double bytes = double(iters) * WARPS * 32 * 128 * 4;   // per CTA (= per SM)
  printf("%-28s warps=%d  %.0f cycles  %.1f B/clk/SM  (%.1f cycles per 16 KB warp-row)  err=%s\n", name, WARPS, cyc, bytes / cyc,
         cyc / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); cudaFree(s);
}

int main() {
  run<4, 0>("x32+wait each");
  run<8, 0>("x32+wait each");
  run<16, 0>("x32+wait each");
  run<4, 1>("4x x32 then wait");
  run<8, 1>("4x x32 then wait");
  run<16, 1>("4x x32 then wait");
  run<4, 2>("x16+wait each");
  run<8, 2>("x16+wait each");
  return 0;
}
