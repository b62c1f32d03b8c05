// Micro-benchmark: cycles per tcgen05.mma (kind::f16, cta_group::1, M = 128) for the shapes the attention kernels issue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate.bin scripts/ubench/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../paintmind_b200/csrc/pm_common.cuh"
using namespace pm;

// mode 0: SS, A K-major, B K-major, N;  mode 1: TS (A from TMEM), B MN-major, N;  mode 2: SS both MN-major;
// mode 3: alternate SS (N) and TS (64) like the backward kernel;  same accumulator unless `rot`
__global__ void k(int mode, int N, int iters, int rot, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t tm = slot;
    const uint32_t id_ss = umma_idesc_bf16(128, N, 0, 0), id_ts = umma_idesc_bf16(128, N, 0, 1), id_mn = umma_idesc_bf16(128, N, 1, 1);
    const uint32_t id_ts64 = umma_idesc_bf16(128, 64, 0, 1);
    const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 32768);
    const uint64_t dam = umma_desc_sw128_mn(base, 8192), dbm = umma_desc_sw128_mn(base + 32768, 8192);
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tm + (rot ? (i & 1) * 256 : 0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (mode == 0) umma_ss(d, da + 2 * kk, db + 2 * kk, id_ss, 1);
        else if (mode == 1) umma_ts(d, tm + 448 + 8 * kk, db + kk * 128, id_ts, 1);
        else if (mode == 2) umma_ss(d, dam + kk * 128, dbm + kk * 128, id_mn, 1);
        else { umma_ss(d, da + 2 * kk, db + 2 * kk, id_ss, 1); umma_ts(tm + 384, tm + 448 + 8 * kk, db + kk * 128, id_ts64, 1); }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[1] = static_cast<long long>(g1 - g0);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 20000;
  struct { int mode, N, rot; const char* name; } cases[] = {
      {0, 64, 0, "SS K-major  128x64x16 "}, {0, 128, 0, "SS K-major  128x128x16"}, {0, 256, 0, "SS K-major  128x256x16"},
      {1, 64, 0, "TS B=MN     128x64x16 "}, {1, 128, 0, "TS B=MN     128x128x16"}, {2, 128, 0, "SS MN/MN    128x128x16"},
      {2, 256, 0, "SS MN/MN    128x256x16"}, {3, 64, 0, "SS64+TS64 interleaved "}, {3, 128, 0, "SS128+TS64 interleaved"},
      {0, 64, 1, "SS 128x64x16 2 accums "}, {1, 64, 1, "TS 128x64x16 2 accums "}};
  for (auto& c : cases) {
    k<<<1, 128, 100 * 1024>>>(c.mode, c.N, iters, c.rot, d); cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%s: %.1f cycles = %.1f ns per MMA%s (SM clock %.0f MHz)\n", c.name, double(h[0]) / (iters * 4), double(h[1]) / (iters * 4),
           c.mode == 3 ? " pair" : "", 1e3 * double(h[0]) / double(h[1]));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
