// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (M = 256 across a CTA pair) for SS and TS shapes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../paintmind_b200/csrc/pm_common.cuh"
using namespace pm;

__device__ __forceinline__ void umma_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// mode 0: SS K-major/K-major N;  mode 1: TS, B MN-major (SW128 atoms; N/2 per CTA -> only meaningful as a timing)
__global__ void __cluster_dims__(2, 1, 1) k(int mode, int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0) { tmem_alloc_2sm(&slot, 512); tmem_relinquish_2sm(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t tm = slot;
    const uint32_t id_ss = umma_idesc_bf16(256, N, 0, 0), id_ts = umma_idesc_bf16(256, N, 0, 1);
    const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 32768);
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (mode == 0) umma_ss_2sm(tm, da + 2 * kk, db + 2 * kk, id_ss, 1);
        else umma_ts_2sm(tm, tm + 448 + 8 * kk, db + kk * 128, id_ts, 1);
      }
    }
    umma_commit_2sm(&bar, 1);
    mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[1] = static_cast<long long>(g1 - g0);
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_2sm(slot, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 20000;
  struct { int mode, N; const char* name; } cases[] = {{0, 64, "SS 2sm 256x64x16 "}, {0, 128, "SS 2sm 256x128x16"}, {0, 256, "SS 2sm 256x256x16"},
                                                      {1, 64, "TS 2sm 256x64x16 "}, {1, 128, "TS 2sm 256x128x16"}};
  for (auto& c : cases) {
    k<<<2, 128, 100 * 1024>>>(c.mode, c.N, iters, d); cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%s: %.1f cycles = %.1f ns per MMA (SM clock %.0f MHz)  (%s)\n", c.name, double(h[0]) / (iters * 4), double(h[1]) / (iters * 4),
           1e3 * double(h[0]) / double(h[1]), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
