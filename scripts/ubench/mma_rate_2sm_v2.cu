// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 256x256x16 (SS, K-major) as the shipped CTA-pair GEMM issues them:
// 4 MMAs per 64-wide k-block out of a ring of operand stages, a commit per k-block, 32 MMAs per accumulator, two
// accumulator stages.  mma_rate_2sm.cu (one operand tile, one accumulator, one commit at the very end) reported 170 cycles;
// the shipped kernel demonstrably sustains ~135 (profiles/r01_gemm_stalls.txt: 4,958 cycles per 256x256x512 tile with 13 %
// operand wait).  Variants: v0 = the old loop; v1 = ring of 4 operand stages; v2 = v1 + commit per k-block (to a barrier
// nobody waits on); v3 = v2 + alternating accumulators every 32 MMAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../paintmind_b200/csrc/pm_common.cuh"
using namespace pm;

__global__ void __cluster_dims__(2, 1, 1) k(int variant, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar, scratch_bar;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0) { tmem_alloc_2sm(&slot, 512); tmem_relinquish_2sm(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&scratch_bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t tm = slot;
    const uint32_t id = umma_idesc_bf16(256, 256, 0, 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {                      // one k-block (4 MMAs) per iteration
      const int st = variant >= 1 ? (i & 3) : 0;
      const uint64_t da = umma_desc_sw128(base + st * 49152), db = umma_desc_sw128(base + st * 49152 + 16384);
      const uint32_t acc = (variant >= 3 && ((i >> 3) & 1)) ? tm + 256 : tm;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_ss_2sm(acc, da + 2 * kk, db + 2 * kk, id, (variant >= 3 && (i & 7) == 0 && kk == 0) ? 0u : 1u);
      if (variant >= 2) umma_commit_2sm(&scratch_bar, 1);
    }
    umma_commit_2sm(&bar, 1);
    mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_2sm(slot, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 20000;
  for (int v = 0; v < 4; ++v) {
    k<<<2, 128, 200 * 1024>>>(v, iters, d); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("2sm 256x256x16 variant %d: %.1f cycles per MMA  (%s)\n", v, double(h) / (iters * 4.0), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
