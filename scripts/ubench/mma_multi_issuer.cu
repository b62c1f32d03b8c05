// Micro-benchmark: is the ~73-cycle cost of a tcgen05.mma (M = 128, N <= 128, measured by mma_rate.cu from ONE issuing
// thread) a property of the tensor pipe or of the issuing thread?  1, 2 or 4 threads (lane 0 of warps 0..3) of one CTA issue
// `iters` MMAs each into their own TMEM accumulator (128 columns apart) from their own shared-memory operand tiles; the time
// until every thread's commit has arrived is reported per MMA (all issuers counted).  If the pipe is the limit the
// cycles-per-MMA stay at ~73; if the issuing thread is, they drop with the number of issuers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../paintmind_b200/csrc/pm_common.cuh"
using namespace pm;

__global__ void k(int n_issuers, int N, int ts, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar[4];
  __shared__ long long t_done[4];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const long long t0 = clock64();
  if (warp < n_issuers && lane == 0) {
    const uint32_t tm = slot + (ts ? warp * 64 : warp * 128);       // TS: accumulators of 64 columns, A operands in the upper half
    const uint32_t id = ts ? umma_idesc_bf16(128, N, 0, 1) : umma_idesc_bf16(128, N, 0, 0);
    const uint64_t da = umma_desc_sw128(base + warp * 32768), db = umma_desc_sw128(base + warp * 32768 + 16384);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (!ts) umma_ss(tm, da + 2 * kk, db + 2 * kk, id, 1);
        else umma_ts(tm, slot + 256 + warp * 64 + 8 * kk, db + kk * 128, id, 1);
      }
    }
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    t_done[warp] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = 0;
    for (int i = 0; i < n_issuers; ++i) m = t_done[i] > m ? t_done[i] : m;
    out[0] = m;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int iters = 10000;
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 128})
      for (int ni : {1, 2, 4}) {
        if (ts && N == 128) continue;
        k<<<1, 128, 170 * 1024>>>(ni, N, ts, iters, d); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%s 128x%dx16, %d issuer thread(s): %.1f cycles per MMA overall, %.1f per MMA per issuer  (%s)\n", ts ? "TS" : "SS", N, ni,
               double(h) / (double(iters) * 4 * ni), double(h) / (double(iters) * 4), cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
