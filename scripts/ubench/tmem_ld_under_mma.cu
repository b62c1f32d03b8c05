// Micro-benchmark: latency of tcgen05.ld (x32, load + wait) while ANOTHER thread keeps `depth` tcgen05.mma instructions queued
// (depth MMAs, commit, wait, repeat).  The loads read TMEM columns the MMAs never touch.  Question: do loads queue behind the
// MMAs that were issued before them (one in-order TMEM pipeline), i.e. does their latency grow with the issuer's run-ahead?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_under_mma scripts/ubench/tmem_ld_under_mma.cu && /tmp/tmem_ld_under_mma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../paintmind_b200/csrc/pm_common.cuh"
using namespace pm;

// kind 0: SS 128x128x16 into columns [0,128); kind 1: TS 128x64x16 (A from TMEM columns [448,..), D = [384,448))
__global__ void k(int kind, int depth, int n_ld, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  __shared__ long long lat[4];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 8) {
    if (lane == 0 && depth > 0) {
      const uint32_t id_ss = umma_idesc_bf16(128, 128, 0, 0), id_ts = umma_idesc_bf16(128, 64, 0, 1);
      const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 32768);
      uint32_t ph = 0;
      while (!stop) {
        for (int i = 0; i < depth; ++i) {
          if (kind == 0) umma_ss(tm, da + 2 * (i & 3), db + 2 * (i & 3), id_ss, 1);
          else umma_ts(tm + 384, tm + 448 + 8 * (i & 3), db + (i & 3) * 128, id_ts, 1);
        }
        umma_commit(&bar);
        mbar_wait(&bar, ph);
        ph ^= 1;
      }
    }
  } else if (warp < 4) {
    const uint32_t b = tm + (static_cast<uint32_t>(warp * 32) << 16) + 256;
    uint32_t acc = 0;
    long long total = 0;
    for (int i = 0; i < n_ld; ++i) {
      uint32_t r[32];
      const long long t0 = clock64();
      tmem_ld_x32(b, r);
      tmem_ld_wait();
      total += clock64() - t0;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
      // a short pause so that the load arrives at a random point of the issuer's burst
      for (int s = 0; s < (i * 7 + warp * 13) % 50; ++s) asm volatile("nanosleep.u32 0;");
    }
    if (lane == 0) lat[warp] = total;
    if (acc == 0x12345678u) sink[0] = acc;
  }
  if (warp < 4) asm volatile("bar.sync 1, 128;" ::: "memory");
  if (threadIdx.x == 0) { stop = 1; out[0] = (lat[0] + lat[1] + lat[2] + lat[3]) / 4; }
  tc_fence_before(); __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n_ld = 4000;
  const char* kn[] = {"SS 128x128x16 (64 cycles each)", "TS 128x64x16 (32 cycles each)"};
  for (int kind = 0; kind < 2; ++kind)
    for (int depth : {0, 1, 2, 4, 8, 16, 32}) {
      k<<<1, 288, 100 * 1024>>>(kind, depth, n_ld, d, s); cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%s, %2d MMAs queued per burst: tcgen05.ld x32 + wait = %.0f cycles (%s)\n", kn[kind], depth, double(h) / n_ld, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
