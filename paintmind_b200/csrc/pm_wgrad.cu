// pm_wgrad.cu — weight-gradient GEMM on tcgen05 (sm_100a) + column-sum (bias-gradient) kernels.
//
// Backward of every nn.Linear / patch Conv2d of the generator path (the reference gets these
// from autograd: utils/trainer.py:205-225 `accelerator.backward(loss)` through modules/attention.py:34-41,
// modules/mlp.py:28-31, stage1/layers.py:82,129, stage1/vqmodel.py:13-14):
//
//     dW[N, K] = dY[M, N]^T · X[M, K]            (contraction over the M = B * tokens rows)
//     db[N]    = sum_m dY[m, N]
//
// Both operands are read in place, token-major, exactly as the forward / dgrad kernels left them: a 3-D TMA
// tensor map views [M, width] as [width / 64 groups][M tokens][64 columns] and drops a box of G groups x 64 tokens
// x 128 B into shared memory, which IS the canonical 128B-swizzled MN-major UMMA operand layout
// ((8, n), (8, k)) : ((1, LBO), (8, SBO)) with LBO = 8 KB (next 64-feature group), SBO = 1 KB (next 8 tokens).
// No transposed copy of dY or X is ever written.  The token dimension is split over `splits` CTAs per output tile
// (a 512 x 512 weight has only 8 tiles); partial tiles go to an fp32 workspace and are summed in a fixed order
// (deterministic, no atomics).
#include "pm_common.cuh"
#include "pm_kernels.h"

namespace pm {

constexpr int WG_TOK = 64;                 // tokens per pipeline stage
constexpr int WG_STAGES = 4;
constexpr int WG_THREADS = 192;            // TMA warp, MMA warp, 4 epilogue warps
constexpr int WG_GROUP_BYTES = WG_TOK * 128;   // one 64-feature group of a stage: 64 tokens x 128 B = 8 KB

template <int GB>
struct WgradCfg {
  static constexpr int A_BYTES = 2 * WG_GROUP_BYTES;        // 128 dY features
  static constexpr int B_BYTES = GB * WG_GROUP_BYTES;       // 64 * GB X features
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = 1024 + WG_STAGES * STAGE_BYTES + 256;
  static constexpr int TMEM_COLS = 64 * GB;
};

template <int GB>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  using Cfg = WgradCfg<GB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_a = smem_u32(smem_raw);
  const uint32_t base = (raw_a + 1023u) & ~1023u;
  const uint32_t sA = base;                                   // [STAGES][A_BYTES]
  const uint32_t sB = base + WG_STAGES * Cfg::A_BYTES;        // [STAGES][B_BYTES]
  const uint32_t bars = base + WG_STAGES * Cfg::STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * WG_STAGES, tfull_bar = empty_bar + 8 * WG_STAGES;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (tfull_bar + 8 - raw_a));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128, k0 = blockIdx.y * 64 * GB, split = blockIdx.z;
  const int blocks_total = (p.M + WG_TOK - 1) / WG_TOK;
  const int per = (blocks_total + p.splits - 1) / p.splits;
  const int kb0 = split * per;
  const int kb1 = kb0 + per < blocks_total ? kb0 + per : blocks_total;
  const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < WG_STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_bar + 8 * i));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty_bar + 8 * i));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tfull_bar));
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Service warps run CONVERGED and an elected lane issues TMA / MMA / commit (-DPM_WG_LANE0=1: the old divergent lane-0 loops,
  // in which every tcgen05.mma costs ~16 instructions of uniform-register plumbing — see pm_attn4.cu)
#ifdef PM_WG_LANE0
#define WG_SERVICE_LANES (lane == 0)
#define WG_ONE
#else
#define WG_SERVICE_LANES true
#define WG_ONE if (elect_one())
#endif
  if (warp == 0) {
    if (WG_SERVICE_LANES) {
      for (int i = 0; i < nkb; ++i) {
        const int st = i % WG_STAGES;
        mbar_wait_a(empty_bar + 8 * st, ((i / WG_STAGES) & 1) ^ 1);
        WG_ONE {
          mbar_arrive_expect_tx_a(full_bar + 8 * st, Cfg::STAGE_BYTES);
          const int tok = (kb0 + i) * WG_TOK;
          tma_load_3d_a(sA + st * Cfg::A_BYTES, &tmA, full_bar + 8 * st, 0, tok, n0 / 64);
          tma_load_3d_a(sB + st * Cfg::B_BYTES, &tmB, full_bar + 8 * st, 0, tok, k0 / 64);
        }
      }
    }
  } else if (warp == 1) {
    if (WG_SERVICE_LANES && nkb > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64 * GB, 1, 1);       // both operands MN-major
      for (int i = 0; i < nkb; ++i) {
        const int st = i % WG_STAGES;
        mbar_wait_a(full_bar + 8 * st, (i / WG_STAGES) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128_mn(sA + st * Cfg::A_BYTES, WG_GROUP_BYTES);
        const uint64_t db = umma_desc_sw128_mn(sB + st * Cfg::B_BYTES, WG_GROUP_BYTES);
        WG_ONE {
#pragma unroll
          for (int k = 0; k < WG_TOK / 16; ++k)       // 16 tokens = 2 KB further along the contraction dimension
            umma_ss(tmem_base, da + k * (2048 >> 4), db + k * (2048 >> 4), idesc, (i | k) != 0 ? 1u : 0u);
          umma_commit_a(empty_bar + 8 * st);
          if (i == nkb - 1) umma_commit_a(tfull_bar);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    float* dst = p.work + (static_cast<size_t>(split) * p.N + n) * p.K + k0;
    if (nkb > 0) {
      mbar_wait_a(tfull_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < 2 * GB; ++c) {
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (n < p.N) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (k0 + c * 32 + j < p.K)
            *reinterpret_cast<uint4*>(dst + c * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// out[n, k] (+)= scale * sum_s work[s, n, k]   (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ work, int splits, long long total4, int K, float* __restrict__ out,
                                    int64_t ld_out, int accumulate, float scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = reinterpret_cast<const float4*>(work)[static_cast<size_t>(s) * total4 + i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const long long e = i * 4;
  const long long row = e / K, col = e - row * K;
  float4* o = reinterpret_cast<float4*>(out + row * ld_out + col);
  if (accumulate) {
    const float4 old = *o;
    acc.x = old.x + scale * acc.x; acc.y = old.y + scale * acc.y; acc.z = old.z + scale * acc.z; acc.w = old.w + scale * acc.w;
  } else {
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
  }
  *o = acc;
}

template <int GB>
static int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgradCfg<GB>;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = pm_make_tmap_3d_box(&tmA, p.dy, 2, (p.N + 63) / 64, p.M, 64, p.lddy, 64, 2, WG_TOK, 64)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d_box(&tmB, p.x, 2, (p.K + 63) / 64, p.M, 64, p.ldx, 64, GB, WG_TOK, 64)) != PM_OK) return rc;
  auto kern = wgrad_kernel<GB>;
  static bool attr_done[PM_MAX_DEVICES] = {};
  if ((rc = pm_ensure_dyn_smem(kern, Cfg::SMEM_BYTES, attr_done)) != 0) return rc;
  dim3 grid((p.N + 127) / 128, (p.K + 64 * GB - 1) / (64 * GB), p.splits);
  kern<<<grid, WG_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return static_cast<int>(cudaGetLastError());
}

int pm_wgrad_splits(int M, int N, int K) {
  const int gb = K >= 256 ? 4 : (K >= 128 ? 2 : 1);
  const int tiles = ((N + 127) / 128) * ((K + 64 * gb - 1) / (64 * gb));
  const int blocks_total = (M + WG_TOK - 1) / WG_TOK;
  int splits = (2 * pm_num_sms() + tiles - 1) / tiles;        // ~two waves' worth of CTAs: tails even out
  if (splits > 64) splits = 64;
  if (splits > blocks_total) splits = blocks_total;
  if (splits < 1) splits = 1;
  const int per = (blocks_total + splits - 1) / splits;
  return (blocks_total + per - 1) / per;                       // no empty split
}

int pm_wgrad_launch(const WgradParams& p_in, cudaStream_t stream) {
  WgradParams p = p_in;
  if (p.dy == nullptr || p.x == nullptr || p.out == nullptr || p.work == nullptr || p.M <= 0 || p.N <= 0 || p.K <= 0) return PM_ERR_INVALID;
  // whole 64-column groups must exist in memory (TMA reads them even when N or K end inside one)
  if ((p.lddy % 8) != 0 || (p.ldx % 8) != 0 || (p.K % 4) != 0 || (p.ld_out % 4) != 0) return PM_ERR_INVALID;
  if (((p.N + 63) / 64) * 64 > p.lddy || ((p.K + 63) / 64) * 64 > p.ldx) return PM_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p.out) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.work) & 15) != 0) return PM_ERR_INVALID;
  if (p.splits <= 0) p.splits = pm_wgrad_splits(p.M, p.N, p.K);
  int rc;
  if (p.K >= 256) rc = launch_wgrad<4>(p, stream);
  else if (p.K >= 128) rc = launch_wgrad<2>(p, stream);
  else rc = launch_wgrad<1>(p, stream);
  if (rc != 0) return rc;
  const long long total4 = static_cast<long long>(p.N) * p.K / 4;
  wgrad_reduce_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(p.work, p.splits, total4, p.K, p.out, p.ld_out,
                                                                                      p.accumulate, p.scale);
  return static_cast<int>(cudaGetLastError());
}

// ----------------------------------------------------------------------------------------------
// Column sums of a bf16 [M, N] matrix (bias gradients; position-embedding gradients when viewed as [B, tokens * D]).
// Two stages, fixed summation order: partial[R, N] then out[N].
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int M, int N, int rows_per, float* __restrict__ partial) {
  __shared__ float red[8][32][8];
  const int cv = blockIdx.x * 32 + threadIdx.x;             // 8-column vector index
  const int r0 = blockIdx.y * rows_per;
  const int r1 = r0 + rows_per < M ? r0 + rows_per : M;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (cv * 8 < N) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * ld + cv * 8);
      acc[0] += bf16lo_to_f32(u.x); acc[1] += bf16hi_to_f32(u.x);
      acc[2] += bf16lo_to_f32(u.y); acc[3] += bf16hi_to_f32(u.y);
      acc[4] += bf16lo_to_f32(u.z); acc[5] += bf16hi_to_f32(u.z);
      acc[6] += bf16lo_to_f32(u.w); acc[7] += bf16hi_to_f32(u.w);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.y][threadIdx.x][k] = acc[k];
  __syncthreads();
  if (threadIdx.y == 0 && cv * 8 < N) {
    float* dst = partial + static_cast<size_t>(blockIdx.y) * N + cv * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x][k];
      dst[k] = s;
    }
  }
}

// out[c] (+)= sum_r partial[r, c], fixed order: block = 32 columns x 32 row lanes (row lane y sums rows y, y + 32, ... and
// the 32 lane sums are added in order) — one thread per column walking hundreds of partial rows was latency-bound
// (~0.1 ms for a [592, 1024] partial, more than the pass that produced it).
__global__ void __launch_bounds__(1024)
colreduce_kernel(const float* __restrict__ partial, int R, int N, float* __restrict__ out, int accumulate) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < N)
    for (int r = threadIdx.y; r < R; r += 32) s += partial[static_cast<size_t>(r) * N + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += red[y][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}

int pm_colreduce_launch(const float* partial, int R, int N, float* out, int accumulate, cudaStream_t stream) {
  colreduce_kernel<<<(N + 31) / 32, dim3(32, 32), 0, stream>>>(partial, R, N, out, accumulate);
  return static_cast<int>(cudaGetLastError());
}

int pm_colsum_rows(int M, int N) {
  const int col_blocks = (N + 255) / 256;
  int R = (4 * pm_num_sms() + col_blocks - 1) / col_blocks;
  const int max_r = (M + 63) / 64;
  if (R > max_r) R = max_r;
  if (R < 1) R = 1;
  const int rows_per = (M + R - 1) / R;
  return (M + rows_per - 1) / rows_per;
}

int pm_colsum_launch(const void* x, int64_t ld, int M, int N, float* partial, float* out, int accumulate, cudaStream_t stream) {
  if (x == nullptr || partial == nullptr || out == nullptr || M <= 0 || N <= 0 || (N % 8) != 0 || (ld % 8) != 0) return PM_ERR_INVALID;
  const int R = pm_colsum_rows(M, N);
  const int rows_per = (M + R - 1) / R;
  dim3 grid((N + 255) / 256, R), block(32, 8);
  colsum_bf16_kernel<<<grid, block, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, M, N, rows_per, partial);
  return pm_colreduce_launch(partial, R, N, out, accumulate, stream);
}

}  // namespace pm
