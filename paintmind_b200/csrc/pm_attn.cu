// pm_attn.cu — flash-style multi-head attention on tcgen05 / TMEM (sm_100a), head_dim = 64.
//
// Replaces CrossAttention.forward / MemoryEfficientCrossAttention.forward between the q/k/v
// projections and to_out (reference modules/attention.py:51-58 and :84-106):
//     out[b, n, h*64:(h+1)*64] = softmax(scale * Q_bh K_bh^T) V_bh        (no mask, no bias)
// Q/K/V are read in place from the token-major projection outputs ([B, N, ld] with the head at
// column h*64) through 3-D TMA tensor maps, so no '(b h) n d' rearrange is ever materialised;
// O is written token-major, ready for the to_out GEMM.
//
// One CTA per (128-query tile, head, batch); two CTAs co-reside per SM so that one CTA's softmax
// overlaps the other's MMAs.  Per CTA (192 threads):
//   warp 0 lane 0 : TMA producer (Q once; K/V tiles of 128 keys in a 2-stage ring)
//   warp 1 lane 0 : MMA issuer   S = Q K^T (SS, 128x128x64) ;  O += P V (TS: P from TMEM,
//                                V as MN-major smem operand, 128x64x128)
//   warps 2..5    : softmax, one thread per query row: online max / sum in fp32 with lazy
//                   rescaling of the TMEM-resident O accumulator, P written back over S as bf16.
#include "pm_common.cuh"
#include "pm_kernels.h"

namespace pm {

constexpr int AT_BM = 128;      // queries per CTA
constexpr int AT_BN = 128;      // keys per tile
constexpr int AT_D = 64;        // head dim
constexpr int AT_TILE_BYTES = 128 * 64 * 2;   // 16 KB (Q, K and V tiles alike)
constexpr int AT_KV_STAGES = 2;
constexpr int AT_THREADS = 192;
constexpr int AT_TMEM_COLS = 256;             // S: [0,128)  O: [128,192)
constexpr int AT_SMEM_BYTES = 1024 + (1 + 2 * AT_KV_STAGES) * AT_TILE_BYTES + 256;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
            const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem;
  uint8_t* smK = smem + AT_TILE_BYTES;
  uint8_t* smV = smK + AT_KV_STAGES * AT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smV + AT_KV_STAGES * AT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                 // [2]
  uint64_t* kv_empty = bars + 3;                // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_kv_tiles = (p.Nk + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);      // one arrival per softmax warp
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, AT_TILE_BYTES);
      tma_load_3d(smQ, &tmQ, q_full, h * AT_D, qt * AT_BM, b);
      for (int j = 0; j < n_kv_tiles; ++j) {
        const int st = j % AT_KV_STAGES;
        mbar_wait(&kv_empty[st], ((j / AT_KV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * AT_TILE_BYTES);
        tma_load_3d(smK + st * AT_TILE_BYTES, &tmK, &kv_full[st], h * AT_D, j * AT_BN, b);
        tma_load_3d(smV + st * AT_TILE_BYTES, &tmV, &kv_full[st], h * AT_D, j * AT_BN, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);   // Q, K both K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_BM, AT_D, 0, 1);    // P K-major (TMEM), V MN-major
      mbar_wait(q_full, 0);
      const uint64_t dq = umma_desc_sw128(smem_u32(smQ));
      for (int j = 0; j < n_kv_tiles; ++j) {
        const int st = j % AT_KV_STAGES;
        mbar_wait(&kv_full[st], (j / AT_KV_STAGES) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(smK + st * AT_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(s_full);
        // softmax turns S into P (bf16, in place) and rescales O when the running max moved
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t dv = umma_desc_sw128(smem_u32(smV + st * AT_TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk) {
          // A: 16 keys = 8 TMEM columns of packed bf16;  B: 16 key rows x 128 B = 2048 B
          umma_ts(tmem_O, tmem_S + 8 * kk, dv + kk * (2048 >> 4), idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        if (j == n_kv_tiles - 1) umma_commit(o_full);
      }
    }
  } else {
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c = p.scale_log2;        // softmax scale * log2(e)
    float m_used = -INFINITY;            // running max (scaled, log2 domain) the accumulators refer to
    float l = 0.0f;                      // running sum of exp2(t - m_used)

    for (int j = 0; j < n_kv_tiles; ++j) {
      const int valid = min(AT_BN, p.Nk - j * AT_BN);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // ---- pass 1: row max of this tile ----
      float mt = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld_x32(tmem_S + lane_off + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (cc * 32 + i < valid) ? __uint_as_float(r[i]) : -INFINITY;
          mt = fmaxf(mt, s);
        }
      }
      const float m_new = fmaxf(m_used, mt * c);
      // lazy rescale: keep the stale max while it is within 2^8 of the true one (exact algebra,
      // bounded magnitude); the decision is made per warp to keep TMEM traffic warp-uniform
      const bool need = (m_new - m_used) > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2_approx(m_used - m_new) : 1.0f;
        if (need) m_used = m_new;
        l *= alpha;
        if (j > 0) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t r[32];
            tmem_ld_x32(tmem_O + lane_off + cc * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_x32(tmem_O + lane_off + cc * 32, r);
          }
        }
      }
      // ---- pass 2: P = exp2(s*c - m_used) -> bf16, written over S (cols 16*cc .. 16*cc+15) ----
      const float moff = m_used;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld_x32(tmem_S + lane_off + cc * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float s0 = (cc * 32 + i < valid) ? __uint_as_float(r[i]) : -INFINITY;
          const float s1 = (cc * 32 + i + 1 < valid) ? __uint_as_float(r[i + 1]) : -INFINITY;
          const float p0 = ex2_approx(fmaf(s0, c, -moff));
          const float p1 = ex2_approx(fmaf(s1, c, -moff));
          l += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st_x16(tmem_S + lane_off + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> swizzled smem (Q buffer) -> TMA store ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    uint8_t* stg = smQ + row_in_tile * 128;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t r[32];
      tmem_ld_x32(tmem_O + lane_off + cc * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int jv = 0; jv < 4; ++jv) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r[jv * 8 + 0]) * inv_l, __uint_as_float(r[jv * 8 + 1]) * inv_l);
        o.y = pack_bf16x2(__uint_as_float(r[jv * 8 + 2]) * inv_l, __uint_as_float(r[jv * 8 + 3]) * inv_l);
        o.z = pack_bf16x2(__uint_as_float(r[jv * 8 + 4]) * inv_l, __uint_as_float(r[jv * 8 + 5]) * inv_l);
        o.w = pack_bf16x2(__uint_as_float(r[jv * 8 + 6]) * inv_l, __uint_as_float(r[jv * 8 + 7]) * inv_l);
        const int chunk = cc * 4 + jv;
        *reinterpret_cast<uint4*>(stg + ((chunk ^ (row_in_tile & 7)) << 4)) = o;
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (warp == 2 && lane == 0) {
      asm volatile(
          "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
          ::"l"(reinterpret_cast<uint64_t>(&tmO)),
          "r"(smem_u32(smQ)), "r"(h * AT_D), "r"(qt * AT_BM), "r"(b)
          : "memory");
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

int pm_attn_launch(const AttnParams& p, cudaStream_t stream) {
  if (p.q == nullptr || p.k == nullptr || p.v == nullptr || p.o == nullptr) return PM_ERR_INVALID;
  if (p.B <= 0 || p.H <= 0 || p.Nq <= 0 || p.Nk <= 0 || p.head_dim != AT_D) return PM_ERR_INVALID;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  const uint64_t inner = static_cast<uint64_t>(p.H) * AT_D;
  if ((rc = pm_make_tmap_3d(&tmQ, p.q, 2, p.B, p.Nq, inner, p.ldq, p.bsq, AT_BM, AT_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmK, p.k, 2, p.B, p.Nk, inner, p.ldk, p.bsk, AT_BN, AT_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmV, p.v, 2, p.B, p.Nk, inner, p.ldv, p.bsv, AT_BN, AT_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmO, p.o, 2, p.B, p.Nq, inner, p.ldo, p.bso, AT_BM, AT_D)) != PM_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  dim3 grid((p.Nq + AT_BM - 1) / AT_BM, p.H, p.B);
  attn_kernel<<<grid, AT_THREADS, AT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
