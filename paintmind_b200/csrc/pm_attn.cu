// pm_attn.cu — flash-style multi-head attention on tcgen05 / TMEM (sm_100a), head_dim = 64.
//
// Replaces CrossAttention.forward / MemoryEfficientCrossAttention.forward between the q/k/v
// projections and to_out (reference modules/attention.py:51-58 and :84-106):
//     out[b, n, h*64:(h+1)*64] = softmax(scale * Q_bh K_bh^T) V_bh        (no mask, no bias)
// Q/K/V are read in place from the token-major projection outputs ([B, N, ld] with the head at
// column h*64) through 3-D TMA tensor maps, so no '(b h) n d' rearrange is ever materialised;
// O is written token-major, ready for the to_out GEMM.
//
// One CTA per (256-query block, head, batch), one CTA per SM.  384 threads (register file rebalanced
// with setmaxnreg: 216 per softmax thread, 64 for the TMA / MMA warpgroup):
//   warps 0-3 / 4-7 : softmax warpgroup 0 / 1 — one thread per query row of Q tile 0 / 1:
//                     fp32 online max / sum with lazy rescaling of the TMEM-resident O accumulator,
//                     exp2 on the pre-scaled scores, P written back over S as packed bf16
//   warp 8 lane 0   : TMA producer — both Q tiles once, K and V tiles of 128 keys in 3-stage rings
//   warp 9 lane 0   : MMA issuer   — S_t = Q_t K^T (SS, 128x128x64) and O_t += P_t V (TS: P from
//                     TMEM, V as MN-major smem operand, 128x64x128), interleaved between the two
//                     Q tiles so the tensor core works on one tile while the other is in softmax
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384); P_t aliases S_t[0,64).
#include "pm_common.cuh"
#include "pm_kernels.h"

namespace pm {

constexpr int AT_BM = 128;      // queries per tile (two tiles per CTA)
constexpr int AT_BN = 128;      // keys per tile
constexpr int AT_D = 64;        // head dim
constexpr int AT_TILE_BYTES = 128 * 64 * 2;   // 16 KB (Q, K and V tiles alike)
constexpr int AT_KV_STAGES = 3;
constexpr int AT_THREADS = 384;            // 3 warpgroups: softmax 0, softmax 1, {TMA, MMA, 2 idle warps}
constexpr int AT_TMEM_COLS = 512;
constexpr int AT_SMEM_BYTES = 1024 + (2 + 2 * AT_KV_STAGES) * AT_TILE_BYTES + 256;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
            const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smQ = smem;                                       // [2][16 KB]
  uint8_t* smK = smem + 2 * AT_TILE_BYTES;                   // [AT_KV_STAGES][16 KB]
  uint8_t* smV = smK + AT_KV_STAGES * AT_TILE_BYTES;         // [AT_KV_STAGES][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smV + AT_KV_STAGES * AT_TILE_BYTES);
  uint64_t* q_full = bars;                                   // [1]
  uint64_t* k_full = bars + 1;                               // [3]
  uint64_t* k_empty = k_full + AT_KV_STAGES;                 // [3]
  uint64_t* v_full = k_empty + AT_KV_STAGES;                 // [3]
  uint64_t* v_empty = v_full + AT_KV_STAGES;                 // [3]
  uint64_t* s_full = v_empty + AT_KV_STAGES;                 // [2]
  uint64_t* p_full = s_full + 2;                             // [2]
  uint64_t* o_full = p_full + 2;                             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Nk + AT_BN - 1) / AT_BN;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);      // one arrival per softmax warp of the group
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 8) {
    // ===================================== TMA producer ======================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * AT_TILE_BYTES);
      tma_load_3d(smQ, &tmQ, q_full, h * AT_D, qb * 2 * AT_BM, b);
      tma_load_3d(smQ + AT_TILE_BYTES, &tmQ, q_full, h * AT_D, qb * 2 * AT_BM + AT_BM, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % AT_KV_STAGES;
        const uint32_t ph = ((j / AT_KV_STAGES) & 1) ^ 1;
        mbar_wait(&k_empty[st], ph);
        mbar_arrive_expect_tx(&k_full[st], AT_TILE_BYTES);
        tma_load_3d(smK + st * AT_TILE_BYTES, &tmK, &k_full[st], h * AT_D, j * AT_BN, b);
        mbar_wait(&v_empty[st], ph);
        mbar_arrive_expect_tx(&v_full[st], AT_TILE_BYTES);
        tma_load_3d(smV + st * AT_TILE_BYTES, &tmV, &v_full[st], h * AT_D, j * AT_BN, b);
      }
    }
  } else if (warp == 9) {
    // ===================================== MMA issuer ========================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);   // Q, K both K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_BM, AT_D, 0, 1);    // P K-major (TMEM), V MN-major
      const uint32_t tS[2] = {tmem_base, tmem_base + 128};
      const uint32_t tO[2] = {tmem_base + 256, tmem_base + 320};
      const uint64_t dq[2] = {umma_desc_sw128(smem_u32(smQ)), umma_desc_sw128(smem_u32(smQ + AT_TILE_BYTES))};

      auto issue_qk = [&](int t, int j) {
        const uint64_t dk = umma_desc_sw128(smem_u32(smK + (j % AT_KV_STAGES) * AT_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_ss(tS[t], dq[t] + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {
        const uint64_t dv = umma_desc_sw128(smem_u32(smV + (j % AT_KV_STAGES) * AT_TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk) {
          // A: 16 keys = 8 TMEM columns of packed bf16;  B: 16 key rows x 128 B = 2048 B
          umma_ts(tO[t], tS[t] + 8 * kk, dv + kk * (2048 >> 4), idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % AT_KV_STAGES;
        const uint32_t ph = (j / AT_KV_STAGES) & 1;
        const bool more = (j + 1 < n_kv);
        const int st1 = (j + 1) % AT_KV_STAGES;
        const uint32_t ph1 = ((j + 1) / AT_KV_STAGES) & 1;
        mbar_wait(&v_full[st], ph);
        // ---- tile 0 ----
        mbar_wait(&p_full[0], j & 1);
        tc_fence_after();
        issue_pv(0, j);
        if (more) {
          mbar_wait(&k_full[st1], ph1);
          tc_fence_after();
          issue_qk(0, j + 1);               // S0 is overwritten only after PV0(j) has consumed P0 (in-order pipe)
        } else {
          umma_commit(&o_full[0]);
        }
        // ---- tile 1 ----
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        issue_pv(1, j);
        umma_commit(&v_empty[st]);
        if (more) {
          issue_qk(1, j + 1);
          umma_commit(&k_empty[st1]);
        } else {
          umma_commit(&o_full[1]);
        }
      }
    }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ===================================== softmax warpgroups ================================
    const int t = warp >> 2;                       // Q tile / warpgroup 0 or 1
    const int q = warp & 3;                        // TMEM lane quarter
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + t * 64 + lane_off;
    const float c = p.scale_log2;                  // softmax scale * log2(e)
    float m_used = -INFINITY;                      // running max (scaled, log2 domain) the accumulators refer to
    float2 l2 = make_float2(0.0f, 0.0f);           // running sum of exp2(s*c - m_used), two partial lanes

    for (int j = 0; j < n_kv; ++j) {
      const int valid = p.Nk - j * AT_BN;          // >= 128 for full tiles
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      // ---- whole score row into registers ----
      uint32_t s[4][32];
      tmem_ld_x32(tS + 0, s[0]);
      tmem_ld_x32(tS + 32, s[1]);
      tmem_ld_x32(tS + 64, s[2]);
      tmem_ld_x32(tS + 96, s[3]);
      tmem_ld_wait();
      if (valid < AT_BN) {                         // ragged last key tile (e.g. 77 text tokens): mask the tail
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ch * 32 + i >= valid) s[ch][i] = 0xff800000u;      // -inf
      }
      // ---- row max (3-input max) ----
      float mt = -INFINITY;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int i = 0; i < 32; i += 2) mt = fmaxf(mt, fmaxf(__uint_as_float(s[ch][i]), __uint_as_float(s[ch][i + 1])));
      const float m_new = fmaxf(m_used, mt * c);
      // lazy rescale: keep the stale max while it is within 2^8 of the true one (exact algebra,
      // bounded magnitude); decided per warp to keep the TMEM traffic warp-uniform
      const bool need = (m_new - m_used) > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2_approx(m_used - m_new) : 1.0f;
        if (need) m_used = m_new;
        l2.x *= alpha;
        l2.y *= alpha;
        if (j > 0) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t r[32];
            tmem_ld_x32(tO + cc * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_x32(tO + cc * 32, r);
          }
        }
      }
      // ---- P = exp2(s*c - m_used) -> packed bf16 over S columns [0, 64) ----
      const float2 cc2 = make_float2(c, c);
      const float2 mm2 = make_float2(-m_used, -m_used);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = make_float2(__uint_as_float(s[ch][i]), __uint_as_float(s[ch][i + 1]));
          const float2 a = __ffma2_rn(x, cc2, mm2);
          const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
          l2 = __fadd2_rn(l2, e);
          pk[i >> 1] = pack_bf16x2(e.x, e.y);
        }
        tmem_st_x16(tS + ch * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }

    // ---- epilogue: O / l -> bf16 -> swizzled smem (this tile's Q buffer) -> TMA store ----
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / (l2.x + l2.y);
    uint8_t* stg_base = smQ + t * AT_TILE_BYTES;
    uint8_t* stg = stg_base + row_in_tile * 128;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t r[32];
      tmem_ld_x32(tO + cc * 32, r);
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
    named_bar_sync(1 + t, 128);
    if (q == 0 && lane == 0) {
      asm volatile(
          "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
          ::"l"(reinterpret_cast<uint64_t>(&tmO)),
          "r"(smem_u32(stg_base)), "r"(h * AT_D), "r"(qb * 2 * AT_BM + t * AT_BM), "r"(b)
          : "memory");
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
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
  dim3 grid((p.Nq + 2 * AT_BM - 1) / (2 * AT_BM), p.H, p.B);
  attn_kernel<<<grid, AT_THREADS, AT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
