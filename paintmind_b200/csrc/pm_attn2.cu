// pm_attn2.cu — round-2 forward attention kernel (head_dim 64): the softmax of pm_attn.cu spread over FOUR warps per SM
// sub-partition instead of two.
//
// Same contract as attn_kernel (pm_attn.cu; reference modules/attention.py:51-58 / :84-106):
//     out[b, n, h*64:(h+1)*64] = softmax(scale * Q_bh K_bh^T) V_bh
// and the same pipeline (persistent CTAs over (256-query block, head, batch) items, TMA rings for K / V, S_t = Q_t K^T as
// SS MMAs, O_t += P_t V as TS MMAs with P in its own TMEM columns, lazy rescaling, 1/4 of the exponentials on the FMA pipe).
//
// What changed and why (profiles/r01_ncu_attn.txt): with one thread per query ROW the kernel had two softmax warps per
// scheduler; each warp's step is a serial chain (TMEM load -> max -> 128 exponentials -> TMEM store) of ~650 instructions
// whose static schedule is ~1230 cycles plus ~630 cycles of exposed TMEM / barrier latency, and two such warps overlapped to
// 2870 cycles per step against 1896 cycles of tensor-pipe work: no pipe was more than 60 % busy.  Here every query row is
// shared by TWO threads (same TMEM lane, warps w and w+4), each owning 64 of the tile's 128 keys:
//   * 16 softmax warps, four per scheduler: twice the latency hiding for the same pipes;
//   * both threads need the row maximum: each reads ALL 128 scores (the partner's half only passes through 32 registers at
//     a time) and reduces them itself — bitwise the same value in both threads, so the lazy-rescale decisions agree without
//     any exchange or barrier; only the 64 own scores are exponentiated;
//   * P_t V is issued in two halves (keys 0-63 as soon as the kh = 0 warps have stored their P columns);
//   * the row sums meet once per item (shared memory) in the epilogue, where each thread also scales and stores its 32 of
//     the 64 output columns.
// 640 threads: warps 0-15 softmax, warp 16 TMA producer, warp 17 / 18 MMA issuers.  Registers: 96 per thread at launch
// (640 x 96 = 61,440 is the CTA's pool); setmaxnreg moves them to 104 per softmax thread and 64 per producer / issuer thread
// (512 x 104 + 128 x 64 = 61,440 exactly — an increase beyond the pool blocks forever).
// TMEM (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <stdio.h>
#include <stdlib.h>

namespace pm {

constexpr int A2_BM = 128;      // queries per tile (two tiles per work item)
constexpr int A2_BN = 128;      // keys per tile
constexpr int A2_D = 64;        // head dim
constexpr int A2_TILE_BYTES = 128 * 64 * 2;   // 16 KB (Q, K, V and O tiles alike)
constexpr int A2_KV_STAGES = 3;
constexpr int A2_Q_STAGES = 2;
constexpr int A2_THREADS = 640;
constexpr int A2_TMEM_COLS = 512;
// smem: Q [2 stages][2 tiles] | K [3] | V [3] | O staging [2 tiles] | row-sum exchange [2 tiles][2 halves][128] | barriers
constexpr int A2_SMEM_BYTES = 1024 + (2 * A2_Q_STAGES + 2 * A2_KV_STAGES + 2) * A2_TILE_BYTES + 2 * 2 * 128 * 4 + 512;

__device__ __forceinline__ float a2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct A2Item {
  int qb, h, b;
};
__device__ __forceinline__ A2Item a2_item(int w, int n_qb, int H) {
  A2Item it;
  it.qb = w % n_qb;
  const int r = w / n_qb;
  it.h = r % H;
  it.b = r / H;
  return it;
}

// EMU   : of every 4 pairs of scores, this many take the FMA-pipe exp2 (exp2_poly2)
// CHAIN : > 0: groups of CHAIN score pairs are chained by a value-neutral dependency so that ptxas keeps the MUFU / FMA
//         mix uniform along the row instead of hoisting every polynomial to the front
// TRAIN : also emit the row log-sum-exp and an fp32 copy of the output (training forward)
template <int EMU, int CHAIN, bool TRAIN>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_a = smem_u32(smem_raw);
  const uint32_t sQ = (raw_a + 1023u) & ~1023u;                          // [A2_Q_STAGES][2][16 KB]
  const uint32_t sK = sQ + 2 * A2_Q_STAGES * A2_TILE_BYTES;              // [A2_KV_STAGES][16 KB]
  const uint32_t sV = sK + A2_KV_STAGES * A2_TILE_BYTES;                 // [A2_KV_STAGES][16 KB]
  const uint32_t sO = sV + A2_KV_STAGES * A2_TILE_BYTES;                 // [2][16 KB] output staging
  const uint32_t sL = sO + 2 * A2_TILE_BYTES;                            // [2 tiles][2 halves][128] fp32 partial row sums
  const uint32_t bars = sL + 2 * 2 * 128 * 4;
  const uint32_t q_full = bars;                                   // [2]
  const uint32_t q_empty = q_full + 8 * A2_Q_STAGES;              // [2]
  const uint32_t k_full = q_empty + 8 * A2_Q_STAGES;              // [3]
  const uint32_t k_empty = k_full + 8 * A2_KV_STAGES;             // [3]
  const uint32_t v_full = k_empty + 8 * A2_KV_STAGES;             // [3]
  const uint32_t v_empty = v_full + 8 * A2_KV_STAGES;             // [3]
  const uint32_t s_full = v_empty + 8 * A2_KV_STAGES;             // [2]     MMA -> softmax: S_t complete
  const uint32_t s_free = s_full + 16;                            // [2]     softmax -> MMA: S_t has been read (8 warps)
  const uint32_t p_full = s_free + 16;                            // [2][2]  softmax -> MMA: P_t columns of key half kh written (4 warps)
  const uint32_t pv_done = p_full + 32;                           // [2]     MMA -> softmax: O_t += P_t V complete
  const uint32_t tmem_slot_a = pv_done + 16;
  uint8_t* const smO = smem_raw + (sO - raw_a);
  float* const smL = reinterpret_cast<float*>(smem_raw + (sL - raw_a));
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot_a - raw_a));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = (p.Nk + A2_BN - 1) / A2_BN;
  const int n_qb = (p.Nq + 2 * A2_BM - 1) / (2 * A2_BM);
  const int total_items = n_qb * p.H * p.B;
  const int my_items = static_cast<int>(blockIdx.x) < total_items
                           ? (total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                           : 0;
  const int total_steps = my_items * n_kv;       // flattened stream of (item, key tile) steps

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    for (int i = 0; i < A2_Q_STAGES; ++i) {
      init(q_full + 8 * i, 1);
      init(q_empty + 8 * i, 1);
    }
    for (int i = 0; i < A2_KV_STAGES; ++i) {
      init(k_full + 8 * i, 1);
      init(k_empty + 8 * i, 1);
      init(v_full + 8 * i, 1);
      init(v_empty + 8 * i, 1);
    }
    for (int t = 0; t < 2; ++t) {
      init(s_full + 8 * t, 1);
      init(s_free + 8 * t, 8);               // one arrival per softmax warp of the tile
      init(p_full + 16 * t, 4);              // key half 0: one arrival per warp of that half
      init(p_full + 16 * t + 8, 4);          // key half 1
      init(pv_done + 8 * t, 1);
    }
    fence_mbar_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_slot, A2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 16) {
      // ===================================== TMA producer ======================================
      if (lane == 0) {
        int g = 0;                                   // running key-tile counter (K/V ring position)
        for (int i = 0; i < my_items; ++i) {
          const A2Item it = a2_item(blockIdx.x + i * gridDim.x, n_qb, p.H);
          const int qs = i % A2_Q_STAGES;
          mbar_wait_a(q_empty + 8 * qs, ((i / A2_Q_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx_a(q_full + 8 * qs, 2 * A2_TILE_BYTES);
          tma_load_3d_a(sQ + (2 * qs) * A2_TILE_BYTES, &tmQ, q_full + 8 * qs, it.h * A2_D, it.qb * 2 * A2_BM, it.b);
          tma_load_3d_a(sQ + (2 * qs + 1) * A2_TILE_BYTES, &tmQ, q_full + 8 * qs, it.h * A2_D, it.qb * 2 * A2_BM + A2_BM, it.b);
          for (int j = 0; j < n_kv; ++j, ++g) {
            const int st = g % A2_KV_STAGES;
            const uint32_t ph = ((g / A2_KV_STAGES) & 1) ^ 1;
            mbar_wait_a(k_empty + 8 * st, ph);
            mbar_arrive_expect_tx_a(k_full + 8 * st, A2_TILE_BYTES);
            tma_load_3d_a(sK + st * A2_TILE_BYTES, &tmK, k_full + 8 * st, it.h * A2_D, j * A2_BN, it.b);
            mbar_wait_a(v_empty + 8 * st, ph);
            mbar_arrive_expect_tx_a(v_full + 8 * st, A2_TILE_BYTES);
            tma_load_3d_a(sV + st * A2_TILE_BYTES, &tmV, v_full + 8 * st, it.h * A2_D, j * A2_BN, it.b);
          }
        }
      }
    } else if (warp == 17) {
      // ===================================== MMA issuer 1: S_t = Q_t K^T ========================
      if (lane == 0 && total_steps > 0) {
        constexpr uint32_t idesc_qk = umma_idesc_bf16(A2_BM, A2_BN, 0, 0);   // Q, K both K-major
        const uint32_t tS[2] = {tmem_base, tmem_base + 128};
        for (int g = 0; g < total_steps; ++g) {
          const int i = g / n_kv, j = g - i * n_kv;
          const int qs = i % A2_Q_STAGES, ks = g % A2_KV_STAGES;
          if (j == 0) mbar_wait_a(q_full + 8 * qs, (i / A2_Q_STAGES) & 1);
          mbar_wait_a(k_full + 8 * ks, (g / A2_KV_STAGES) & 1);
          const uint64_t dk = umma_desc_sw128(sK + ks * A2_TILE_BYTES);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (g > 0) mbar_wait_a(s_free + 8 * t, (g - 1) & 1);      // S_t of the previous step has been read by all 8 warps
            tc_fence_after();
            const uint64_t dq = umma_desc_sw128(sQ + (2 * qs + t) * A2_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < A2_D / 16; ++k) umma_ss(tS[t], dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
            umma_commit_a(s_full + 8 * t);
          }
          umma_commit_a(k_empty + 8 * ks);
          if (j == n_kv - 1) umma_commit_a(q_empty + 8 * qs);
        }
      }
    } else if (warp == 18) {
      // ===================================== MMA issuer 2: O_t (+)= P_t V ========================
      if (lane == 0 && total_steps > 0) {
        constexpr uint32_t idesc_pv = umma_idesc_bf16(A2_BM, A2_D, 0, 1);    // P K-major (TMEM), V MN-major
        const uint32_t tP[2] = {tmem_base + 256, tmem_base + 320};
        const uint32_t tO[2] = {tmem_base + 384, tmem_base + 448};
        for (int g = 0; g < total_steps; ++g) {
          const int j = g % n_kv;
          const int vs = g % A2_KV_STAGES;
          mbar_wait_a(v_full + 8 * vs, (g / A2_KV_STAGES) & 1);
          const uint64_t dv = umma_desc_sw128(sV + vs * A2_TILE_BYTES);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            // keys 0-63: the kh = 0 warps (which also did any rescaling of O_t) are done.  On the first step of an item the
            // first MMA OVERWRITES O_t, whose previous contents the kh = 1 warps may still be reading out: wait for both.
            mbar_wait_a(p_full + 16 * t, g & 1);
            if (j == 0) mbar_wait_a(p_full + 16 * t + 8, g & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ts(tO[t], tP[t] + 8 * kk, dv + kk * (2048 >> 4), idesc_pv, (j | kk) != 0 ? 1u : 0u);
            if (j != 0) {
              mbar_wait_a(p_full + 16 * t + 8, g & 1);
              tc_fence_after();
            }
#pragma unroll
            for (int kk = 4; kk < 8; ++kk) umma_ts(tO[t], tP[t] + 8 * kk, dv + kk * (2048 >> 4), idesc_pv, 1u);
            umma_commit_a(pv_done + 8 * t);
          }
          umma_commit_a(v_empty + 8 * vs);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================================== softmax warps =====================================
    const int t = warp >> 3;                       // Q tile 0 / 1
    const int kh = (warp >> 2) & 1;                // key half of every 128-key tile owned by this thread
    const int q = warp & 3;                        // TMEM lane quarter
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS_own = tmem_base + t * 128 + kh * 64 + lane_off;
    const uint32_t tS_oth = tmem_base + t * 128 + (kh ^ 1) * 64 + lane_off;
    const uint32_t tP = tmem_base + 256 + t * 64 + kh * 32 + lane_off;
    const uint32_t tO = tmem_base + 384 + t * 64 + lane_off;
    const uint32_t b_s_full = s_full + 8 * t, b_s_free = s_free + 8 * t;
    const uint32_t b_p_full = p_full + 16 * t + 8 * kh, b_pv_done = pv_done + 8 * t;
    const float c = p.scale_log2;                  // softmax scale * log2(e)
    uint8_t* const stg = smO + t * A2_TILE_BYTES + row_in_tile * 128;
    float* const l_mine = smL + (t * 2 + kh) * 128 + row_in_tile;
    float* const l_other = smL + (t * 2 + (kh ^ 1)) * 128 + row_in_tile;
    const bool tile_leader = (kh == 0 && q == 0 && lane == 0);
    int g = 0;                                     // flattened step counter (barrier phases)

    for (int i = 0; i < my_items; ++i) {
      const A2Item it = a2_item(blockIdx.x + i * gridDim.x, n_qb, p.H);
      float m_used = -INFINITY;                    // running max (scaled, log2 domain) the accumulators refer to
      float2 la = make_float2(0.0f, 0.0f);         // partial sum of exp2(s*c - m_used) over this thread's keys
      float2 lb = make_float2(0.0f, 0.0f);

      for (int j = 0; j < n_kv; ++j, ++g) {
        const int valid = p.Nk - j * A2_BN;        // >= 128 for full tiles
        mbar_wait_a(b_s_full, g & 1);
        tc_fence_after();
        // ---- the partner's 64 scores: only their maximum is needed ----
        float mo;
        {
          uint32_t o0[32], o1[32];
          tmem_ld_x32(tS_oth, o0);
          tmem_ld_x32(tS_oth + 32, o1);
          tmem_ld_wait();
          if (valid < A2_BN) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              if ((kh ^ 1) * 64 + e >= valid) o0[e] = 0xff800000u;
              if ((kh ^ 1) * 64 + 32 + e >= valid) o1[e] = 0xff800000u;
            }
          }
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(o0[e]), __uint_as_float(o0[e + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(o1[e]), __uint_as_float(o1[e + 1])));
          }
          mo = fmaxf(m0, m1);
        }
        // ---- own 64 scores into registers ----
        // The load address carries a data dependency on `mo` (0 unless the partner's scores are all NaN): the 64 registers of
        // the partner's scores must be dead before the 64 own scores arrive — ptxas otherwise hoists these loads above the max
        // tree, needs 128 + 40 registers and spills the score rows to local memory.
        const uint32_t dep = (mo != mo) ? 1u : 0u;
        uint32_t s[2][32];
        tmem_ld_x32(tS_own + dep, s[0]);
        tmem_ld_x32(tS_own + 32 + dep, s[1]);
        tmem_ld_wait();
        // every score this warp needs has been read: S_t may be overwritten by the next Q K^T (8 arrivals)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(b_s_free);
        if (valid < A2_BN) {                       // ragged last key tile (e.g. 77 text tokens): mask the tail
#pragma unroll
          for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (kh * 64 + ch * 32 + e >= valid) s[ch][e] = 0xff800000u;      // -inf
        }
        float mp0 = -INFINITY, mp1 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mp0 = fmaxf(mp0, fmaxf(__uint_as_float(s[0][e]), __uint_as_float(s[0][e + 1])));
          mp1 = fmaxf(mp1, fmaxf(__uint_as_float(s[1][e]), __uint_as_float(s[1][e + 1])));
        }
        // the row maximum: max is exact, so both threads of the row hold bitwise the same value whatever the order
        const float mt = fmaxf(mo, fmaxf(mp0, mp1));
        const float m_new = fmaxf(m_used, mt * c);
        // lazy rescale: keep the stale max while it is within 2^8 of the true one; decided per warp — the partner warp
        // (same rows, same values) takes the same decision
        const bool need = (m_new - m_used) > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? a2_ex2(m_used - m_new) : 1.0f;
          if (need) m_used = m_new;
          la.x *= alpha;
          la.y *= alpha;
          lb.x *= alpha;
          lb.y *= alpha;
          if (j > 0 && kh == 0) {
            // the kh = 0 warp rescales all 64 columns of O_t: the P_t V issuer only waits for kh = 0 before its first half
            mbar_wait_a(b_pv_done, (g - 1) & 1);     // O_t is still being accumulated by the previous step until this fires
            tc_fence_after();
            // 16 columns at a time: this rare path must not raise the register pressure of the loop (the score row stays live)
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
              uint32_t r[16];
              tmem_ld_x16(tO + cc * 16, r);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
              tmem_st_x16(tO + cc * 16, r);
            }
          }
        }
        // ---- P = exp2(s*c - m_used) -> packed bf16 into this thread's 32 P_t columns ----
        const float2 cc2 = make_float2(c, c);
        float2 mm2 = make_float2(-m_used, -m_used);
        uint32_t pk[2][16];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 x = make_float2(__uint_as_float(s[ch][e]), __uint_as_float(s[ch][e + 1]));
            const float2 a = __ffma2_rn(x, cc2, mm2);
            const float2 ex = (((e >> 1) & 3) < EMU) ? exp2_poly2(a) : make_float2(a2_ex2(a.x), a2_ex2(a.y));
            if (CHAIN > 0 && ((e >> 1) % (CHAIN > 0 ? CHAIN : 1)) == CHAIN - 1 && !(ch == 1 && e + 2 * CHAIN >= 32)) {
              // value-neutral ordering dependency (ex is finite: ex * 0 + mm == mm), see pm_attn.cu
              mm2 = __ffma2_rn(ex, make_float2(0.0f, 0.0f), mm2);
            }
            if ((e >> 1) & 1) lb = __fadd2_rn(lb, ex);
            else la = __fadd2_rn(la, ex);
            pk[ch][e >> 1] = pack_bf16x2(ex.x, ex.y);
          }
        }
        if (j > 0) {
          // the P_t columns are read by P_t V of the previous step until this fires (this late it practically always has)
          mbar_wait_a(b_pv_done, (g - 1) & 1);
          tc_fence_after();
        }
        tmem_st_x16(tP, pk[0]);
        tmem_st_x16(tP + 16, pk[1]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(b_p_full);
      }

      // ---- item epilogue: O / l -> bf16 -> swizzled smem staging -> TMA store; each thread handles 32 of the 64 columns ----
      *l_mine = (la.x + la.y) + (lb.x + lb.y);
      if (tile_leader) tma_store_wait_read<0>();      // previous item's store has left the staging tile
      named_bar_sync(1 + t, 256);                     // partial sums visible; staging tile free
      // fixed order (half 0 + half 1) so that both threads of a row form the same sum
      const float l_sum = kh == 0 ? (*l_mine + *l_other) : (*l_other + *l_mine);
      const float inv_l = 1.0f / l_sum;
      mbar_wait_a(b_pv_done, (g - 1) & 1);
      tc_fence_after();
      const int qrow = it.qb * 2 * A2_BM + t * A2_BM + row_in_tile;
      if (TRAIN && p.lse != nullptr && kh == 0) {
        // training forward: base-2 log-sum-exp of the scaled score row, consumed by pm_attn_bwd
        if (qrow < p.Nq) p.lse[(static_cast<size_t>(it.b) * p.H + it.h) * p.lse_ld + qrow] = m_used + log2f(l_sum);
      }
      uint32_t r0[32];
      tmem_ld_x32(tO + kh * 32, r0);
      tmem_ld_wait();
      // (O_t is free again: the next item's first P V is only issued after BOTH halves' next p_full arrivals)
      if (TRAIN && p.o32 != nullptr) {
        // training forward: an fp32 copy of the output rows (keeps delta = rowsum(dO * O) free of O's bf16 rounding)
        if (qrow < p.Nq) {
          float4* dst = reinterpret_cast<float4*>(p.o32 + (static_cast<size_t>(it.b) * p.Nq + qrow) * p.ldo32 + it.h * A2_D + kh * 32);
#pragma unroll
          for (int jv = 0; jv < 8; ++jv)
            __stcs(dst + jv, make_float4(__uint_as_float(r0[jv * 4 + 0]) * inv_l, __uint_as_float(r0[jv * 4 + 1]) * inv_l,
                                         __uint_as_float(r0[jv * 4 + 2]) * inv_l, __uint_as_float(r0[jv * 4 + 3]) * inv_l));
        }
      }
#pragma unroll
      for (int jv = 0; jv < 4; ++jv) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r0[jv * 8 + 0]) * inv_l, __uint_as_float(r0[jv * 8 + 1]) * inv_l);
        o.y = pack_bf16x2(__uint_as_float(r0[jv * 8 + 2]) * inv_l, __uint_as_float(r0[jv * 8 + 3]) * inv_l);
        o.z = pack_bf16x2(__uint_as_float(r0[jv * 8 + 4]) * inv_l, __uint_as_float(r0[jv * 8 + 5]) * inv_l);
        o.w = pack_bf16x2(__uint_as_float(r0[jv * 8 + 6]) * inv_l, __uint_as_float(r0[jv * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stg + (((kh * 4 + jv) ^ (row_in_tile & 7)) << 4)) = o;
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 256);
      if (tile_leader) {
        asm volatile(
            "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
            ::"l"(reinterpret_cast<uint64_t>(&tmO)),
            "r"(sO + t * A2_TILE_BYTES), "r"(it.h * A2_D), "r"(it.qb * 2 * A2_BM + t * A2_BM), "r"(it.b)
            : "memory");
        tma_store_commit();
      }
    }
    if (tile_leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, A2_TMEM_COLS);
  }
}

using Attn2KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);

struct Attn2Variant {
  int emu, chain;
  Attn2KernelFn fn;
};
// The first entry is the default.  PM_ATTN2_VARIANT="emu,chain" picks another one (tuning aid; every variant computes the
// same function).
static const Attn2Variant kAttn2Variants[] = {
    {1, 8, attn2_kernel<1, 8, false>},
    {1, 0, attn2_kernel<1, 0, false>},
    {1, 4, attn2_kernel<1, 4, false>},
    {2, 8, attn2_kernel<2, 8, false>},
    {2, 0, attn2_kernel<2, 0, false>},
    {0, 0, attn2_kernel<0, 0, false>},
};

int pm_attn2_launch(const AttnParams& p, cudaStream_t stream) {
  if (p.q == nullptr || p.k == nullptr || p.v == nullptr || p.o == nullptr) return PM_ERR_INVALID;
  if (p.B <= 0 || p.H <= 0 || p.Nq <= 0 || p.Nk <= 0 || p.head_dim != A2_D) return PM_ERR_INVALID;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  const uint64_t inner = static_cast<uint64_t>(p.H) * A2_D;
  if ((rc = pm_make_tmap_3d(&tmQ, p.q, 2, p.B, p.Nq, inner, p.ldq, p.bsq, A2_BM, A2_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmK, p.k, 2, p.B, p.Nk, inner, p.ldk, p.bsk, A2_BN, A2_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmV, p.v, 2, p.B, p.Nk, inner, p.ldv, p.bsv, A2_BN, A2_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmO, p.o, 2, p.B, p.Nq, inner, p.ldo, p.bso, A2_BM, A2_D)) != PM_OK) return rc;
  static Attn2KernelFn fn = nullptr;
  if (fn == nullptr) {
    const Attn2Variant* v = &kAttn2Variants[0];
    const char* env = getenv("PM_ATTN2_VARIANT");
    if (env != nullptr) {
      int e = -9, ch = -9;
      if (sscanf(env, "%d,%d", &e, &ch) != 2) return PM_ERR_INVALID;
      v = nullptr;
      for (const Attn2Variant& c : kAttn2Variants)
        if (c.emu == e && c.chain == ch) v = &c;
      if (v == nullptr) return PM_ERR_INVALID;
    }
    fn = v->fn;
  }
  static bool attr_done[PM_MAX_DEVICES] = {}, attr_done_train[PM_MAX_DEVICES] = {};
  const bool train = p.lse != nullptr || p.o32 != nullptr;
  Attn2KernelFn kern = fn;
  if (train) {
    kern = attn2_kernel<1, 8, true>;
    if ((rc = pm_ensure_dyn_smem(kern, A2_SMEM_BYTES, attr_done_train)) != 0) return rc;
  } else if ((rc = pm_ensure_dyn_smem(fn, A2_SMEM_BYTES, attr_done)) != 0) {
    return rc;
  }
  const long long items = static_cast<long long>((p.Nq + 2 * A2_BM - 1) / (2 * A2_BM)) * p.H * p.B;
  const int grid = items < pm_num_sms() ? static_cast<int>(items) : pm_num_sms();
  kern<<<grid, A2_THREADS, A2_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
