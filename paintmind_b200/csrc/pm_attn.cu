// pm_attn.cu — flash-style multi-head attention on tcgen05 / TMEM (sm_100a), head_dim = 64.
//
// Replaces CrossAttention.forward / MemoryEfficientCrossAttention.forward between the q/k/v
// projections and to_out (reference modules/attention.py:51-58 and :84-106):
//     out[b, n, h*64:(h+1)*64] = softmax(scale * Q_bh K_bh^T) V_bh        (no mask, no bias)
// Q/K/V are read in place from the token-major projection outputs ([B, N, ld] with the head at
// column h*64) through 3-D TMA tensor maps, so no '(b h) n d' rearrange is ever materialised;
// O is written token-major, ready for the to_out GEMM.
//
// Persistent: one CTA per SM walks work items (256-query block, head, batch); the stream of (item, key tile)
// steps is flattened so that the next item's Q/K loads and first Q K^T overlap the current item's epilogue.
// 384 threads (register file rebalanced
// with setmaxnreg: 216 per softmax thread, 64 for the TMA / MMA warpgroup):
//   warps 0-3 / 4-7 : softmax warpgroup 0 / 1 — one thread per query row of Q tile 0 / 1:
//                     fp32 online max / sum with lazy rescaling of the TMEM-resident O accumulator,
//                     exp2 on the pre-scaled scores, P written back over S as packed bf16
//   warp 8 lane 0   : TMA producer — both Q tiles once, K and V tiles of 128 keys in 3-stage rings
//   warp 9 lane 0   : MMA issuer 1 — S_t = Q_t K^T (SS, 128x128x64), issued as soon as the previous S_t has been
//                     read into registers
//   warp 10 lane 0  : MMA issuer 2 — O_t += P_t V (TS: P from TMEM, V as MN-major smem operand, 128x64x128)
// TMEM (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
// P lives in its own columns, so S_t is handed back to the MMA warp as soon as the softmax threads have the score
// row in registers: Q_t K(j+1)^T runs underneath the exponentials of tile j and both warpgroups stay busy.
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <stdio.h>
#include <stdlib.h>

namespace pm {

constexpr int AT_BM = 128;      // queries per tile (two tiles per work item)
constexpr int AT_BN = 128;      // keys per tile
constexpr int AT_D = 64;        // head dim
constexpr int AT_TILE_BYTES = 128 * 64 * 2;   // 16 KB (Q, K, V and O tiles alike)
constexpr int AT_KV_STAGES = 3;
constexpr int AT_Q_STAGES = 2;
constexpr int AT_THREADS = 384;            // 3 warpgroups: softmax 0, softmax 1, {TMA, MMA, 2 idle warps}
constexpr int AT_TMEM_COLS = 512;
// smem: Q [2 stages][2 tiles] | K [3] | V [3] | O staging [2 tiles] | barriers
constexpr int AT_SMEM_BYTES = 1024 + (2 * AT_Q_STAGES + 2 * AT_KV_STAGES + 2) * AT_TILE_BYTES + 512;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnItem {
  int qb, h, b;
};
// work item w -> (query block, head, batch); consecutive items share (b, h) so K/V stay hot in L2
__device__ __forceinline__ AttnItem attn_item(int w, int n_qb, int H) {
  AttnItem it;
  it.qb = w % n_qb;
  const int r = w / n_qb;
  it.h = r % H;
  it.b = r / H;
  return it;
}

// A never-taken, compiler-opaque branch: ends the basic block so that ptxas schedules the exponentials of one
// sub-block (MUFU + FMA-pipe mix) before the next one instead of hoisting every polynomial chain to the front and
// leaving a MUFU-only tail (measured: the two softmax warps of an SM sub-partition then fight for the 4-lane MUFU
// while the FMA pipe idles, and vice versa).
__device__ __forceinline__ void sched_fence(uint32_t tok) {
  uint32_t v;
  asm volatile("mov.b32 %0, %1;" : "=r"(v) : "r"(tok));
  if (v == 0xffffffffu) __trap();
}

// EMU   : of every 4 pairs of scores, this many take the FMA-pipe exp2
// SPLIT : sub-blocks of the exponential phase separated by scheduling fences (1 = none, 4 = per 32 keys, 8 = per 16)
// CHAIN : > 0: groups of CHAIN score pairs are chained by a value-neutral dependency (see the exponential loop)
// PING  : 2 = the exponential phases of the two softmax warps that share an SM sub-partition (warp q of Q tile 0 and
//         warp q of Q tile 1) are made mutually exclusive with a pair of named barriers (ping-pong): measured, the two
//         run in lock-step otherwise (the MMA warps hand S_0 and S_1 out back to back), so both sit in their MUFU-bound
//         phase at the same time (MUFU saturated) and both in their load / max / store phases at the same time (MUFU
//         idle).  1 = a one-time half-step delay of tile 1 instead.  0 = neither.
// DEFER : where the wait for the previous P_t V sits: -1 = before the exponentials, c = 0..3 = after the
//         exponentials of 32-key chunk c (P chunks are stored from there on; 3 = whole row kept in registers)
// TRAIN : also emit the row log-sum-exp and an fp32 copy of the output (training forward; a separate instantiation so that
//         the inference kernel's register allocation is untouched: the two epilogue branches cost it 1.5 %)
template <int EMU, int SPLIT, int DEFER, int CHAIN, int PING, bool TRAIN = false>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
            const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // everything is addressed by 32-bit shared::cta addresses (no generic-window rebuilds in the hot loops)
  const uint32_t raw_a = smem_u32(smem_raw);
  const uint32_t sQ = (raw_a + 1023u) & ~1023u;                          // [AT_Q_STAGES][2][16 KB]
  const uint32_t sK = sQ + 2 * AT_Q_STAGES * AT_TILE_BYTES;              // [AT_KV_STAGES][16 KB]
  const uint32_t sV = sK + AT_KV_STAGES * AT_TILE_BYTES;                 // [AT_KV_STAGES][16 KB]
  const uint32_t sO = sV + AT_KV_STAGES * AT_TILE_BYTES;                 // [2][16 KB] output staging
  const uint32_t bars = sO + 2 * AT_TILE_BYTES;
  const uint32_t q_full = bars;                                   // [2]
  const uint32_t q_empty = q_full + 8 * AT_Q_STAGES;              // [2]
  const uint32_t k_full = q_empty + 8 * AT_Q_STAGES;              // [3]
  const uint32_t k_empty = k_full + 8 * AT_KV_STAGES;             // [3]
  const uint32_t v_full = k_empty + 8 * AT_KV_STAGES;             // [3]
  const uint32_t v_empty = v_full + 8 * AT_KV_STAGES;             // [3]
  const uint32_t s_full = v_empty + 8 * AT_KV_STAGES;             // [2]  MMA -> softmax: S_t complete
  const uint32_t s_free = s_full + 16;                            // [2]  softmax -> MMA: S_t is in registers
  const uint32_t p_full = s_free + 16;                            // [2]  softmax -> MMA: P_t written (and O_t rescaled)
  const uint32_t pv_done = p_full + 16;                           // [2]  MMA -> softmax: O_t += P_t V complete
  const uint32_t tmem_slot_a = pv_done + 16;
  uint8_t* const smO = smem_raw + (sO - raw_a);                   // generic view of the staging tiles
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot_a - raw_a));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = (p.Nk + AT_BN - 1) / AT_BN;
  const int n_qb = (p.Nq + 2 * AT_BM - 1) / (2 * AT_BM);
  const int total_items = n_qb * p.H * p.B;
  // persistent: this CTA handles items blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_items = static_cast<int>(blockIdx.x) < total_items
                           ? (total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                           : 0;
  const int total_steps = my_items * n_kv;       // flattened stream of (item, key tile) steps

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    for (int i = 0; i < AT_Q_STAGES; ++i) {
      init(q_full + 8 * i, 1);
      init(q_empty + 8 * i, 1);
    }
    for (int i = 0; i < AT_KV_STAGES; ++i) {
      init(k_full + 8 * i, 1);
      init(k_empty + 8 * i, 1);
      init(v_full + 8 * i, 1);
      init(v_empty + 8 * i, 1);
    }
    for (int t = 0; t < 2; ++t) {
      init(s_full + 8 * t, 1);
      init(s_free + 8 * t, 4);      // one arrival per softmax warp of the group
      init(p_full + 8 * t, 4);
      init(pv_done + 8 * t, 1);
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
        int g = 0;                                   // running key-tile counter (K/V ring position)
        for (int i = 0; i < my_items; ++i) {
          const AttnItem it = attn_item(blockIdx.x + i * gridDim.x, n_qb, p.H);
          const int qs = i % AT_Q_STAGES;
          mbar_wait_a(q_empty + 8 * qs, ((i / AT_Q_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx_a(q_full + 8 * qs, 2 * AT_TILE_BYTES);
          tma_load_3d_a(sQ + (2 * qs) * AT_TILE_BYTES, &tmQ, q_full + 8 * qs, it.h * AT_D, it.qb * 2 * AT_BM, it.b);
          tma_load_3d_a(sQ + (2 * qs + 1) * AT_TILE_BYTES, &tmQ, q_full + 8 * qs, it.h * AT_D, it.qb * 2 * AT_BM + AT_BM, it.b);
          for (int j = 0; j < n_kv; ++j, ++g) {
            const int st = g % AT_KV_STAGES;
            const uint32_t ph = ((g / AT_KV_STAGES) & 1) ^ 1;
            mbar_wait_a(k_empty + 8 * st, ph);
            mbar_arrive_expect_tx_a(k_full + 8 * st, AT_TILE_BYTES);
            tma_load_3d_a(sK + st * AT_TILE_BYTES, &tmK, k_full + 8 * st, it.h * AT_D, j * AT_BN, it.b);
            mbar_wait_a(v_empty + 8 * st, ph);
            mbar_arrive_expect_tx_a(v_full + 8 * st, AT_TILE_BYTES);
            tma_load_3d_a(sV + st * AT_TILE_BYTES, &tmV, v_full + 8 * st, it.h * AT_D, j * AT_BN, it.b);
          }
        }
      }
    } else if (warp == 9) {
      // ===================================== MMA issuer 1: S_t = Q_t K^T ========================
      // Two issuer threads (this one and warp 10) so that a Q K^T is never queued behind a wait for P:
      // S, P and O live in disjoint TMEM columns, the two MMA streams are independent.
      if (lane == 0 && total_steps > 0) {
        constexpr uint32_t idesc_qk = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);   // Q, K both K-major
        const uint32_t tS[2] = {tmem_base, tmem_base + 128};
        for (int g = 0; g < total_steps; ++g) {
          const int i = g / n_kv, j = g - i * n_kv;
          const int qs = i % AT_Q_STAGES, ks = g % AT_KV_STAGES;
          if (j == 0) mbar_wait_a(q_full + 8 * qs, (i / AT_Q_STAGES) & 1);
          mbar_wait_a(k_full + 8 * ks, (g / AT_KV_STAGES) & 1);
          const uint64_t dk = umma_desc_sw128(sK + ks * AT_TILE_BYTES);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            // S_t of the previous step must be in the softmax threads' registers before it is overwritten
            if (g > 0) mbar_wait_a(s_free + 8 * t, (g - 1) & 1);
            tc_fence_after();
            const uint64_t dq = umma_desc_sw128(sQ + (2 * qs + t) * AT_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < AT_D / 16; ++k) umma_ss(tS[t], dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
            umma_commit_a(s_full + 8 * t);
          }
          umma_commit_a(k_empty + 8 * ks);                            // both tiles have used K(g)
          if (j == n_kv - 1) umma_commit_a(q_empty + 8 * qs);         // last use of this item's Q
        }
      }
    } else if (warp == 10) {
      // ===================================== MMA issuer 2: O_t (+)= P_t V ========================
      if (lane == 0 && total_steps > 0) {
        constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_BM, AT_D, 0, 1);    // P K-major (TMEM), V MN-major
        const uint32_t tP[2] = {tmem_base + 256, tmem_base + 320};
        const uint32_t tO[2] = {tmem_base + 384, tmem_base + 448};
        for (int g = 0; g < total_steps; ++g) {
          const int j = g % n_kv;
          const int vs = g % AT_KV_STAGES;
          mbar_wait_a(v_full + 8 * vs, (g / AT_KV_STAGES) & 1);
          const uint64_t dv = umma_desc_sw128(sV + vs * AT_TILE_BYTES);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait_a(p_full + 8 * t, g & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < AT_BN / 16; ++kk) {
              // A: 16 keys = 8 TMEM columns of packed bf16;  B: 16 key rows x 128 B = 2048 B
              umma_ts(tO[t], tP[t] + 8 * kk, dv + kk * (2048 >> 4), idesc_pv, (j | kk) != 0 ? 1u : 0u);
            }
            umma_commit_a(pv_done + 8 * t);
          }
          umma_commit_a(v_empty + 8 * vs);
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
    const uint32_t tP = tmem_base + 256 + t * 64 + lane_off;
    const uint32_t tO = tmem_base + 384 + t * 64 + lane_off;
    const uint32_t b_s_full = s_full + 8 * t, b_s_free = s_free + 8 * t;
    const uint32_t b_p_full = p_full + 8 * t, b_pv_done = pv_done + 8 * t;
    const float c = p.scale_log2;                  // softmax scale * log2(e)
    uint8_t* stg_base = smO + t * AT_TILE_BYTES;
    uint8_t* stg = stg_base + row_in_tile * 128;
    int g = 0;                                     // flattened step counter (barrier phases)
    if (PING == 2 && t == 1 && total_steps > 0) asm volatile("bar.arrive %0, 64;" ::"r"(3 + 2 * q) : "memory");   // tile 0 goes first
    if (PING == 1 && t == 1) {
      const long long t0 = clock64();
      while (clock64() - t0 < 900) {
      }
    }

    for (int i = 0; i < my_items; ++i) {
      const AttnItem it = attn_item(blockIdx.x + i * gridDim.x, n_qb, p.H);
      float m_used = -INFINITY;                    // running max (scaled, log2 domain) the accumulators refer to
      float2 la = make_float2(0.0f, 0.0f);         // running sum of exp2(s*c - m_used): four partial sums
      float2 lb = make_float2(0.0f, 0.0f);         //   (two packed accumulators = two short dependency chains)

      for (int j = 0; j < n_kv; ++j, ++g) {
        const int valid = p.Nk - j * AT_BN;        // >= 128 for full tiles
        mbar_wait_a(b_s_full, g & 1);
        tc_fence_after();
        // ---- whole score row into registers ----
        uint32_t s[4][32];
        tmem_ld_x32(tS + 0, s[0]);
        tmem_ld_x32(tS + 32, s[1]);
        tmem_ld_x32(tS + 64, s[2]);
        tmem_ld_x32(tS + 96, s[3]);
        tmem_ld_wait();
        // the score row is in registers: S_t may be overwritten by the next Q K^T
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(b_s_free);
        if (valid < AT_BN) {                       // ragged last key tile (e.g. 77 text tokens): mask the tail
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (ch * 32 + e >= valid) s[ch][e] = 0xff800000u;      // -inf
        }
        // ---- row max (3-input max) ----
        float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // independent chains (latency, not throughput)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
          for (int e = 0; e < 32; e += 2) mp[ch] = fmaxf(mp[ch], fmaxf(__uint_as_float(s[ch][e]), __uint_as_float(s[ch][e + 1])));
        const float mt = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
        const float m_new = fmaxf(m_used, mt * c);
        // lazy rescale: keep the stale max while it is within 2^8 of the true one (exact algebra,
        // bounded magnitude); decided per warp to keep the TMEM traffic warp-uniform
        const bool need = (m_new - m_used) > 8.0f;
        if (DEFER < 0 && j > 0) {
          // O_t and P_t are still being read by P_t V of the previous step until this fires
          mbar_wait_a(b_pv_done, (g - 1) & 1);
          tc_fence_after();
        }
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2_approx(m_used - m_new) : 1.0f;
          if (need) m_used = m_new;
          la.x *= alpha;
          la.y *= alpha;
          lb.x *= alpha;
          lb.y *= alpha;
          if (j > 0) {
            if (DEFER >= 0) {
              // O_t is still being accumulated by P_t V of the previous step until this fires
              mbar_wait_a(b_pv_done, (g - 1) & 1);
              tc_fence_after();
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t r[32];
              tmem_ld_x32(tO + cc * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
              tmem_st_x32(tO + cc * 32, r);
            }
          }
        }
        // ---- P = exp2(s*c - m_used) -> packed bf16 into the P_t columns ----
        if (PING == 2) asm volatile("bar.sync %0, 64;" ::"r"(3 + 2 * q + t) : "memory");      // my turn on the MUFU
        const float2 cc2 = make_float2(c, c);
        float2 mm2 = make_float2(-m_used, -m_used);
        uint32_t pk[4][16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            if (SPLIT == 8 && e == 16) sched_fence(static_cast<uint32_t>(g));
            const float2 x = make_float2(__uint_as_float(s[ch][e]), __uint_as_float(s[ch][e + 1]));
            const float2 a = __ffma2_rn(x, cc2, mm2);
            const float2 ex = (((e >> 1) & 3) < EMU) ? exp2_poly2(a) : make_float2(ex2_approx(a.x), ex2_approx(a.y));
            if (CHAIN > 0 && ((e >> 1) % (CHAIN > 0 ? CHAIN : 1)) == CHAIN - 1 && !(ch == 3 && e + 2 * CHAIN >= 32)) {
              // Ordering dependency (value-neutral: ex is finite, ex * 0 + mm == mm): the next group's scores are scaled
              // with an offset that "depends" on a MUFU result of this group, so ptxas cannot hoist every polynomial
              // chain of the row to the front and leave a MUFU-only tail; each group keeps its own MUFU / FMA mix.
              mm2 = __ffma2_rn(ex, make_float2(0.0f, 0.0f), mm2);
            }
            if ((e >> 1) & 1) lb = __fadd2_rn(lb, ex);
            else la = __fadd2_rn(la, ex);
            pk[ch][e >> 1] = pack_bf16x2(ex.x, ex.y);
          }
          if (DEFER >= 0 && ch == DEFER) {
            if (j > 0) {
              // the P_t columns are read by P_t V of the previous step until this fires — this late it practically
              // always has (the wait used to sit in front of the exponentials and cost ~11 % of the softmax time)
              mbar_wait_a(b_pv_done, (g - 1) & 1);
              tc_fence_after();
            }
#pragma unroll
            for (int c2 = 0; c2 < ch; ++c2) tmem_st_x16(tP + c2 * 16, pk[c2]);
          }
          if (DEFER < 0 || ch >= DEFER) tmem_st_x16(tP + ch * 16, pk[ch]);
          if (SPLIT >= 4 && ch < 3) sched_fence(static_cast<uint32_t>(g));
        }
        if (PING == 2 && !(t == 1 && g == total_steps - 1))
          asm volatile("bar.arrive %0, 64;" ::"r"(3 + 2 * q + (t ^ 1)) : "memory");           // the partner warp's turn
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(b_p_full);
      }

      // ---- item epilogue: O / l -> bf16 -> swizzled smem staging -> TMA store ----
      mbar_wait_a(b_pv_done, (g - 1) & 1);
      tc_fence_after();
      const float l_sum = (la.x + la.y) + (lb.x + lb.y);
      const float inv_l = 1.0f / l_sum;
      if (TRAIN && p.lse != nullptr) {
        // training forward: base-2 log-sum-exp of the scaled score row, consumed by pm_attn_bwd
        const int qrow = it.qb * 2 * AT_BM + t * AT_BM + row_in_tile;
        if (qrow < p.Nq) p.lse[(static_cast<size_t>(it.b) * p.H + it.h) * p.lse_ld + qrow] = m_used + log2f(l_sum);
      }
      uint32_t r0[32], r1[32];
      tmem_ld_x32(tO, r0);
      tmem_ld_x32(tO + 32, r1);
      tmem_ld_wait();
      // (O_t is free again: the next item's first P V is only issued after this thread's next p_full arrival)
      if (TRAIN && p.o32 != nullptr) {
        // training forward: an fp32 copy of the output rows, so that delta = rowsum(dO * O) of the backward pass is not
        // limited by O's bf16 rounding (dS = P (dP - delta) cancels to ~1e-3 of its terms when attention is diffuse)
        const int qrow = it.qb * 2 * AT_BM + t * AT_BM + row_in_tile;
        if (qrow < p.Nq) {
          float4* dst = reinterpret_cast<float4*>(p.o32 + (static_cast<size_t>(it.b) * p.Nq + qrow) * p.ldo32 + it.h * AT_D);
#pragma unroll
          for (int jv = 0; jv < 8; ++jv) {
            __stcs(dst + jv, make_float4(__uint_as_float(r0[jv * 4 + 0]) * inv_l, __uint_as_float(r0[jv * 4 + 1]) * inv_l,
                                         __uint_as_float(r0[jv * 4 + 2]) * inv_l, __uint_as_float(r0[jv * 4 + 3]) * inv_l));
            __stcs(dst + 8 + jv, make_float4(__uint_as_float(r1[jv * 4 + 0]) * inv_l, __uint_as_float(r1[jv * 4 + 1]) * inv_l,
                                             __uint_as_float(r1[jv * 4 + 2]) * inv_l, __uint_as_float(r1[jv * 4 + 3]) * inv_l));
          }
        }
      }
      if (q == 0 && lane == 0) tma_store_wait_read<0>();      // previous item's store has left the staging tile
      named_bar_sync(1 + t, 128);
#pragma unroll
      for (int jv = 0; jv < 4; ++jv) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r0[jv * 8 + 0]) * inv_l, __uint_as_float(r0[jv * 8 + 1]) * inv_l);
        o.y = pack_bf16x2(__uint_as_float(r0[jv * 8 + 2]) * inv_l, __uint_as_float(r0[jv * 8 + 3]) * inv_l);
        o.z = pack_bf16x2(__uint_as_float(r0[jv * 8 + 4]) * inv_l, __uint_as_float(r0[jv * 8 + 5]) * inv_l);
        o.w = pack_bf16x2(__uint_as_float(r0[jv * 8 + 6]) * inv_l, __uint_as_float(r0[jv * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stg + ((jv ^ (row_in_tile & 7)) << 4)) = o;
      }
#pragma unroll
      for (int jv = 0; jv < 4; ++jv) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r1[jv * 8 + 0]) * inv_l, __uint_as_float(r1[jv * 8 + 1]) * inv_l);
        o.y = pack_bf16x2(__uint_as_float(r1[jv * 8 + 2]) * inv_l, __uint_as_float(r1[jv * 8 + 3]) * inv_l);
        o.z = pack_bf16x2(__uint_as_float(r1[jv * 8 + 4]) * inv_l, __uint_as_float(r1[jv * 8 + 5]) * inv_l);
        o.w = pack_bf16x2(__uint_as_float(r1[jv * 8 + 6]) * inv_l, __uint_as_float(r1[jv * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stg + (((4 + jv) ^ (row_in_tile & 7)) << 4)) = o;
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if (q == 0 && lane == 0) {
        asm volatile(
            "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
            ::"l"(reinterpret_cast<uint64_t>(&tmO)),
            "r"(sO + t * AT_TILE_BYTES), "r"(it.h * AT_D), "r"(it.qb * 2 * AT_BM + t * AT_BM), "r"(it.b)
            : "memory");
        tma_store_commit();
      }
    }
    if (q == 0 && lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

using AttnKernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);

struct AttnVariant {
  int emu, split, defer, chain, ping;
  AttnKernelFn fn;
};
#define PM_ATTN_V(E, S, D, C, P) {E, S, D, C, P, attn_kernel<E, S, D, C, P>}
// The first entry is the default.  PM_ATTN_VARIANT="emu,split,defer,chain,ping" picks another one (tuning aid; every
// variant computes the same function).
static const AttnVariant kAttnVariants[] = {
    // measured on B200 (B = 256, H = 8, N = 1024; scripts/attn_variants.py, profiles/r01_attn_variants.txt):
    PM_ATTN_V(1, 1, 3, 8, 0),    // 0.656 ms  default: deferred P_t V wait, groups of 8 pairs chained
    PM_ATTN_V(1, 1, -1, 0, 0),   // 0.684-0.690 ms  the previous kernel (wait in front of the exponentials)
    PM_ATTN_V(1, 1, 3, 0, 0),    // 0.661-0.668 ms  deferred wait only
    PM_ATTN_V(1, 1, 3, 8, 2),    // 0.661 ms  + ping-pong barriers between the two warps of a sub-partition (no gain)
    PM_ATTN_V(2, 1, 3, 8, 0),    //           half of the exponentials on the FMA pipe
    PM_ATTN_V(0, 1, 3, 0, 0),    // 0.736 ms  every exponential on the MUFU
};
#undef PM_ATTN_V

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
  static AttnKernelFn fn = nullptr;
  if (fn == nullptr) {
    const AttnVariant* v = &kAttnVariants[0];
    const char* env = getenv("PM_ATTN_VARIANT");
    if (env != nullptr) {
      int e = -9, s = -9, d = -9, ch = -9, pg = -9;
      if (sscanf(env, "%d,%d,%d,%d,%d", &e, &s, &d, &ch, &pg) != 5) return PM_ERR_INVALID;
      v = nullptr;
      for (const AttnVariant& c : kAttnVariants)
        if (c.emu == e && c.split == s && c.defer == d && c.chain == ch && c.ping == pg) v = &c;
      if (v == nullptr) return PM_ERR_INVALID;
    }
    fn = v->fn;
  }
  static bool attr_done[PM_MAX_DEVICES] = {}, attr_done_train[PM_MAX_DEVICES] = {};
  const bool train = p.lse != nullptr || p.o32 != nullptr;
  AttnKernelFn kern = fn;
  if (train) {
    kern = attn_kernel<1, 1, 3, 8, 0, true>;
    if ((rc = pm_ensure_dyn_smem(kern, AT_SMEM_BYTES, attr_done_train)) != 0) return rc;
  } else if ((rc = pm_ensure_dyn_smem(fn, AT_SMEM_BYTES, attr_done)) != 0) {
    return rc;
  }
  const long long items = static_cast<long long>((p.Nq + 2 * AT_BM - 1) / (2 * AT_BM)) * p.H * p.B;
  const int grid = items < pm_num_sms() ? static_cast<int>(items) : pm_num_sms();
  kern<<<grid, AT_THREADS, AT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
