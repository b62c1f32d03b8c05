// pm_attn3.cu — forward attention for PRE-SCALED queries (sm_100a, head_dim = 64): the row maximum is subtracted by the
// tensor core, the softmax threads only exponentiate.
//
// Same function as pm_attn.cu (reference modules/attention.py:51-58 / :84-106):
//     out[b, n, h*64:(h+1)*64] = softmax(scale * Q_bh K_bh^T) V_bh
// with the contract that the caller's Q already carries scale * log2(e) (engine.py folds it into the to_q rows of the packed
// q|k|v weight, so it costs nothing at run time): S = Q K^T is then the base-2 logit.
//
// Why: at head_dim 64 the forward kernel is bound by the softmax threads' instruction stream, not by the tensor pipe
// (profiles/r02_attention.md: 614 warp-instructions per 128-key step and warp, of which 96 MUFU; issue 56 %, XU 60 %,
// tensor 39 %).  Of those 614, 135 are the per-score `s * c - m` FFMA2s and 99 are the row-maximum pass.  Here
//   * the MMA warp issues ONE more K = 16 step per score tile: S_t = Q_t K^T + A_t B^T, where A_t [128 x 16] holds
//     (-m_row, 1, 0, ...) per query row (written by the softmax thread that owns the row, 4 bytes, only when its m changes)
//     and B [128 keys x 16] is a constant (1, mask, 0, ...): the scores arrive as s - m, with masked keys (ragged last
//     tile) at -3.4e38 — no scaling FFMA2, no masking selects;
//   * m is fixed by the first key tile of a work item (its exact row maximum, rounded to bf16 so that A holds it exactly):
//     tiles 1.. take no maximum at all.  exp2 of fp32 has 127 binades of head-room above it, and P, l and O are scale-free,
//     so a stale maximum costs no accuracy.
// Robustness: a row with a logit more than ~2^100 above its first tile's maximum would overflow; this is detected from the
// row sum in the item epilogue (> 1e30, inf or NaN), the item is marked in a shared-memory bitmap, and after its last item the
// CTA re-runs the marked items in EXACT mode (maximum of every tile with lazy rescaling, as pm_attn.cu does).  Tested with
// crafted inputs (tests/test_gpu_attention3.py); never taken on real activations.
//
// Thread roles, TMEM map, K/V rings and the epilogue are those of pm_attn.cu.
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <stdio.h>
#include <stdlib.h>

namespace pm {

constexpr int A3_BM = 128;
constexpr int A3_BN = 128;
constexpr int A3_D = 64;
constexpr int A3_TILE_BYTES = 128 * 64 * 2;
constexpr int A3_KV_STAGES = 3;
constexpr int A3_Q_STAGES = 2;
constexpr int A3_THREADS = 384;
constexpr int A3_TMEM_COLS = 512;
constexpr int A3_BIAS_BYTES = 128 * 16 * 2;       // one [128 x 16] bf16 K-major no-swizzle operand: 2 k-chunks x 128 rows x 16 B
constexpr int A3_MAX_ITEMS = 4096;                // per CTA (bitmap of items to re-run in exact mode)
// smem: Q [2 stages][2 tiles] | K [3] | V [3] | O staging [2] | A_0 A_1 B_full B_last | redo bitmap | barriers
constexpr int A3_SMEM_BYTES = 1024 + (2 * A3_Q_STAGES + 2 * A3_KV_STAGES + 2) * A3_TILE_BYTES + 4 * A3_BIAS_BYTES + A3_MAX_ITEMS / 8 + 512;

__device__ __forceinline__ float a3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major operand without swizzle, K = 16 (two 8-element chunks): row r of chunk c at  c * 2048 + r * 16  bytes
// (canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units with SBO = 8 rows x 16 B = 128 B, LBO = 2048 B).
__device__ __forceinline__ uint64_t a3_desc_k16(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(2048 >> 4) << 16;      // LBO: between the two k-chunks
  d |= static_cast<uint64_t>(128 >> 4) << 32;       // SBO: between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;
  return d;                                         // swizzle mode 0
}

#ifdef PM_A3_DEBUG
// bring-up cycle counters (scripts/attn3_debug.py; build with PM_NVCC_EXTRA=-DPM_A3_DEBUG).  Lane 0 of softmax warps 0 / 4, the
// two / three issuer threads and the producer add their per-phase cycles; see the script for the slot names.
__device__ unsigned long long g_a3_dbg[32];
#define A3_DBG_DECL unsigned long long dbg_loc[32] = {}; long long dbg_t = clock64();
#define A3_DBG_LAP(i) { const long long now_ = clock64(); dbg_loc[i] += static_cast<unsigned long long>(now_ - dbg_t); dbg_t = now_; }
#define A3_DBG_FLUSH                                   \
  for (int di = 0; di < 32; ++di)                      \
    if (dbg_loc[di] != 0ull) atomicAdd(&g_a3_dbg[di], dbg_loc[di]);
#else
#define A3_DBG_DECL
#define A3_DBG_LAP(i)
#define A3_DBG_FLUSH
#endif

struct A3Item {
  int qb, h, b;
};
__device__ __forceinline__ A3Item a3_item(int w, int n_qb, int H) {
  A3Item it;
  it.qb = w % n_qb;
  const int r = w / n_qb;
  it.h = r % H;
  it.b = r / H;
  return it;
}

// exp2_poly2 (pm_common.cuh) clamps from below only: above 2^127 its exponent arithmetic wraps to garbage that the overflow
// check of this kernel could miss (a small negative number instead of inf).  Clamp at 128: 2^128 has the bit pattern of inf /
// NaN, which is what the MUFU lanes produce and what the end-of-item check looks for.
__device__ __forceinline__ float2 a3_exp2_poly2(float2 a) {
  a.x = fminf(a.x, 128.0f);
  a.y = fminf(a.y, 128.0f);
  return exp2_poly2(a);
}

__device__ __forceinline__ float a3_bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// The exponential phase of one 128-key step: P = exp2(s [+ delta]) -> packed bf16, row sums into la / lb, P stored to TMEM once
// P_t V of the previous step has finished with the P columns.
//   SHIFT : add `delta` to every score first (first tile of an item, exact mode, the rare re-centring step)
//   EMU   : of every 4 score pairs, this many take the FMA-pipe exp2
//   DEFER : the wait for the previous P_t V (and the first P stores) sits after the exponentials of 32-key chunk DEFER
// Both barrier polls a step needs (P_t V of the previous step done; the next score tile complete) are ISSUED here up front and
// consumed later — a try_wait costs the warp ~100-200 cycles even when the phase completed long ago.  Returns the poll of the
// next score tile.  (Measured and dropped: refilling the score registers of finished chunks with the next tile's scores under
// the remaining exponentials — tcgen05.ld stalls the issuing warp for its ~100 cycles either way: 0.72-0.76 vs 0.64 ms.)
template <bool SHIFT, int EMU, int DEFER>
__device__ __forceinline__ uint32_t a3_exps(uint32_t (&s)[4][32], float delta, float2& la, float2& lb, uint32_t tP, uint32_t b_pv_done,
                                            bool wait_pv, uint32_t pv_parity, bool poll_next, uint32_t b_s_full, uint32_t s_parity,
                                            unsigned long long* dbg) {
#ifdef PM_A3_DEBUG
  long long dbg_t = clock64();
#define A3X_LAP(i) { const long long now_ = clock64(); dbg[i] += static_cast<unsigned long long>(now_ - dbg_t); dbg_t = now_; }
#else
#define A3X_LAP(i)
#endif
  const float2 dd2 = make_float2(delta, delta);
  uint32_t ok_pv = 1, ok_s = 0;
  if (wait_pv) ok_pv = mbar_try_wait_a(b_pv_done, pv_parity);
  uint32_t pk[4][16];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      float2 a = make_float2(__uint_as_float(s[ch][e]), __uint_as_float(s[ch][e + 1]));
      if (SHIFT) a = __fadd2_rn(a, dd2);
      const bool poly = ((e >> 1) & 3) < EMU;
      const float2 ex = poly ? a3_exp2_poly2(a) : make_float2(a3_ex2(a.x), a3_ex2(a.y));
      if ((e >> 1) & 1) lb = __fadd2_rn(lb, ex);
      else la = __fadd2_rn(la, ex);
      pk[ch][e >> 1] = pack_bf16x2(ex.x, ex.y);
    }
    A3X_LAP(27 + (ch < 3 ? ch : 2))
    if (ch == 2 && poll_next) ok_s = mbar_try_wait_a(b_s_full, s_parity);
    if (ch == DEFER) {
      if (wait_pv && !ok_pv) {
        // the P_t columns are read by P_t V of the previous step until this fires (this late it practically always has)
        mbar_wait_a(b_pv_done, pv_parity);
      }
      tc_fence_after();
      A3X_LAP(30)
#pragma unroll
      for (int c2 = 0; c2 < ch; ++c2) tmem_st_x16(tP + c2 * 16, pk[c2]);
    }
    if (ch >= DEFER) tmem_st_x16(tP + ch * 16, pk[ch]);
  }
  A3X_LAP(31)
  return ok_s;
}

template <int EMU, int DEFER, bool PREF, bool PV2>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
             const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_a = smem_u32(smem_raw);
  const uint32_t sQ = (raw_a + 1023u) & ~1023u;
  const uint32_t sK = sQ + 2 * A3_Q_STAGES * A3_TILE_BYTES;
  const uint32_t sV = sK + A3_KV_STAGES * A3_TILE_BYTES;
  const uint32_t sO = sV + A3_KV_STAGES * A3_TILE_BYTES;
  const uint32_t sA = sO + 2 * A3_TILE_BYTES;                     // [2] per-tile (-m, 1, 0...) operands
  const uint32_t sBfull = sA + 2 * A3_BIAS_BYTES;                 // (1, 0, 0...) for every key
  const uint32_t sBlast = sBfull + A3_BIAS_BYTES;                 // (1, key >= valid ? -max : 0, 0...) for the ragged last tile
  const uint32_t sRedo = sBlast + A3_BIAS_BYTES;                  // bitmap [A3_MAX_ITEMS] + flag word behind it
  const uint32_t bars = sRedo + A3_MAX_ITEMS / 8 + 16;
  const uint32_t q_full = bars;
  const uint32_t q_empty = q_full + 8 * A3_Q_STAGES;
  const uint32_t k_full = q_empty + 8 * A3_Q_STAGES;
  const uint32_t k_empty = k_full + 8 * A3_KV_STAGES;
  const uint32_t v_full = k_empty + 8 * A3_KV_STAGES;
  const uint32_t v_empty = v_full + 8 * A3_KV_STAGES;
  const uint32_t s_full = v_empty + 8 * A3_KV_STAGES;
  const uint32_t s_free = s_full + 16;
  const uint32_t p_full = s_free + 16;
  const uint32_t pv_done = p_full + 16;
  const uint32_t tmem_slot_a = pv_done + 16;
  const uint32_t pass_done = tmem_slot_a + 16;                    //         softmax warps -> everyone: pass 0 complete, redo bitmap final
  uint8_t* const smO = smem_raw + (sO - raw_a);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot_a - raw_a));
  uint32_t* const redo_bits = reinterpret_cast<uint32_t*>(smem_raw + (sRedo - raw_a));
  volatile uint32_t* const redo_any = redo_bits + A3_MAX_ITEMS / 32;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = (p.Nk + A3_BN - 1) / A3_BN;
  const int n_qb = (p.Nq + 2 * A3_BM - 1) / (2 * A3_BM);
  const int total_items = n_qb * p.H * p.B;
  const int my_items = static_cast<int>(blockIdx.x) < total_items
                           ? (total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                           : 0;
  const int valid_last = p.Nk - (n_kv - 1) * A3_BN;      // keys of the last tile, 1..128

  // ---- constant / initial bias operands, redo bitmap ----
  for (int i = threadIdx.x; i < A3_MAX_ITEMS / 32 + 1; i += A3_THREADS) redo_bits[i] = 0;
  if (threadIdx.x < 256) {
    // A_t row r: chunk 0 = (-m = 0, 1, 0, ...), chunk 1 = 0
    const uint32_t a = sA + (threadIdx.x >> 7) * A3_BIAS_BYTES + (threadIdx.x & 127) * 16;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(a), "r"(0x3F800000u), "r"(0u) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a + 2048), "r"(0u) : "memory");
  } else {
    const int r = threadIdx.x - 256;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(sBfull + r * 16), "r"(0x00003F80u), "r"(0u) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sBfull + 2048 + r * 16), "r"(0u) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(sBlast + r * 16), "r"(r < valid_last ? 0x00003F80u : 0xFF7F3F80u), "r"(0u)
                 : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sBlast + 2048 + r * 16), "r"(0u) : "memory");
  }
  fence_proxy_async_smem();

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    for (int i = 0; i < A3_Q_STAGES; ++i) {
      init(q_full + 8 * i, 1);
      init(q_empty + 8 * i, 1);
    }
    for (int i = 0; i < A3_KV_STAGES; ++i) {
      init(k_full + 8 * i, 1);
      init(k_empty + 8 * i, 1);
      init(v_full + 8 * i, 1);
      init(v_empty + 8 * i, PV2 ? 2 : 1);
    }
    for (int t = 0; t < 2; ++t) {
      init(s_full + 8 * t, 1);
      init(s_free + 8 * t, 4);
      init(p_full + 8 * t, 4);
      init(pv_done + 8 * t, 1);
    }
    init(pass_done, 8);                   // one arrival per softmax warp
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, A3_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    int ic = 0, g = 0;                               // running item / step counters (barrier phases) across both passes
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) {
        // pass 0 complete in every softmax warp, bitmap final.  One lane per role warp polls (a warp without a role would
        // otherwise poll from the first cycle of the kernel and take issue slots from the softmax warps of its scheduler)
        if (lane == 0 && (warp <= 10 || PV2)) mbar_wait_a(pass_done, 0);
        __syncwarp();
        if (*redo_any == 0) break;
      }
      if (lane == 0) {
        if (warp == 8) {
          // ===================================== TMA producer ======================================
          for (int i = 0; i < my_items; ++i) {
            if (pass == 1 && ((redo_bits[i >> 5] >> (i & 31)) & 1u) == 0) continue;
            const A3Item it = a3_item(blockIdx.x + i * gridDim.x, n_qb, p.H);
            const int qs = ic % A3_Q_STAGES;
            mbar_wait_a(q_empty + 8 * qs, ((ic / A3_Q_STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx_a(q_full + 8 * qs, 2 * A3_TILE_BYTES);
            tma_load_3d_a(sQ + (2 * qs) * A3_TILE_BYTES, &tmQ, q_full + 8 * qs, it.h * A3_D, it.qb * 2 * A3_BM, it.b);
            tma_load_3d_a(sQ + (2 * qs + 1) * A3_TILE_BYTES, &tmQ, q_full + 8 * qs, it.h * A3_D, it.qb * 2 * A3_BM + A3_BM, it.b);
            for (int j = 0; j < n_kv; ++j, ++g) {
              const int st = g % A3_KV_STAGES;
              const uint32_t ph = ((g / A3_KV_STAGES) & 1) ^ 1;
              mbar_wait_a(k_empty + 8 * st, ph);
              mbar_arrive_expect_tx_a(k_full + 8 * st, A3_TILE_BYTES);
              tma_load_3d_a(sK + st * A3_TILE_BYTES, &tmK, k_full + 8 * st, it.h * A3_D, j * A3_BN, it.b);
              mbar_wait_a(v_empty + 8 * st, ph);
              mbar_arrive_expect_tx_a(v_full + 8 * st, A3_TILE_BYTES);
              tma_load_3d_a(sV + st * A3_TILE_BYTES, &tmV, v_full + 8 * st, it.h * A3_D, j * A3_BN, it.b);
            }
            ++ic;
          }
        } else if (warp == 9) {
          // ============================ MMA issuer 1: S_t = Q_t K^T + A_t B^T =======================
          constexpr uint32_t idesc_qk = umma_idesc_bf16(A3_BM, A3_BN, 0, 0);
          const uint32_t tS[2] = {tmem_base, tmem_base + 128};
          const uint64_t d_bfull = a3_desc_k16(sBfull), d_blast = a3_desc_k16(sBlast);
          const uint64_t d_a[2] = {a3_desc_k16(sA), a3_desc_k16(sA + A3_BIAS_BYTES)};
          A3_DBG_DECL
          for (int i = 0; i < my_items; ++i) {
            if (pass == 1 && ((redo_bits[i >> 5] >> (i & 31)) & 1u) == 0) continue;
            const int qs = ic % A3_Q_STAGES;
            for (int j = 0; j < n_kv; ++j, ++g) {
              const int ks = g % A3_KV_STAGES;
              if (j == 0) mbar_wait_a(q_full + 8 * qs, (ic / A3_Q_STAGES) & 1);
              mbar_wait_a(k_full + 8 * ks, (g / A3_KV_STAGES) & 1);
              A3_DBG_LAP(8)
              const uint64_t dk = umma_desc_sw128(sK + ks * A3_TILE_BYTES);
              const uint64_t db = (j == n_kv - 1) ? d_blast : d_bfull;
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                // S_t of the previous step is in the softmax threads' registers and A_t holds the offsets for this step
                if (g > 0) mbar_wait_a(s_free + 8 * t, (g - 1) & 1);
                A3_DBG_LAP(9)
                tc_fence_after();
                const uint64_t dq = umma_desc_sw128(sQ + (2 * qs + t) * A3_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < A3_D / 16; ++k) umma_ss(tS[t], dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
                umma_ss(tS[t], d_a[t], db, idesc_qk, 1u);
                A3_DBG_LAP(10)
                umma_commit_a(s_full + 8 * t);
                A3_DBG_LAP(11)
              }
              umma_commit_a(k_empty + 8 * ks);
              if (j == n_kv - 1) umma_commit_a(q_empty + 8 * qs);
              A3_DBG_LAP(11)
            }
            ++ic;
          }
          A3_DBG_FLUSH
        } else if (warp == 10 || (PV2 && warp == 11)) {
          // ============================ MMA issuer 2: O_t (+)= P_t V ============================
          // (one issuer thread per Q tile was measured: 0.69 vs 0.65 ms — the issuer is not what the step waits for)
          constexpr uint32_t idesc_pv = umma_idesc_bf16(A3_BM, A3_D, 0, 1);
          A3_DBG_DECL
          for (int i = 0; i < my_items; ++i) {
            if (pass == 1 && ((redo_bits[i >> 5] >> (i & 31)) & 1u) == 0) continue;
            for (int j = 0; j < n_kv; ++j, ++g) {
              const int vs = g % A3_KV_STAGES;
              mbar_wait_a(v_full + 8 * vs, (g / A3_KV_STAGES) & 1);
              A3_DBG_LAP(12)
              const uint64_t dv = umma_desc_sw128(sV + vs * A3_TILE_BYTES);
              for (int t = (PV2 ? warp - 10 : 0); t <= (PV2 ? warp - 10 : 1); ++t) {
                const uint32_t tP_t = tmem_base + 256 + 64 * t, tO_t = tmem_base + 384 + 64 * t;
                mbar_wait_a(p_full + 8 * t, g & 1);
                A3_DBG_LAP(13)
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < A3_BN / 16; ++kk)
                  umma_ts(tO_t, tP_t + 8 * kk, dv + kk * (2048 >> 4), idesc_pv, (j | kk) != 0 ? 1u : 0u);
                A3_DBG_LAP(14)
                umma_commit_a(pv_done + 8 * t);
                A3_DBG_LAP(15)
              }
              umma_commit_a(v_empty + 8 * vs);
              A3_DBG_LAP(15)
            }
          }
          A3_DBG_FLUSH
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ===================================== softmax warpgroups ================================
    const int t = warp >> 2;
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;
    const uint32_t tP = tmem_base + 256 + t * 64 + lane_off;
    const uint32_t tO = tmem_base + 384 + t * 64 + lane_off;
    const uint32_t b_s_full = s_full + 8 * t, b_s_free = s_free + 8 * t;
    const uint32_t b_p_full = p_full + 8 * t, b_pv_done = pv_done + 8 * t;
    const uint32_t a_row = sA + t * A3_BIAS_BYTES + row_in_tile * 16;      // this row's (-m, 1) word
    uint8_t* stg_base = smO + t * A3_TILE_BYTES;
    uint8_t* stg = stg_base + row_in_tile * 128;
    int g = 0;
    A3_DBG_DECL
    float m_baked = 0.0f;                            // what A_t holds (this thread is its only writer), i.e. what the MMA warp
                                                     // subtracts from the next score tile it issues for this row
    uint32_t s_polled = 0;                           // the poll of the score tile at hand was issued (and succeeded) last step

    // ---- item epilogue: O / l -> bf16 -> swizzled smem staging -> TMA store.  DEFERRED: it runs inside the NEXT item's first
    // step, after that step's exponentials and before its P is published (the next P_t V is what overwrites O_t), so the wait
    // for the item's last P_t V (~1,400 cycles when taken right after the last step) is hidden behind useful work. ----
    bool ep_pending = false;
    float ep_l = 1.0f;
    int ep_i = 0;
    auto epilogue = [&](uint32_t pv_parity) {
      mbar_wait_a(b_pv_done, pv_parity);
      A3_DBG_LAP(5)
      tc_fence_after();
      const A3Item it = a3_item(blockIdx.x + ep_i * gridDim.x, n_qb, p.H);
      const float inv_l = 1.0f / ep_l;
      uint32_t r0[32], r1[32];
      tmem_ld_x32(tO, r0);
      tmem_ld_x32(tO + 32, r1);
      tmem_ld_wait();
      if (q == 0 && lane == 0) tma_store_wait_read<0>();
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
            "r"(sO + t * A3_TILE_BYTES), "r"(it.h * A3_D), "r"(it.qb * 2 * A3_BM + t * A3_BM), "r"(it.b)
            : "memory");
        tma_store_commit();
      }
      ep_pending = false;
      A3_DBG_LAP(6)
    };

    for (int pass = 0; pass < 2; ++pass) {
      const bool exact = pass == 1;
      if (pass == 1) {
        if (q == 0 && lane == 0) tma_store_wait_all<0>();      // pass-0 stores of re-run items must not land after the new ones
        // (an mbarrier, not a block barrier: the two warp roles would reach it from different program locations, which
        //  compute-sanitizer's synccheck reports as a divergent barrier)
        __syncwarp();
        if (lane == 0) mbar_arrive_a(pass_done);
        mbar_wait_a(pass_done, 0);
        if (*redo_any == 0) break;
      }
      for (int i = 0; i < my_items; ++i) {
        if (exact && ((redo_bits[i >> 5] >> (i & 31)) & 1u) == 0) continue;
        float m_used = 0.0f;                         // what the accumulators refer to (set by the first tile)
        float2 la = make_float2(0.0f, 0.0f);
        float2 lb = make_float2(0.0f, 0.0f);

        for (int j = 0; j < n_kv; ++j, ++g) {
          if (!s_polled) mbar_wait_a(b_s_full, g & 1);
          A3_DBG_LAP(0)
          tc_fence_after();
          uint32_t s[4][32];
          tmem_ld_x32(tS + 0, s[0]);
          tmem_ld_x32(tS + 32, s[1]);
          tmem_ld_x32(tS + 64, s[2]);
          tmem_ld_x32(tS + 96, s[3]);
          tmem_ld_wait();
          A3_DBG_LAP(1)
          const bool last = j == n_kv - 1;
          const float m_tile = m_baked;              // the offset this tile was issued with
          bool shifted;
          bool waited_pv = false;
          if (j == 0 || exact) {
            // ---- exact row maximum of this tile (masked keys sit at -3.4e38) ----
            float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
#pragma unroll
              for (int e = 0; e < 32; e += 2) mp[ch] = fmaxf(mp[ch], fmaxf(__uint_as_float(s[ch][e]), __uint_as_float(s[ch][e + 1])));
            const float mt = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3])) + m_tile;       // true base-2 logit maximum
            if (j == 0) {
              m_used = a3_bf16_round(mt);
            } else {
              // exact mode, later tiles: lazy rescaling as in pm_attn.cu
              const float m_new = fmaxf(m_used, mt);
              const bool need = (m_new - m_used) > 8.0f;
              if (__any_sync(0xffffffffu, need)) {
                const float alpha = need ? a3_ex2(m_used - m_new) : 1.0f;
                if (need) m_used = m_new;
                la.x *= alpha; la.y *= alpha; lb.x *= alpha; lb.y *= alpha;
                mbar_wait_a(b_pv_done, (g - 1) & 1);
                tc_fence_after();
                waited_pv = true;
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
            shifted = true;
          } else {
            // ---- fast mode, tiles 1..: the scores arrived as s - m_used; nothing to decide ----
            shifted = false;
          }
          A3_DBG_LAP(24)
          // ---- publish the offset the NEXT score tile of this Q tile is issued with (fast pass only), then hand S_t back ----
          // every item's FIRST tile is issued with offset 0 (written back on the item's last tile), so that a row's result does
          // not depend on what the CTA processed before: batch-invariant bits
          if (!exact && (j == 0 || last)) {
            const float m_next = (!last && fabsf(m_used) < 3.0e38f) ? m_used : 0.0f;      // never publish inf / NaN
            if (m_next != m_baked) {
              const uint32_t w = 0x3F800000u | (static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(-m_next))));
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_row), "r"(w) : "memory");
              fence_proxy_async_smem();
              m_baked = m_next;
            }
          }
          const float delta = m_tile - m_used;
          A3_DBG_LAP(25)
          tc_fence_before();
          A3_DBG_LAP(26)
          if (j == 0 || last || exact) {
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_s_free);                            // release: orders the lanes' A_t writes
          } else if (lane == 0) {
            // nothing but completed tcgen05.ld reads to publish (ordered by the tcgen05 fence above)
            asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(b_s_free) : "memory");
          }
          A3_DBG_LAP(2)

          // (first step of an item: the previous item's last P_t V — its epilogue is still pending — reads the P columns too)
          const bool wait_pv = (j > 0 || ep_pending) && !waited_pv;
          // is there a next step in this pass?  (its s_full poll is issued under this step's exponentials)
          const bool poll_next = j + 1 < n_kv || (!exact && i + 1 < my_items);
#ifdef PM_A3_DEBUG
          unsigned long long* dbg_wait = dbg_loc;
#else
          unsigned long long* dbg_wait = nullptr;
#endif
          if (shifted) s_polled = a3_exps<true, EMU, DEFER>(s, delta, la, lb, tP, b_pv_done, wait_pv, (g - 1) & 1, poll_next, b_s_full, (g + 1) & 1, dbg_wait);
          else s_polled = a3_exps<false, EMU, DEFER>(s, 0.0f, la, lb, tP, b_pv_done, wait_pv, (g - 1) & 1, poll_next, b_s_full, (g + 1) & 1, dbg_wait);
          A3_DBG_LAP(3)
          tmem_st_wait();
          // the previous item's output leaves O_t before this item's first P_t V (issued on p_full) overwrites it
          if (j == 0 && ep_pending) epilogue((g - 1) & 1);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(b_p_full);
          A3_DBG_LAP(4)
        }

        // ---- end of item: its epilogue is deferred into the next item's first step (or runs after the loop) ----
        const float l_sum = (la.x + la.y) + (lb.x + lb.y);
        if (!exact) {
          // overflow: some logit of tiles 1.. lies more than ~2^100 above the first tile's maximum (P, l or O may have left
          // fp32's range): huge / inf / NaN row sum
          const bool bad = !(l_sum < 1.0e30f);
          if (__any_sync(0xffffffffu, bad) && lane == 0) {
            atomicOr(&redo_bits[i >> 5], 1u << (i & 31));
            *redo_any = 1;
          }
        }
        ep_pending = true;
        ep_l = l_sum;
        ep_i = i;
        if (!PREF) epilogue((g - 1) & 1);
      }
      if (ep_pending) epilogue((g - 1) & 1);
    }
#ifdef PM_A3_DEBUG
    if (lane == 0 && q == 0) {
      if (t == 1)
        for (int di = 0; di < 8; ++di) { dbg_loc[16 + di] = dbg_loc[di]; dbg_loc[di] = 0; }
      A3_DBG_FLUSH
    }
#endif
    if (q == 0 && lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, A3_TMEM_COLS);
  }
}

using Attn3KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);

struct Attn3Variant {
  int emu, defer, pref, pv2;
  Attn3KernelFn fn;
};
#define PM_ATTN3_V(E, D, P, V) {E, D, P, V, attn3_kernel<E, D, (P) != 0, (V) != 0>}
// The first entry is the default; PM_ATTN3_VARIANT="emu,defer,pref,pv2" picks another one (tuning aid, same function).
static const Attn3Variant kAttn3Variants[] = {
    PM_ATTN3_V(1, 3, 1, 0),
    PM_ATTN3_V(1, 3, 1, 1),
    PM_ATTN3_V(2, 3, 1, 0),
    PM_ATTN3_V(2, 3, 1, 1),
    PM_ATTN3_V(1, 2, 1, 1),
    PM_ATTN3_V(2, 2, 1, 1),
    PM_ATTN3_V(1, 3, 0, 0),
};
#undef PM_ATTN3_V

bool pm_attn3_supported(const AttnParams& p) {
  if (p.lse != nullptr || p.o32 != nullptr) return false;        // training outputs: pm_attn.cu
  const long long items = static_cast<long long>((p.Nq + 2 * A3_BM - 1) / (2 * A3_BM)) * p.H * p.B;
  const int sms = pm_num_sms();
  return (items + sms - 1) / sms <= A3_MAX_ITEMS;
}

int pm_attn3_launch(const AttnParams& p, cudaStream_t stream) {
  if (p.q == nullptr || p.k == nullptr || p.v == nullptr || p.o == nullptr) return PM_ERR_INVALID;
  if (p.B <= 0 || p.H <= 0 || p.Nq <= 0 || p.Nk <= 0 || p.head_dim != A3_D) return PM_ERR_INVALID;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  const uint64_t inner = static_cast<uint64_t>(p.H) * A3_D;
  if ((rc = pm_make_tmap_3d(&tmQ, p.q, 2, p.B, p.Nq, inner, p.ldq, p.bsq, A3_BM, A3_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmK, p.k, 2, p.B, p.Nk, inner, p.ldk, p.bsk, A3_BN, A3_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmV, p.v, 2, p.B, p.Nk, inner, p.ldv, p.bsv, A3_BN, A3_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmO, p.o, 2, p.B, p.Nq, inner, p.ldo, p.bso, A3_BM, A3_D)) != PM_OK) return rc;
  static Attn3KernelFn fn = nullptr;
  if (fn == nullptr) {
    const Attn3Variant* v = &kAttn3Variants[0];
    const char* env = getenv("PM_ATTN3_VARIANT");
    if (env != nullptr) {
      int e = -9, df = -9, pf = -9, pv = -9;
      if (sscanf(env, "%d,%d,%d,%d", &e, &df, &pf, &pv) != 4) return PM_ERR_INVALID;
      v = nullptr;
      for (const Attn3Variant& c : kAttn3Variants)
        if (c.emu == e && c.defer == df && c.pref == pf && c.pv2 == pv) v = &c;
      if (v == nullptr) return PM_ERR_INVALID;
    }
    fn = v->fn;
  }
  static bool attr_done[PM_MAX_DEVICES] = {};
  if ((rc = pm_ensure_dyn_smem(fn, A3_SMEM_BYTES, attr_done)) != 0) return rc;
  const long long items = static_cast<long long>((p.Nq + 2 * A3_BM - 1) / (2 * A3_BM)) * p.H * p.B;
  const int grid = items < pm_num_sms() ? static_cast<int>(items) : pm_num_sms();
  fn<<<grid, A3_THREADS, A3_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  return static_cast<int>(cudaGetLastError());
}

#ifdef PM_A3_DEBUG
extern "C" int pm_debug_a3_counters(unsigned long long* out32, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out32, g_a3_dbg, sizeof(g_a3_dbg));
  if (e == cudaSuccess && reset) {
    unsigned long long z[32] = {};
    e = cudaMemcpyToSymbol(g_a3_dbg, z, sizeof(z));
  }
  return static_cast<int>(e);
}
#endif

}  // namespace pm
