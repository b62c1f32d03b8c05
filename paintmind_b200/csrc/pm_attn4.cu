// pm_attn4.cu — forward attention for PRE-SCALED queries (head_dim 64) with SIXTEEN softmax warps: the bias-MMA scheme of
// pm_attn3.cu (row maximum subtracted by the tensor core, no per-tile maximum after the first key tile, overflow -> exact re-run)
// on the four-warps-per-scheduler layout of pm_attn2.cu.
//
// Same function as pm_attn.cu (reference modules/attention.py:51-58 / :84-106), contract of pm_attn3.cu: Q carries
// scale * log2(e).
//
// Why both together.  Per-phase cycle counters in pm_attn3.cu (scripts/attn3_debug.py) showed where a 128-key step of the
// 8-warp kernel goes: ~1,500 cycles in the exponential phase — the MUFU pipe's floor for two warps per scheduler (96 MUFU.EX2
// per warp and step at 8 cycles each; skipping 3/4 of the exponentials only brought 0.65 -> 0.54 ms) — plus ~1,000 cycles of
// serial latency per warp and step (two mbarrier polls at 100-200 cycles each, the tcgen05.ld of the score row, the P store,
// fences, the epilogue), during which BOTH warps of a scheduler sit idle because they run in lock-step.  Making the two phases
// exclusive (named-barrier ping-pong) is slower: one warp alone cannot keep the MUFU busy (in-order issue, ~1,350 cycles for
// its 96 exponentials and their consumers).  Four warps per scheduler hide one warp's serial phase behind the others'
// exponentials; pm_attn2.cu had that layout but paid for it with redundant row maxima (each of the two threads of a row
// reduced all 128 scores: 6,400 vs 5,200 warp-instructions per step).  With the maximum gone from the steady state the
// redundancy is gone too: a thread loads, exponentiates and stores only its own 64 keys.
//
// Every query row is shared by two threads (same TMEM lane, warps w and w + 4), each owning 64 of the tile's 128 keys.
//   * first key tile of an item (and every tile of the exact pass): both threads reduce all 128 scores (bitwise the same
//     maximum, no exchange); the kh = 0 thread writes the row's (-m, 1) word of A_t before S_t is handed back;
//   * later tiles: tcgen05.ld of the own 64 scores (as two halves of 32, so that the step fits its 104 registers without local
//     memory), exp2, row sums, bf16 P into the own 32 P columns — nothing else; the loop's barrier / TMEM addresses are kept in
//     registers (opaque to ptxas, which otherwise re-derives them from %tid in front of every wait, load, store and arrival);
//   * P_t V is issued in two halves; row sums meet once per item in shared memory; each thread stores 32 output columns.
// 640 threads: warps 0-15 softmax, 16 TMA producer, 17 / 18 MMA issuers.  Registers 104 / 64 after setmaxnreg (640 x 96 pool).
// TMEM (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <stdio.h>
#include <stdlib.h>

namespace pm {

constexpr int A4_BM = 128;
constexpr int A4_BN = 128;
constexpr int A4_D = 64;
constexpr int A4_TILE_BYTES = 128 * 64 * 2;
constexpr int A4_KV_STAGES = 3;
constexpr int A4_Q_STAGES = 2;
constexpr int A4_THREADS = 640;
constexpr int A4_TMEM_COLS = 512;
constexpr int A4_BIAS_BYTES = 128 * 16 * 2;       // one [128 x 16] bf16 K-major no-swizzle operand (see pm_attn3.cu)
constexpr int A4_MAX_ITEMS = 4096;                // per CTA (bitmap of items to re-run in exact mode)
// smem: Q [2][2] | K [3] | V [3] | O staging [2] | A_0 A_1 B_full B_last | row sums [2][2][128] | redo bitmap | barriers
constexpr int A4_SMEM_BYTES = 1024 + (2 * A4_Q_STAGES + 2 * A4_KV_STAGES + 2) * A4_TILE_BYTES + 4 * A4_BIAS_BYTES + 2 * 2 * 128 * 4 +
                              A4_MAX_ITEMS / 8 + 16 + 512;

__device__ __forceinline__ float a4_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// K-major operand without swizzle, K = 16: row r of k-chunk c at c * 2048 + r * 16 bytes (SBO = 128 B, LBO = 2048 B)
__device__ __forceinline__ uint64_t a4_desc_k16(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(2048 >> 4) << 16;
  d |= static_cast<uint64_t>(128 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// (exp2_poly2 clamps from below only; 2^128 has the bit pattern of inf, which is what the overflow check looks for)
__device__ __forceinline__ float2 a4_exp2_poly2(float2 a) {
  a.x = fminf(a.x, 128.0f);
  a.y = fminf(a.y, 128.0f);
  return exp2_poly2(a);
}
__device__ __forceinline__ float a4_bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Every polling thread takes issue slots from the four softmax warps of its scheduler (measured: tight try_wait / test_wait loops in
// the two MMA issuer threads 0.649 -> 0.758 ms; an idle warp polling for the whole kernel +2 %).  The TMA producer runs three stages
// ahead and is latency-insensitive: it sleeps PM_A4_PROD_SLEEP ns after a failed poll (0.7603 -> 0.7565 ms sustained with 200,
// 0.7587 with 1000).  The same idea on the MMA issuers (a nanosleep between issuing a step and polling for the next) is a loss:
// 300 ns in the P.V issuer 0.786 ms, 700 ns 0.880 ms — their wake-up latency is on the critical path (profiles/r02_attention.md).
#ifndef PM_A4_LD2_AT
#define PM_A4_LD2_AT 24
#endif
#ifndef PM_A4_PROD_SLEEP
#define PM_A4_PROD_SLEEP 200
#endif
__device__ __forceinline__ void a4_issuer_wait(uint32_t bar, uint32_t parity) { mbar_wait_a(bar, parity); }
// The service warps run their role loops CONVERGED (all 32 lanes wait and compute descriptors in the uniform datapath) and one
// elected lane issues the TMA / MMA / commit instructions, instead of lane 0 running the whole loop in divergent code (where
// every tcgen05.mma costs ~16 instructions of R2UR / PLOP3 / ELECT plumbing on the critical path): 0.7623 -> 0.7431 ms
// sustained (-DPM_A4_LANE0=1 restores the old form).
#ifndef PM_A4_LANE0
#define A4_SERVICE_LANES true
#define A4_ONE if (elect_one())
#else
#define A4_SERVICE_LANES (lane == 0)
#define A4_ONE
#endif
__device__ __forceinline__ void a4_producer_wait(uint32_t bar, uint32_t parity) {
#if PM_A4_PROD_SLEEP > 0
  if (mbar_try_wait_a(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_a(bar, parity)) {
    __nanosleep(PM_A4_PROD_SLEEP);
    if (clock64() - t0 > PM_MBAR_TIMEOUT_CYCLES) __trap();
  }
#else
  mbar_wait_a(bar, parity);
#endif
}

struct A4Item {
  int qb, h, b;
};
__device__ __forceinline__ A4Item a4_item(int w, int n_qb, int H) {
  A4Item it;
  it.qb = w % n_qb;
  const int r = w / n_qb;
  it.h = r % H;
  it.b = r / H;
  return it;
}

// exponentials of this thread's 64 keys: P = exp2(s [+ delta]) -> packed bf16 (pk), row sums into la / lb
template <bool SHIFT, int EMU>
__device__ __forceinline__ void a4_exps(const uint32_t (&s)[2][32], float delta, float2& la, float2& lb, uint32_t (&pk)[2][16]) {
  const float2 dd2 = make_float2(delta, delta);
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      float2 a = make_float2(__uint_as_float(s[ch][e]), __uint_as_float(s[ch][e + 1]));
      if (SHIFT) a = __fadd2_rn(a, dd2);
      const bool poly = ((e >> 1) & 3) < EMU;
#if defined(PM_A4_EXPERIMENT) && PM_A4_EXPERIMENT >= 1
      // energy experiments (wrong results): 1 = no exponentials, 2 = no exponentials and no row sums, 3 = also no bf16 packing
      const float2 ex = a;
#else
      const float2 ex = poly ? a4_exp2_poly2(a) : make_float2(a4_ex2(a.x), a4_ex2(a.y));
#endif
#if !defined(PM_A4_EXPERIMENT) || PM_A4_EXPERIMENT < 2
      if ((e >> 1) & 1) lb = __fadd2_rn(lb, ex);
      else la = __fadd2_rn(la, ex);
#endif
#if defined(PM_A4_EXPERIMENT) && PM_A4_EXPERIMENT >= 3
      pk[ch][e >> 1] = __float_as_uint(ex.x) ^ __float_as_uint(ex.y);
#else
      pk[ch][e >> 1] = pack_bf16x2(ex.x, ex.y);
#endif
    }
  }
}

// the same for 32 keys (the split fast path below), elements [E0, E1)
template <int EMU, int E0 = 0, int E1 = 32>
__device__ __forceinline__ void a4_exps32(const uint32_t (&s)[32], float2& la, float2& lb, uint32_t (&pk)[16]) {
#pragma unroll
  for (int e = E0; e < E1; e += 2) {
    const float2 a = make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1]));
    const bool poly = ((e >> 1) & 3) < EMU;
    const float2 ex = poly ? a4_exp2_poly2(a) : make_float2(a4_ex2(a.x), a4_ex2(a.y));
    if ((e >> 1) & 1) lb = __fadd2_rn(lb, ex);
    else la = __fadd2_rn(la, ex);
    pk[e >> 1] = pack_bf16x2(ex.x, ex.y);
  }
}

// TRAIN : also emit the row log-sum-exp (base 2, of the pre-scaled logits) and an fp32 copy of the output (training forward)
// SPLIT : 1 = the steady-state step loads and exponentiates its 64 scores as two halves of 32 (see the fast path)
template <int EMU, bool TRAIN, int SPLIT>
__global__ void __launch_bounds__(A4_THREADS, 1)
attn4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_a = smem_u32(smem_raw);
  const uint32_t sQ = (raw_a + 1023u) & ~1023u;
  const uint32_t sK = sQ + 2 * A4_Q_STAGES * A4_TILE_BYTES;
  const uint32_t sV = sK + A4_KV_STAGES * A4_TILE_BYTES;
  const uint32_t sO = sV + A4_KV_STAGES * A4_TILE_BYTES;
  const uint32_t sA = sO + 2 * A4_TILE_BYTES;                     // [2] per-tile (-m, 1, 0...) operands
  const uint32_t sBfull = sA + 2 * A4_BIAS_BYTES;                 // (1, 0, 0...) for every key
  const uint32_t sBlast = sBfull + A4_BIAS_BYTES;                 // (1, key >= valid ? -max : 0, 0...): ragged last tile
  const uint32_t sL = sBlast + A4_BIAS_BYTES;                     // [2 tiles][2 halves][128] fp32 partial row sums
  const uint32_t sRedo = sL + 2 * 2 * 128 * 4;                    // bitmap [A4_MAX_ITEMS] + flag word behind it
  const uint32_t bars = sRedo + A4_MAX_ITEMS / 8 + 16;
  const uint32_t q_full = bars;
  const uint32_t q_empty = q_full + 8 * A4_Q_STAGES;
  const uint32_t k_full = q_empty + 8 * A4_Q_STAGES;
  const uint32_t k_empty = k_full + 8 * A4_KV_STAGES;
  const uint32_t v_full = k_empty + 8 * A4_KV_STAGES;
  const uint32_t v_empty = v_full + 8 * A4_KV_STAGES;
  const uint32_t s_full = v_empty + 8 * A4_KV_STAGES;             // [2]     MMA -> softmax: S_t complete
  const uint32_t s_free = s_full + 16;                            // [2]     softmax -> MMA: S_t has been read (8 warps)
  const uint32_t p_full = s_free + 16;                            // [2][2]  softmax -> MMA: P_t columns of key half kh written
  const uint32_t pv_done = p_full + 32;                           // [2]     MMA -> softmax: O_t += P_t V complete
  const uint32_t tmem_slot_a = pv_done + 16;
  const uint32_t pass_done = tmem_slot_a + 16;                    //         softmax warps -> everyone: pass 0 complete, redo bitmap final
  uint8_t* const smO = smem_raw + (sO - raw_a);
  float* const smL = reinterpret_cast<float*>(smem_raw + (sL - raw_a));
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot_a - raw_a));
  uint32_t* const redo_bits = reinterpret_cast<uint32_t*>(smem_raw + (sRedo - raw_a));
  volatile uint32_t* const redo_any = redo_bits + A4_MAX_ITEMS / 32;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = (p.Nk + A4_BN - 1) / A4_BN;
  const int n_qb = (p.Nq + 2 * A4_BM - 1) / (2 * A4_BM);
  const int total_items = n_qb * p.H * p.B;
  const int my_items = static_cast<int>(blockIdx.x) < total_items
                           ? (total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                           : 0;
  const int valid_last = p.Nk - (n_kv - 1) * A4_BN;      // keys of the last tile, 1..128

  // ---- constant / initial bias operands, redo bitmap ----
  for (int i = threadIdx.x; i < A4_MAX_ITEMS / 32 + 1; i += A4_THREADS) redo_bits[i] = 0;
  if (threadIdx.x < 256) {
    // A_t row r: chunk 0 = (-m = 0, 1, 0, ...), chunk 1 = 0
    const uint32_t a = sA + (threadIdx.x >> 7) * A4_BIAS_BYTES + (threadIdx.x & 127) * 16;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(a), "r"(0x3F800000u), "r"(0u) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a + 2048), "r"(0u) : "memory");
  } else if (threadIdx.x < 384) {
    const int r = threadIdx.x - 256;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(sBfull + r * 16), "r"(0x00003F80u), "r"(0u) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sBfull + 2048 + r * 16), "r"(0u) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(sBlast + r * 16), "r"(r < valid_last ? 0x00003F80u : 0xFF7F3F80u), "r"(0u)
                 : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sBlast + 2048 + r * 16), "r"(0u) : "memory");
  }
  fence_proxy_async_smem();

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    auto init = [](uint32_t bar, uint32_t count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    };
    for (int i = 0; i < A4_Q_STAGES; ++i) {
      init(q_full + 8 * i, 1);
      init(q_empty + 8 * i, 1);
    }
    for (int i = 0; i < A4_KV_STAGES; ++i) {
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
    init(pass_done, 16);                   // one arrival per softmax warp
    fence_mbar_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_slot, A4_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    int ic = 0, g = 0;                               // running item / step counters (barrier phases) across both passes
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) {
        // pass 0 complete in every softmax warp, bitmap final.  One lane per role warp polls (a warp without a role would
        // otherwise poll from the first cycle of the kernel and take issue slots from the softmax warps of its scheduler)
        if (A4_SERVICE_LANES && warp <= 18) mbar_wait_a(pass_done, 0);
        __syncwarp();
        if (*redo_any == 0) break;
      }
      if (A4_SERVICE_LANES) {
        if (warp == 16) {
          // ===================================== TMA producer ======================================
          for (int i = 0; i < my_items; ++i) {
            if (pass == 1 && ((redo_bits[i >> 5] >> (i & 31)) & 1u) == 0) continue;
            const A4Item it = a4_item(blockIdx.x + i * gridDim.x, n_qb, p.H);
            const int qs = ic % A4_Q_STAGES;
            a4_producer_wait(q_empty + 8 * qs, ((ic / A4_Q_STAGES) & 1) ^ 1);
            A4_ONE {
              mbar_arrive_expect_tx_a(q_full + 8 * qs, 2 * A4_TILE_BYTES);
              tma_load_3d_a(sQ + (2 * qs) * A4_TILE_BYTES, &tmQ, q_full + 8 * qs, it.h * A4_D, it.qb * 2 * A4_BM, it.b);
              tma_load_3d_a(sQ + (2 * qs + 1) * A4_TILE_BYTES, &tmQ, q_full + 8 * qs, it.h * A4_D, it.qb * 2 * A4_BM + A4_BM, it.b);
            }
            for (int j = 0; j < n_kv; ++j, ++g) {
              const int st = g % A4_KV_STAGES;
              const uint32_t ph = ((g / A4_KV_STAGES) & 1) ^ 1;
              a4_producer_wait(k_empty + 8 * st, ph);
              A4_ONE {
                mbar_arrive_expect_tx_a(k_full + 8 * st, A4_TILE_BYTES);
                tma_load_3d_a(sK + st * A4_TILE_BYTES, &tmK, k_full + 8 * st, it.h * A4_D, j * A4_BN, it.b);
              }
              a4_producer_wait(v_empty + 8 * st, ph);
              A4_ONE {
                mbar_arrive_expect_tx_a(v_full + 8 * st, A4_TILE_BYTES);
                tma_load_3d_a(sV + st * A4_TILE_BYTES, &tmV, v_full + 8 * st, it.h * A4_D, j * A4_BN, it.b);
              }
            }
            ++ic;
          }
        } else if (warp == 17) {
          // ============================ MMA issuer 1: S_t = Q_t K^T + A_t B^T =======================
          constexpr uint32_t idesc_qk = umma_idesc_bf16(A4_BM, A4_BN, 0, 0);
          const uint32_t tS[2] = {tmem_base, tmem_base + 128};
          const uint64_t d_bfull = a4_desc_k16(sBfull), d_blast = a4_desc_k16(sBlast);
          const uint64_t d_a[2] = {a4_desc_k16(sA), a4_desc_k16(sA + A4_BIAS_BYTES)};
          for (int i = 0; i < my_items; ++i) {
            if (pass == 1 && ((redo_bits[i >> 5] >> (i & 31)) & 1u) == 0) continue;
            const int qs = ic % A4_Q_STAGES;
            for (int j = 0; j < n_kv; ++j, ++g) {
              const int ks = g % A4_KV_STAGES;
              if (j == 0) a4_issuer_wait(q_full + 8 * qs, (ic / A4_Q_STAGES) & 1);
              a4_issuer_wait(k_full + 8 * ks, (g / A4_KV_STAGES) & 1);
              const uint64_t dk = umma_desc_sw128(sK + ks * A4_TILE_BYTES);
              const uint64_t db = (j == n_kv - 1) ? d_blast : d_bfull;
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                // S_t of the previous step has been read by all 8 warps and A_t holds the offsets for this step
                if (g > 0) a4_issuer_wait(s_free + 8 * t, (g - 1) & 1);
                tc_fence_after();
                const uint64_t dq = umma_desc_sw128(sQ + (2 * qs + t) * A4_TILE_BYTES);
                A4_ONE {
#pragma unroll
                  for (int k = 0; k < A4_D / 16; ++k) umma_ss(tS[t], dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
                  umma_ss(tS[t], d_a[t], db, idesc_qk, 1u);
                  umma_commit_a(s_full + 8 * t);
                  if (t == 1) {
                    umma_commit_a(k_empty + 8 * ks);
                    if (j == n_kv - 1) umma_commit_a(q_empty + 8 * qs);
                  }
                }
              }

            }
            ++ic;
          }
        } else if (warp == 18) {
          // ===================================== MMA issuer 2: O_t (+)= P_t V ========================
          constexpr uint32_t idesc_pv = umma_idesc_bf16(A4_BM, A4_D, 0, 1);
          const uint32_t tP[2] = {tmem_base + 256, tmem_base + 320};
          const uint32_t tO[2] = {tmem_base + 384, tmem_base + 448};
          for (int i = 0; i < my_items; ++i) {
            if (pass == 1 && ((redo_bits[i >> 5] >> (i & 31)) & 1u) == 0) continue;
            for (int j = 0; j < n_kv; ++j, ++g) {
              const int vs = g % A4_KV_STAGES;
              a4_issuer_wait(v_full + 8 * vs, (g / A4_KV_STAGES) & 1);
              const uint64_t dv = umma_desc_sw128(sV + vs * A4_TILE_BYTES);
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                // keys 0-63: the kh = 0 warps (which also did any rescaling of O_t) are done.  On the first step of an item the
                // first MMA OVERWRITES O_t, whose previous contents the kh = 1 warps may still be reading out: wait for both.
                a4_issuer_wait(p_full + 16 * t, g & 1);
                if (j == 0) a4_issuer_wait(p_full + 16 * t + 8, g & 1);
                tc_fence_after();
                A4_ONE {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    umma_ts(tO[t], tP[t] + 8 * kk, dv + kk * (2048 >> 4), idesc_pv, (j | kk) != 0 ? 1u : 0u);
                }
                if (j != 0) {
                  a4_issuer_wait(p_full + 16 * t + 8, g & 1);
                  tc_fence_after();
                }
                A4_ONE {
#pragma unroll
                  for (int kk = 4; kk < 8; ++kk) umma_ts(tO[t], tP[t] + 8 * kk, dv + kk * (2048 >> 4), idesc_pv, 1u);
                  umma_commit_a(pv_done + 8 * t);
                  if (t == 1) umma_commit_a(v_empty + 8 * vs);
                }
              }
            }
          }
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================================== softmax warps =====================================
    const int t = warp >> 3;                       // Q tile 0 / 1
    const int kh = (warp >> 2) & 1;                // key half of every 128-key tile owned by this thread
    const int q = warp & 3;                        // TMEM lane quarter
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    // (opaque, like the barrier bases below: ptxas otherwise re-derives these per-thread constants from %tid — S2R, five dependent
    //  integer operations, R2UR — in front of the score load and of the P store of every step)
    uint32_t tS_own = tmem_base + t * 128 + kh * 64 + lane_off;
    uint32_t tP = tmem_base + 256 + t * 64 + kh * 32 + lane_off;
    asm volatile("" : "+r"(tS_own), "+r"(tP));
    const uint32_t tS_oth = tmem_base + t * 128 + (kh ^ 1) * 64 + lane_off;
    const uint32_t tO = tmem_base + 384 + t * 64 + lane_off;
    // Two opaque base registers for the four barriers of the hot loop: left to itself ptxas re-derives each address from %tid and
    // %cluster_ctaid (two S2R, ~25 cycles each) in front of the waits and arrivals of every step.
    uint32_t bar_t = s_full + 8 * t, bar_p = p_full + 16 * t + 8 * kh;
    asm volatile("" : "+r"(bar_t), "+r"(bar_p));
    const uint32_t b_s_full = bar_t, b_s_free = bar_t + (s_free - s_full);
    const uint32_t b_p_full = bar_p, b_pv_done = bar_t + (pv_done - s_full);
    const uint32_t a_row = sA + t * A4_BIAS_BYTES + row_in_tile * 16;      // this row's (-m, 1) word (written by the kh = 0 thread)
    uint8_t* const stg = smO + t * A4_TILE_BYTES + row_in_tile * 128;
    float* const l_mine = smL + (t * 2 + kh) * 128 + row_in_tile;
    float* const l_other = smL + (t * 2 + (kh ^ 1)) * 128 + row_in_tile;
    const bool tile_leader = (kh == 0 && q == 0 && lane == 0);
    int g = 0;                                     // flattened step counter (barrier phases)
    float m_baked = 0.0f;                          // what A_t holds for this row (both threads of the row track the same value)

    for (int pass = 0; pass < 2; ++pass) {
      const bool exact = pass == 1;
      if (pass == 1) {
        if (tile_leader) tma_store_wait_all<0>();    // pass-0 stores of re-run items must not land after the new ones
        // (an mbarrier, not a block barrier: the two warp roles would reach it from different program locations, which
        //  compute-sanitizer's synccheck reports as a divergent barrier)
        __syncwarp();
        if (lane == 0) mbar_arrive_a(pass_done);
        mbar_wait_a(pass_done, 0);
        if (*redo_any == 0) break;
      }
      for (int i = 0; i < my_items; ++i) {
        if (exact && ((redo_bits[i >> 5] >> (i & 31)) & 1u) == 0) continue;
        const A4Item it = a4_item(blockIdx.x + i * gridDim.x, n_qb, p.H);
        float m_used = 0.0f;                         // what the accumulators refer to (set by the first tile)
        float2 la = make_float2(0.0f, 0.0f);         // partial sum of exp2(s - m_used) over this thread's keys
        float2 lb = make_float2(0.0f, 0.0f);

        for (int j = 0; j < n_kv; ++j, ++g) {
          uint32_t ok_pv = 1;
          mbar_wait_a(b_s_full, g & 1);
          tc_fence_after();
          uint32_t s[2][32];
          uint32_t pk[2][16];
          if (j == 0 || exact) {
            const float m_tile = m_baked;            // the offset this tile was issued with
            // ---- the partner's 64 scores: only their maximum is needed (masked keys sit at -3.4e38) ----
            float mo;
            {
              uint32_t o0[32], o1[32];
              tmem_ld_x32(tS_oth, o0);
              tmem_ld_x32(tS_oth + 32, o1);
              tmem_ld_wait();
              float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                m0 = fmaxf(m0, fmaxf(__uint_as_float(o0[e]), __uint_as_float(o0[e + 1])));
                m1 = fmaxf(m1, fmaxf(__uint_as_float(o1[e]), __uint_as_float(o1[e + 1])));
              }
              mo = fmaxf(m0, m1);
            }
            // (address dependency on `mo`, 0 unless NaN: keeps the partner's 64 registers dead before the own 64 arrive)
            const uint32_t dep = (mo != mo) ? 1u : 0u;
            tmem_ld_x32(tS_own + dep, s[0]);
            tmem_ld_x32(tS_own + 32 + dep, s[1]);
            tmem_ld_wait();
            float mp0 = -INFINITY, mp1 = -INFINITY;
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              mp0 = fmaxf(mp0, fmaxf(__uint_as_float(s[0][e]), __uint_as_float(s[0][e + 1])));
              mp1 = fmaxf(mp1, fmaxf(__uint_as_float(s[1][e]), __uint_as_float(s[1][e + 1])));
            }
            // max is exact: both threads of the row hold bitwise the same value whatever the order
            const float mt = fmaxf(mo, fmaxf(mp0, mp1)) + m_tile;          // true base-2 logit maximum of the tile
            if (j == 0) {
              m_used = a4_bf16_round(mt);
              if (!exact) {
                // offset for the item's remaining tiles; a single-tile item leaves A_t at 0 (every item's FIRST tile is issued
                // with offset 0, so that a row's result does not depend on what the CTA processed before: batch-invariant bits)
                const float m_next = (n_kv > 1 && fabsf(m_used) < 3.0e38f) ? m_used : 0.0f;      // never publish inf / NaN
                if (m_next != m_baked) {
                  if (kh == 0) {
                    const uint32_t w = 0x3F800000u | (static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(-m_next))));
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_row), "r"(w) : "memory");
                    fence_proxy_async_smem();
                  }
                  m_baked = m_next;
                }
              }
            } else {
              // exact pass, later tiles: lazy rescaling as in pm_attn.cu; the partner warp takes the same decision
              const float m_new = fmaxf(m_used, mt);
              const bool need = (m_new - m_used) > 8.0f;
              if (__any_sync(0xffffffffu, need)) {
                const float alpha = need ? a4_ex2(m_used - m_new) : 1.0f;
                if (need) m_used = m_new;
                la.x *= alpha; la.y *= alpha; lb.x *= alpha; lb.y *= alpha;
                if (kh == 0) {
                  // the kh = 0 warp rescales all 64 columns of O_t (the P_t V issuer waits for kh = 0 before its first half)
                  mbar_wait_a(b_pv_done, (g - 1) & 1);
                  tc_fence_after();
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
            }
            // S_t has been read and A_t holds the next tile's offset: hand it back (release: orders the A_t writes)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_s_free);
            if (j > 0) ok_pv = mbar_try_wait_a(b_pv_done, (g - 1) & 1);
            a4_exps<true, EMU>(s, m_tile - m_used, la, lb, pk);
          } else {
            // ---- fast pass, tiles 1..: the scores arrive as s - m_used; load, exponentiate, store ----
            if (SPLIT) {
              // Two halves of 32 keys: at most 32 scores + 16 packed words are live instead of 64 + 32.  With all 64 scores in
              // registers the loop does not fit 104 registers: ptxas kept the step counter in LOCAL memory (four LDL and one STL per
              // step, ncu: 4.5 M local loads per launch) and rebuilt every barrier address from %tid / %cluster_ctaid in front of
              // each wait and arrive — on the serial part of the step.  The second load's address depends on the first half's sums
              // (0 unless NaN), otherwise ptxas hoists it above them and the 64 registers are back.
              tmem_ld_x32(tS_own, s[0]);
              tmem_ld_wait();
              // the second load is issued when PM_A4_LD2_AT of the 32 first-half scores are done: its latency hides under the rest
              // (24: 0.611 -> 0.606 ms against issuing it after all 32)
              a4_exps32<EMU, 0, PM_A4_LD2_AT>(s[0], la, lb, pk[0]);
              const float chk = (la.x + la.y) + (lb.x + lb.y);
              const uint32_t dep2 = (chk != chk) ? 1u : 0u;
              tmem_ld_x32(tS_own + 32 + dep2, s[1]);
              a4_exps32<EMU, PM_A4_LD2_AT, 32>(s[0], la, lb, pk[0]);
              tmem_ld_wait();
              tc_fence_before();
            } else {
              tmem_ld_x32(tS_own, s[0]);
              tmem_ld_x32(tS_own + 32, s[1]);
              tmem_ld_wait();
              tc_fence_before();
            }
            if (j == n_kv - 1) {
              // last tile of the item: the next item's first tile must be issued with offset 0 (see above)
              if (m_baked != 0.0f) {
                if (kh == 0) {
                  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_row), "r"(0x3F800000u) : "memory");
                  fence_proxy_async_smem();
                }
                m_baked = 0.0f;
              }
              __syncwarp();
              if (lane == 0) mbar_arrive_a(b_s_free);                          // release: orders the A_t writes
            } else if (lane == 0) {
              // nothing but completed tcgen05.ld reads to publish (ordered by the tcgen05 fence)
              asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(b_s_free) : "memory");
            }
            ok_pv = mbar_try_wait_a(b_pv_done, (g - 1) & 1);          // poll issued now, consumed after the exponentials
            if (SPLIT) a4_exps32<EMU>(s[1], la, lb, pk[1]);
            else a4_exps<false, EMU>(s, 0.0f, la, lb, pk);
          }
          if (j > 0 && !ok_pv) mbar_wait_a(b_pv_done, (g - 1) & 1);   // P_t V of the previous step reads the P columns until this fires
          tc_fence_after();
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
        if (!exact) {
          // overflow: some logit of tiles 1.. lies more than ~2^100 above the first tile's maximum: huge / inf / NaN row sum
          const bool bad = !(l_sum < 1.0e30f);
          if (__any_sync(0xffffffffu, bad) && lane == 0) {
            atomicOr(&redo_bits[i >> 5], 1u << (i & 31));
            *redo_any = 1;
          }
        }
        const float inv_l = 1.0f / l_sum;
        mbar_wait_a(b_pv_done, (g - 1) & 1);
        tc_fence_after();
        const int qrow = it.qb * 2 * A4_BM + t * A4_BM + row_in_tile;
        if (TRAIN && p.lse != nullptr && kh == 0) {
          // training forward: base-2 log-sum-exp of the score row, consumed by pm_attn_bwd (a flagged item's value is
          // overwritten by the exact pass: same thread, program order)
          if (qrow < p.Nq) p.lse[(static_cast<size_t>(it.b) * p.H + it.h) * p.lse_ld + qrow] = m_used + log2f(l_sum);
        }
        uint32_t r0[32];
        tmem_ld_x32(tO + kh * 32, r0);
        tmem_ld_wait();
        // (O_t is free again: the next item's first P V is only issued after BOTH halves' next p_full arrivals)
        if (TRAIN && p.o32 != nullptr) {
          // training forward: an fp32 copy of the output rows (keeps delta = rowsum(dO * O) free of O's bf16 rounding)
          if (qrow < p.Nq) {
            float4* dst = reinterpret_cast<float4*>(p.o32 + (static_cast<size_t>(it.b) * p.Nq + qrow) * p.ldo32 + it.h * A4_D + kh * 32);
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
              "r"(sO + t * A4_TILE_BYTES), "r"(it.h * A4_D), "r"(it.qb * 2 * A4_BM + t * A4_BM), "r"(it.b)
              : "memory");
          tma_store_commit();
        }
      }
    }
    if (tile_leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, A4_TMEM_COLS);
  }
}

using Attn4KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);

struct Attn4Variant {
  int emu, split;
  Attn4KernelFn fn;
};
// The first entry is the default; PM_ATTN4_VARIANT="emu" / PM_ATTN4_SPLIT=0|1 pick another one (tuning aid, same function).
static const Attn4Variant kAttn4Variants[] = {
    {1, 1, attn4_kernel<1, false, 1>},
    {1, 0, attn4_kernel<1, false, 0>},
    {0, 0, attn4_kernel<0, false, 0>},
    {2, 0, attn4_kernel<2, false, 0>},
    {0, 1, attn4_kernel<0, false, 1>},
    {2, 1, attn4_kernel<2, false, 1>},
};

bool pm_attn4_supported(const AttnParams& p) {
  const long long items = static_cast<long long>((p.Nq + 2 * A4_BM - 1) / (2 * A4_BM)) * p.H * p.B;
  const int sms = pm_num_sms();
  return (items + sms - 1) / sms <= A4_MAX_ITEMS;
}

int pm_attn4_launch(const AttnParams& p, cudaStream_t stream) {
  if (p.q == nullptr || p.k == nullptr || p.v == nullptr || p.o == nullptr) return PM_ERR_INVALID;
  if (p.B <= 0 || p.H <= 0 || p.Nq <= 0 || p.Nk <= 0 || p.head_dim != A4_D) return PM_ERR_INVALID;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  const uint64_t inner = static_cast<uint64_t>(p.H) * A4_D;
  if ((rc = pm_make_tmap_3d(&tmQ, p.q, 2, p.B, p.Nq, inner, p.ldq, p.bsq, A4_BM, A4_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmK, p.k, 2, p.B, p.Nk, inner, p.ldk, p.bsk, A4_BN, A4_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmV, p.v, 2, p.B, p.Nk, inner, p.ldv, p.bsv, A4_BN, A4_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tmO, p.o, 2, p.B, p.Nq, inner, p.ldo, p.bso, A4_BM, A4_D)) != PM_OK) return rc;
  static Attn4KernelFn fn = nullptr;
  if (fn == nullptr) {
    const Attn4Variant* v = &kAttn4Variants[0];
    const char* env = getenv("PM_ATTN4_VARIANT");
    const char* env_s = getenv("PM_ATTN4_SPLIT");
    if (env != nullptr || env_s != nullptr) {
      const int e = env != nullptr ? atoi(env) : v->emu;
      const int sp = env_s != nullptr ? atoi(env_s) : v->split;
      v = nullptr;
      for (const Attn4Variant& c : kAttn4Variants)
        if (c.emu == e && c.split == sp) v = &c;
      if (v == nullptr) return PM_ERR_INVALID;
    }
    fn = v->fn;
  }
  static bool attr_done[PM_MAX_DEVICES] = {}, attr_done_train[PM_MAX_DEVICES] = {};
  Attn4KernelFn kern = fn;
  if (p.lse != nullptr || p.o32 != nullptr) {
    kern = attn4_kernel<1, true, 1>;   // (a separate instantiation: the inference kernel's register allocation is untouched)
    if ((rc = pm_ensure_dyn_smem(kern, A4_SMEM_BYTES, attr_done_train)) != 0) return rc;
  } else if ((rc = pm_ensure_dyn_smem(fn, A4_SMEM_BYTES, attr_done)) != 0) {
    return rc;
  }
  const long long items = static_cast<long long>((p.Nq + 2 * A4_BM - 1) / (2 * A4_BM)) * p.H * p.B;
  const int grid = items < pm_num_sms() ? static_cast<int>(items) : pm_num_sms();
  kern<<<grid, A4_THREADS, A4_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
