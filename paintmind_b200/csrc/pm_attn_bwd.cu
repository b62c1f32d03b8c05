// pm_attn_bwd.cu — backward of the flash attention core on tcgen05 / TMEM (sm_100a), head_dim = 64.
//
// What autograd derives in the reference for modules/attention.py:52-57 (sim = einsum(q * scale, k); softmax;
// out = einsum(attn, v)) — SURVEY.md §8f row 4.  With P = softmax(scale * Q K^T) recomputed from the forward's
// row log-sum-exp (base 2, `lse`) and delta = rowsum(dO ⊙ O):
//     dV = P^T dO          dS = P ⊙ (dO V^T - delta) * scale          dQ = dS K          dK = dS^T Q
// Two launches of one persistent kernel template, no atomics, deterministic.  Token counts need not be multiples of 128:
// TMA zero-fills / clips the ragged tiles and the tail columns are masked.
//   DKV = false : work item = (128-query tile, head, image), loops over key tiles.   rows = queries, columns = keys
//                 S' = Q K_j^T,  dP' = dO V_j^T,  dQ += dS' K_j
//   DKV = true  : work item = (128-key tile, head, image), loops over query tiles.  rows = keys, columns = queries
//                 S' = K Q_i^T (= S^T),  dP' = V dO_i^T,  dV += P' dO_i,  dK += dS' Q_i
// In both, the two score-shaped products are SS MMAs (both operands K-major, 128 x 128 x 64) into TMEM; 256 threads
// (two warpgroups: one thread per TMEM lane = row and 64-column half) turn them into P' and dS' (packed bf16, written back
// to TMEM), which then feed TS MMAs (A from TMEM, B = the very same shared-memory tile re-read as an MN-major operand —
// no transposes anywhere).
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <type_traits>

namespace pm {

constexpr int AB_T = 128;                      // tile edge (rows and columns)
constexpr int AB_D = 64;
constexpr int AB_TILE = AB_T * AB_D * 2;       // 16 KB
constexpr int AB_NST = 3;                      // column-operand ring depth
#ifndef AB_EMU
#define AB_EMU 0                               // of every 4 score pairs, this many take the FMA-pipe exp2 (pm_common.cuh)
#endif
// cycle counters of the compute warps (scripts/attn_bwd_stalls.py) cost registers in the hot loop: compile with
// -DPM_AB_TIMERS to get them; the MMA thread's counters are always on
#ifdef PM_AB_TIMERS
#define AB_TIMER(...) __VA_ARGS__
#else
#define AB_TIMER(...)
#endif
constexpr int AB_NWG = 2;                      // compute warpgroups: each owns 128 / AB_NWG columns of every tile.  (Four groups
                                               // of 32 columns measured the same in the one-CTA-per-item kernel, but 18 warps
                                               // put 5 on one SM sub-partition: 96 registers per thread, and the persistent loop spills.)
constexpr int AB_COLS = AB_T / AB_NWG;         // 64
constexpr int AB_CH = AB_COLS / 32;            // 32-column chunks per thread and step
constexpr int AB_CW = 4 * AB_NWG;              // compute warps
constexpr int AB_THREADS = 32 * (AB_CW + 2);   // compute warpgroups, TMA warp, MMA warp
constexpr int AB_SMEM = 1024 + 4 * AB_TILE + AB_NST * 2 * AB_TILE + 2 * AB_TILE + AB_NST * 1024 + 256;   // 197 KB

__device__ __forceinline__ float ab_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (ex2.approx.f16x2 — two exponentials per instruction — was tried: ptxas splits it into two MUFU.EX2.F16 operations,
//  no throughput gain on sm_100a, 3.27 ms against 2.96 ms.)

// Work split inside a CTA: the two score-shaped products of a step are issued for the whole 128-column tile (measured on
// B200, scripts/ubench/mma_rate.cu: one tcgen05.mma with M = 128 costs ~73 cycles (SS) / ~84 cycles (TS) whatever N <= 128
// is, so instruction COUNT is what the tensor pipe is bound by and N = 128 is the efficient shape), then compute
// warpgroup g (warps 4g .. 4g+3, one thread per row) turns columns [64g, 64g + 64) into P' / dS': two compute warps share
// every SM sub-partition (one warp per scheduler cannot hide its own MUFU / TMEM latencies).  The accumulating TS MMAs of
// step t are queued BEHIND the SS MMAs of step t + 1, so the exponentials of t + 1 run underneath them; P' / dS' of t + 1
// wait in registers until those TS MMAs have read the previous ones.
// TMEM: S' [0,128) dP' [128,256) P' [256,320) dS' [320,384) acc1 [384,448) acc2 [448,512).
static_assert(AB_COLS % 32 == 0, "the compute loop works in 32-column chunks");

// Persistent: one CTA per SM walks work items (row tile, head, image) — consecutive items share (head, image), so the
// column tiles stay hot in L2.  The (item, column tile) steps form one flattened stream: the next item's row tiles are
// prefetched into the other half of a two-stage buffer, its first score products are issued while the current item's
// accumulators are still being drained, and the results leave through a dedicated staging buffer.  (One CTA per item cost
// ~6000 of ~23000 cycles in launch, TMEM allocation, barrier set-up and the exposed first loads.)
struct AbItem {
  int rt, h, b;
};
__device__ __forceinline__ AbItem ab_item(int w, int n_rt, int H) {
  AbItem it;
  it.rt = w % n_rt;
  const int r = w / n_rt;
  it.h = r % H;
  it.b = r / H;
  return it;
}

template <bool DKV>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                const __grid_constant__ CUtensorMap tmC1, const __grid_constant__ CUtensorMap tmC2,
                const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_a = smem_u32(smem_raw);
  const uint32_t sR = (raw_a + 1023u) & ~1023u;            // [2 stages][R1 16 KB | R2 16 KB]
  const uint32_t sC = sR + 4 * AB_TILE;                    // [NST][C1 16 KB | C2 16 KB]
  const uint32_t sStg = sC + AB_NST * 2 * AB_TILE;         // [2][16 KB] output staging
  const uint32_t sVec = sStg + 2 * AB_TILE;                // [NST][-lse 512 B | -scale*delta 512 B]
  const uint32_t bars = sVec + AB_NST * 1024;
  const uint32_t r_full = bars, r_empty = r_full + 16, c_full = r_empty + 16, c_empty = c_full + 8 * AB_NST;
  const uint32_t s_full = c_empty + 8 * AB_NST, p_full = s_full + 8, acc_done = p_full + 8, s_free = acc_done + 8;
  const uint32_t tmem_slot_a = s_free + 8;                 // p_full / s_free count the compute warps
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot_a - raw_a));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = DKV ? p.Nk : p.Nq, n_cols = DKV ? p.Nq : p.Nk;
  const int n_rt = (n_rows + AB_T - 1) / AB_T;
  const int T = (n_cols + AB_T - 1) / AB_T;               // column tiles per item
  const int total_items = n_rt * p.H * p.B;
  const int my_items = static_cast<int>(blockIdx.x) < total_items
                           ? (total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                           : 0;
  const int total_steps = my_items * T;

  if (warp == AB_CW && lane == 0) {
    tma_prefetch_desc(&tmR1); tma_prefetch_desc(&tmR2); tma_prefetch_desc(&tmC1); tma_prefetch_desc(&tmC2);
    tma_prefetch_desc(&tmO1);
    if (DKV) tma_prefetch_desc(&tmO2);
    auto init = [](uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); };
    for (int i = 0; i < 2; ++i) { init(r_full + 8 * i, 1); init(r_empty + 8 * i, 1); }
    for (int i = 0; i < AB_NST; ++i) { init(c_full + 8 * i, 1); init(c_empty + 8 * i, 1); }
    init(s_full, 1);
    init(p_full, AB_CW);
    init(acc_done, 1);
    init(s_free, AB_CW);
    fence_mbar_init();
  }
  if (warp == AB_CW + 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tA1 = tmem_base + 384, tA2 = tmem_base + 448;

  // Service warps run CONVERGED and an elected lane issues TMA / MMA / commit (-DPM_AB_LANE0=1: the old divergent lane-0 loops,
  // in which every tcgen05.mma costs ~16 instructions of uniform-register plumbing — see pm_attn4.cu)
#ifdef PM_AB_LANE0
#define AB_SERVICE_LANES (lane == 0)
#define AB_ONE
#else
#define AB_SERVICE_LANES true
#define AB_ONE if (elect_one())
#endif
  if (warp == AB_CW) {
    // ===================================== TMA producer ======================================
    if (AB_SERVICE_LANES) {
      int g = 0;                                             // running column-tile counter (ring position)
      for (int i = 0; i < my_items; ++i) {
        const AbItem it = ab_item(blockIdx.x + i * gridDim.x, n_rt, p.H);
        const int rs = i & 1;
        mbar_wait_a(r_empty + 8 * rs, ((i >> 1) & 1) ^ 1);
        AB_ONE {
          mbar_arrive_expect_tx_a(r_full + 8 * rs, 2 * AB_TILE);
          tma_load_3d_a(sR + rs * 2 * AB_TILE, &tmR1, r_full + 8 * rs, it.h * AB_D, it.rt * AB_T, it.b);
          tma_load_3d_a(sR + rs * 2 * AB_TILE + AB_TILE, &tmR2, r_full + 8 * rs, it.h * AB_D, it.rt * AB_T, it.b);
        }
        const size_t vec_base = (static_cast<size_t>(it.b) * p.H + it.h) * p.lse_ld;
        for (int t = 0; t < T; ++t, ++g) {
          const int st = g % AB_NST;
          mbar_wait_a(c_empty + 8 * st, ((g / AB_NST) & 1) ^ 1);
          AB_ONE {
          mbar_arrive_expect_tx_a(c_full + 8 * st, 2 * AB_TILE + (DKV ? 1024 : 0));
          tma_load_3d_a(sC + st * 2 * AB_TILE, &tmC1, c_full + 8 * st, it.h * AB_D, t * AB_T, it.b);
          tma_load_3d_a(sC + st * 2 * AB_TILE + AB_TILE, &tmC2, c_full + 8 * st, it.h * AB_D, t * AB_T, it.b);
          if (DKV) {
            // the 128 per-query (-lse, -scale * delta) values of this column tile
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sVec + st * 1024), "l"(reinterpret_cast<uint64_t>(p.nlse + vec_base + t * AB_T)), "r"(512u), "r"(c_full + 8 * st)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sVec + st * 1024 + 512), "l"(reinterpret_cast<uint64_t>(p.nds + vec_base + t * AB_T)), "r"(512u), "r"(c_full + 8 * st)
                         : "memory");
          }
          }
        }
      }
    }
  } else if (warp == AB_CW + 1) {
    // ===================================== MMA issuer ========================================
    if (AB_SERVICE_LANES && total_steps > 0) {
      constexpr uint32_t idesc_ss = umma_idesc_bf16(AB_T, AB_T, 0, 0);     // 128 x 128 x 16, both K-major
      constexpr uint32_t idesc_ts = umma_idesc_bf16(AB_T, AB_D, 0, 1);     // A from TMEM, B MN-major
      const uint32_t tS = tmem_base, tDP = tmem_base + 128, tP = tmem_base + 256, tDS = tmem_base + 320;
      long long w_c = 0, w_p = 0, w_f = 0;
      const long long t_begin = clock64();
      auto issue_ss = [&](int g) {
        const int i = g / T, t = g - i * T;
        const int st = g % AB_NST, rs = i & 1;
        if (t == 0) mbar_wait_a(r_full + 8 * rs, (i >> 1) & 1);            // a new item's row tiles
        long long t0 = clock64();
        mbar_wait_a(c_full + 8 * st, (g / AB_NST) & 1);
        w_c += clock64() - t0;
        tc_fence_after();
        const uint64_t dr1 = umma_desc_sw128(sR + rs * 2 * AB_TILE), dr2 = umma_desc_sw128(sR + rs * 2 * AB_TILE + AB_TILE);
        const uint64_t dc1 = umma_desc_sw128(sC + st * 2 * AB_TILE), dc2 = umma_desc_sw128(sC + st * 2 * AB_TILE + AB_TILE);
        AB_ONE {
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k) umma_ss(tS, dr1 + 2 * k, dc1 + 2 * k, idesc_ss, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k) umma_ss(tDP, dr2 + 2 * k, dc2 + 2 * k, idesc_ss, k != 0 ? 1u : 0u);
          umma_commit_a(s_full);
        }
      };
      // acc1 (+)= dS' C1 (and acc2 (+)= P' C2): the column operands re-read as MN-major, 128 contraction rows
      auto issue_ts = [&](int g) {
        const int i = g / T, t = g - i * T;
        const int st = g % AB_NST;
        const uint64_t dc1 = umma_desc_sw128(sC + st * 2 * AB_TILE), dc2 = umma_desc_sw128(sC + st * 2 * AB_TILE + AB_TILE);
        AB_ONE {
#pragma unroll
          for (int kk = 0; kk < AB_T / 16; ++kk) {
            umma_ts(tA1, tDS + 8 * kk, dc1 + kk * (2048 >> 4), idesc_ts, (t | kk) != 0 ? 1u : 0u);
            if (DKV) umma_ts(tA2, tP + 8 * kk, dc2 + kk * (2048 >> 4), idesc_ts, (t | kk) != 0 ? 1u : 0u);
          }
          umma_commit_a(acc_done);
          umma_commit_a(c_empty + 8 * st);
          if (t == T - 1) umma_commit_a(r_empty + 8 * (i & 1));             // every MMA that read this item's row tiles is done
        }
      };
      issue_ss(0);
      for (int g = 0; g < total_steps; ++g) {
        if (g + 1 < total_steps) {
          long long t0 = clock64();
          mbar_wait_a(s_free, g & 1);                  // S'/dP'(g) sit in the compute threads' registers:
          w_f += clock64() - t0;
          issue_ss(g + 1);                             //   the next step's scores run underneath this step's exponentials
        }
        const long long t0 = clock64();
        mbar_wait_a(p_full, g & 1);                    // P'/dS'(g) written (and, at an item's first step, its predecessor drained)
        w_p += clock64() - t0;
        tc_fence_after();
        issue_ts(g);                                   // this step's accumulation runs underneath the next step's exponentials
      }
      if (p.debug != nullptr && lane == 0) {
        long long* d = p.debug + 8 * static_cast<size_t>(blockIdx.x);
        d[0] = w_f; d[1] = w_c; d[2] = w_p; d[3] = clock64() - t_begin;
      }
    }
  } else {
    // ===================================== compute warpgroups ================================
    const int gq = warp >> 2;                                 // column slice
    const int q = warp & 3;                                   // TMEM lane quarter
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    // per-thread constants of the step loop, opaque to the optimiser: ptxas otherwise re-derives them from %tid / %cluster_ctaid
    // (S2R + a chain of integer operations + R2UR) in front of the loads, waits, stores and arrivals of every step (see pm_attn4.cu)
    uint32_t tSg_o = tmem_base + gq * AB_COLS + lane_off, tPg_o = tmem_base + 256 + gq * (AB_COLS / 2) + lane_off, bar_o = s_full;
#ifndef PM_ABWD_PLAIN
    asm volatile("" : "+r"(tSg_o), "+r"(tPg_o), "+r"(bar_o));
#endif
    const uint32_t tSg = tSg_o, tDPg = tSg + 128;
    const uint32_t tPg = tPg_o, tDSg = tPg + 64;
    const uint32_t s_full = bar_o, p_full = bar_o + 8, acc_done = bar_o + 16, s_free = bar_o + 24, c_full = bar_o - 16 * AB_NST;
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2), sc2 = make_float2(p.scale, p.scale);
    AB_TIMER(long long w_s = 0, w_a = 0, c_math = 0, c_epi = 0;)
    int g = 0;
    for (int i = 0; i < my_items; ++i) {
      const AbItem it = ab_item(blockIdx.x + i * gridDim.x, n_rt, p.H);
      float2 nl_row = make_float2(0.f, 0.f), nd_row = make_float2(0.f, 0.f);
      if (!DKV && it.rt * AB_T + row_in_tile < p.Nq) {
        const size_t at = (static_cast<size_t>(it.b) * p.H + it.h) * p.lse_ld + it.rt * AB_T + row_in_tile;
        const float a = p.nlse[at], d = p.nds[at];
        nl_row = make_float2(a, a);
        nd_row = make_float2(d, d);
      }
      for (int t = 0; t < T; ++t, ++g) {
        const int st = g % AB_NST;
        const float* nl_s = reinterpret_cast<const float*>(smem_raw + (sVec + st * 1024 - raw_a)) + gq * AB_COLS;
        const float* nd_s = nl_s + 128;
        const int valid = n_cols - t * AB_T - gq * AB_COLS;      // < AB_COLS on a ragged last column tile (TMA zero-filled it)
        AB_TIMER(long long t0 = clock64();)
        mbar_wait_a(s_full, g & 1);
        AB_TIMER(w_s += clock64() - t0;)
        tc_fence_after();
        if (DKV) mbar_wait_a(c_full + 8 * st, (g / AB_NST) & 1);     // this thread reads the TMA-written vectors itself
        uint32_t pk[AB_CH][16], dk[AB_CH][16];
        uint32_t s[AB_CH][32], dp[AB_CH][32];
#pragma unroll
        for (int ch = 0; ch < AB_CH; ++ch) {
          tmem_ld_x32(tSg + ch * 32, s[ch]);
          tmem_ld_x32(tDPg + ch * 32, dp[ch]);
        }
        tmem_ld_wait();
        // both score slices of this step are in registers: hand the TMEM columns back for the next step's SS MMAs
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(s_free);
        // P' = exp2(s * scale * log2e - lse);  dS' = P' * (dP' - delta) * scale   (packed fp32 pairs around the two MUFU ops)
        auto tile_math = [&](auto masked) {
#pragma unroll
          for (int ch = 0; ch < AB_CH; ++ch) {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              float2 nl = nl_row, nd = nd_row;
              if (DKV) {
                nl = *reinterpret_cast<const float2*>(nl_s + ch * 32 + e);
                nd = *reinterpret_cast<const float2*>(nd_s + ch * 32 + e);
              }
              const float2 a = __ffma2_rn(make_float2(__uint_as_float(s[ch][e]), __uint_as_float(s[ch][e + 1])), c2, nl);
              float2 pp = (((e >> 1) & 3) < AB_EMU) ? exp2_poly2(a) : make_float2(ab_ex2(a.x), ab_ex2(a.y));
              float2 gg = __fmul2_rn(pp, __ffma2_rn(make_float2(__uint_as_float(dp[ch][e]), __uint_as_float(dp[ch][e + 1])), sc2, nd));
              if (decltype(masked)::value) {                    // ragged last tile: columns past the end contribute nothing
                if (ch * 32 + e >= valid) pp.x = 0.f, gg.x = 0.f;
                if (ch * 32 + e + 1 >= valid) pp.y = 0.f, gg.y = 0.f;
              }
              pk[ch][e >> 1] = pack_bf16x2(pp.x, pp.y);
              dk[ch][e >> 1] = pack_bf16x2(gg.x, gg.y);
            }
          }
        };
        AB_TIMER(t0 = clock64();)
        if (valid < AB_COLS) tile_math(std::true_type{});
        else tile_math(std::false_type{});
        AB_TIMER(c_math += clock64() - t0;)
        if (g > 0) {
          // the TS MMAs of the previous step (queued behind this step's SS MMAs) read P' / dS' until this fires
          AB_TIMER(t0 = clock64();)
          mbar_wait_a(acc_done, (g - 1) & 1);
          AB_TIMER(w_a += clock64() - t0;)
          tc_fence_after();
        }
#pragma unroll
        for (int ch = 0; ch < AB_CH; ++ch) {
          tmem_st_x16(tDSg + ch * 16, dk[ch]);
          if (DKV) tmem_st_x16(tPg + ch * 16, pk[ch]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(p_full);
      }
      // ---- item epilogue: accumulators -> bf16 -> swizzled staging -> TMA store ----
      // The next item's first TS MMA (which overwrites the accumulators) is only issued
      // after every compute warp's next p_full arrival, i.e. after these reads.
      AB_TIMER(long long t1 = clock64();)
      mbar_wait_a(acc_done, (g - 1) & 1);
      tc_fence_after();
      // two compute warpgroups: DKV: group 0 drains acc1 (dK), group 1 acc2 (dV); dQ: each group drains 32 of acc1's 64 columns
      static_assert(AB_NWG == 2, "the epilogue split below assumes two compute warpgroups");
      uint32_t r0[32], r1[32];
      const bool second = DKV && gq == 1;
      const int col0 = DKV ? 0 : gq * 32;
      tmem_ld_x32((second ? tA2 : tA1) + lane_off + col0, r0);
      if (DKV) tmem_ld_x32((second ? tA2 : tA1) + lane_off + 32, r1);
      tmem_ld_wait();
      if (threadIdx.x == 0) tma_store_wait_read<0>();          // the previous item's stores have left the staging tiles
      named_bar_sync(1, 32 * AB_CW);
      {
        uint8_t* stg = smem_raw + (sStg + (second ? AB_TILE : 0) - raw_a) + row_in_tile * 128;
#pragma unroll
        for (int jv = 0; jv < 4; ++jv) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r0[jv * 8 + 0]), __uint_as_float(r0[jv * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(r0[jv * 8 + 2]), __uint_as_float(r0[jv * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(r0[jv * 8 + 4]), __uint_as_float(r0[jv * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(r0[jv * 8 + 6]), __uint_as_float(r0[jv * 8 + 7]));
          *reinterpret_cast<uint4*>(stg + (((col0 / 8 + jv) ^ (row_in_tile & 7)) << 4)) = o;
        }
        if (DKV) {
#pragma unroll
          for (int jv = 0; jv < 4; ++jv) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r1[jv * 8 + 0]), __uint_as_float(r1[jv * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(r1[jv * 8 + 2]), __uint_as_float(r1[jv * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(r1[jv * 8 + 4]), __uint_as_float(r1[jv * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(r1[jv * 8 + 6]), __uint_as_float(r1[jv * 8 + 7]));
            *reinterpret_cast<uint4*>(stg + (((4 + jv) ^ (row_in_tile & 7)) << 4)) = o;
          }
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 32 * AB_CW);
      if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&tmO1)), "r"(sStg), "r"(it.h * AB_D), "r"(it.rt * AB_T), "r"(it.b) : "memory");
        if (DKV)
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmO2)), "r"(sStg + AB_TILE), "r"(it.h * AB_D), "r"(it.rt * AB_T), "r"(it.b) : "memory");
        tma_store_commit();
      }
      AB_TIMER(c_epi += clock64() - t1;)
    }
    if (threadIdx.x == 0) {
      tma_store_wait_all<0>();
      AB_TIMER(if (p.debug != nullptr) {
        long long* d = p.debug + 8 * static_cast<size_t>(blockIdx.x);
        d[4] = w_s; d[5] = w_a; d[6] = c_math; d[7] = c_epi;
      })
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AB_CW + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int pm_attn_bwd_launch(const AttnBwdParams& p, cudaStream_t stream) {
  if (p.q == nullptr || p.k == nullptr || p.v == nullptr || p.dO == nullptr || p.nlse == nullptr || p.nds == nullptr) return PM_ERR_INVALID;
  if (p.dq == nullptr || p.dk == nullptr || p.dv == nullptr) return PM_ERR_INVALID;
  if (p.B <= 0 || p.H <= 0 || p.Nq <= 0 || p.Nk <= 0 || p.head_dim != AB_D) return PM_ERR_INVALID;
  if (p.lse_ld < (p.Nq + AB_T - 1) / AB_T * AB_T || (p.lse_ld % 4) != 0) return PM_ERR_INVALID;     // whole 512-byte vector tiles are copied
  CUtensorMap tQ, tK, tV, tDO, tDQ, tDK, tDV;
  int rc;
  const uint64_t inner = static_cast<uint64_t>(p.H) * AB_D;
  if ((rc = pm_make_tmap_3d(&tQ, p.q, 2, p.B, p.Nq, inner, p.ldq, p.bsq, AB_T, AB_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tK, p.k, 2, p.B, p.Nk, inner, p.ldk, p.bsk, AB_T, AB_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tV, p.v, 2, p.B, p.Nk, inner, p.ldv, p.bsv, AB_T, AB_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tDO, p.dO, 2, p.B, p.Nq, inner, p.lddo, p.bsdo, AB_T, AB_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tDQ, p.dq, 2, p.B, p.Nq, inner, p.lddq, p.bsdq, AB_T, AB_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tDK, p.dk, 2, p.B, p.Nk, inner, p.lddk, p.bsdk, AB_T, AB_D)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_3d(&tDV, p.dv, 2, p.B, p.Nk, inner, p.lddv, p.bsdv, AB_T, AB_D)) != PM_OK) return rc;
  static bool attr_a[PM_MAX_DEVICES] = {}, attr_b[PM_MAX_DEVICES] = {};
  if ((rc = pm_ensure_dyn_smem(attn_bwd_kernel<false>, AB_SMEM, attr_a)) != 0) return rc;
  if ((rc = pm_ensure_dyn_smem(attn_bwd_kernel<true>, AB_SMEM, attr_b)) != 0) return rc;
  // dK / dV: rows = keys (R1 = K, R2 = V), columns = queries (C1 = Q, C2 = dO); acc1 = dS' Q = dK, acc2 = P' dO = dV
  const long long items_kv = static_cast<long long>((p.Nk + AB_T - 1) / AB_T) * p.H * p.B;
  const long long items_q = static_cast<long long>((p.Nq + AB_T - 1) / AB_T) * p.H * p.B;
  const int grid_kv = items_kv < pm_num_sms() ? static_cast<int>(items_kv) : pm_num_sms();
  const int grid_q = items_q < pm_num_sms() ? static_cast<int>(items_q) : pm_num_sms();
  attn_bwd_kernel<true><<<grid_kv, AB_THREADS, AB_SMEM, stream>>>(tK, tV, tQ, tDO, tDK, tDV, p);
  if ((rc = static_cast<int>(cudaGetLastError())) != 0) return rc;
  // dQ: rows = queries (R1 = Q, R2 = dO), columns = keys (C1 = K, C2 = V); acc1 = dS' K = dQ
  attn_bwd_kernel<false><<<grid_q, AB_THREADS, AB_SMEM, stream>>>(tQ, tDO, tK, tV, tDQ, tDQ, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
