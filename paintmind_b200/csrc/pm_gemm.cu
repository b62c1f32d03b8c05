// pm_gemm.cu — persistent, warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   D[M, N] = A[M, K] (bf16, row-major)  x  W[N, K]^T (bf16, row-major, i.e. nn.Linear layout)
//
// One kernel serves every dense projection of the tokenizer hot path (SURVEY.md §8a rows
// a5, a8, a9, a11, a15, a16): the epilogue is selected by GemmParams flags:
//   * LayerNorm folded into the projection:  LN(x) W^T = rstd * (x W'^T - mu * colsum(W')) + b'
//     with W' = gamma ⊙ W, b' = b + W beta  (reference: stage1/layers.py:49-58 norm1/norm2
//     feeding modules/attention.py:34-36 and modules/mlp.py:28)
//   * + bias, + position embedding (stage1/layers.py:108,146), + residual (layers.py:55-56)
//   * SwiGLU  silu(x1) * x2 on the two halves of w12 (modules/mlp.py:27-31); the weight rows
//     are repacked so that one 256-wide tile holds 128 gate rows followed by their 128 value rows
//   * fp32 row-major store (prev_quant, vqmodel.py:23) or un-patchify + clamp to NCHW fp32
//     (stage1/layers.py:150 + vqmodel.py:30)
//
// Structure per CTA (320 threads, 1 CTA / SM, grid = min(#tiles, #SMs); optionally CTA pairs, see GemmCfg):
//   warp 0 lane 0 : TMA producer  — A and W tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1 lane 0 : MMA issuer    — tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM,
//                                   two accumulator stages so the epilogue overlaps the next tile
//   warps 2..9    : epilogue      — two groups of four warps on alternating 64-column chunks:
//                                   tcgen05.ld -> registers -> fused math -> swizzled smem staging ->
//                                   TMA store (bf16 outputs); residual tiles are TMA-prefetched one
//                                   chunk ahead into the staging buffer the result is stored from.
#include "pm_common.cuh"
#include "pm_kernels.h"

namespace pm {

constexpr int BM = 128;                       // accumulator rows per CTA (UMMA M = 128, or 256 across a CTA pair)
constexpr int BK = 64;
constexpr int STG_BYTES = BM * 128;           // one epilogue staging buffer: 128 rows x 128 B = 16 KB
constexpr int GEMM_THREADS = 320;            // TMA warp, MMA warp, 8 epilogue warps

// CTA2 = true: two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 of 256 x BN x 16.  Each CTA stages its own
// 128 rows of A and HALF of the W tile (BN/2 rows), which halves the W bytes every SM pulls through L2 and smem
// (the 1-CTA kernel is bound by L2->SM bandwidth: 128x256 tiles need ~12 TB/s at 1 PFLOP/s).
template <int BN, bool CTA2, bool SWIGLU = false>
struct GemmCfg {
  // epilogue staging buffers per epilogue warp group.  The SwiGLU projection never carries a residual (which is
  // prefetched into the staging buffer one chunk ahead and therefore needs two): one buffer per group there, and the
  // 32 KB saved buy a sixth operand stage (the MMA issuer of w12 waits 13 % of its time for operands with five).
  static constexpr int NSTG_G = SWIGLU ? 1 : 2;
  static constexpr int NSTG = 2 * NSTG_G;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // 227 KB total: staging + column vectors + barriers leave this much for the ring
  static constexpr int RING_BUDGET = 232448 - NSTG * STG_BYTES - 2 * BN * 4 - 256;
  static constexpr int STAGES_RAW = RING_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NSTG * STG_BYTES + 2 * BN * 4 + 256;
};

struct TileCoord {
  int m0, n0;
};

// bm = rows per tile (128, or 256 for a CTA pair), m_off = this CTA's row offset inside the tile
__device__ __forceinline__ TileCoord tile_coord(int tile, int n_tiles, int bn, int bm, int m_off) {
  TileCoord t;
  t.m0 = (tile / n_tiles) * bm + m_off;
  t.n0 = (tile % n_tiles) * bn;
  return t;
}

// silu(x) = x * sigmoid(x) with two MUFU ops (ex2.approx, rcp.approx; ~1e-6 relative error)
__device__ __forceinline__ float silu_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2): one issue slot per two elements.  The epilogues are issue-bound CUDA-core
// code racing the tensor pipe (at K = 512 a 128 x 256 tile is ~4100 MMA cycles), so halving their instruction count is
// what keeps the narrow-K projections (to_out, w12) from waiting on the epilogue.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2u(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }
// silu(a) * b for two elements: the two MUFU ops per element stay scalar, everything around them is packed
__device__ __forceinline__ float2 silu_mul2(float2 a, float2 b) {
  const float2 t = __fmul2_rn(a, f2(-1.4426950408889634f, -1.4426950408889634f));
  float2 e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
  const float2 d = __fadd2_rn(e, f2(1.0f, 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
  return __fmul2_rn(__fmul2_rn(a, r), b);
}

template <int BN, int OUT_MODE, bool SWIGLU, bool CTA2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
            const GemmParams p) {
  using Cfg = GemmCfg<BN, CTA2, SWIGLU>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NSTG = Cfg::NSTG;
  constexpr int NSTG_G = Cfg::NSTG_G;
  constexpr int TILE_M = CTA2 ? 2 * BM : BM;
  constexpr int OUT_COLS = SWIGLU ? BN / 2 : BN;           // output columns per tile
  constexpr int CHUNKS = (OUT_COLS + 63) / 64;             // 64-column output chunks per tile
  static_assert(!SWIGLU || (BN % 128 == 0), "SwiGLU tiles need BN multiple of 128");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled TMA / UMMA tiles need a 1024-byte aligned base
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smC = smem + STAGES * Cfg::STAGE_BYTES;
  float* colvec = reinterpret_cast<float*>(smC + NSTG * STG_BYTES);   // [2][BN]: bias, colsum
  uint64_t* bars = reinterpret_cast<uint64_t*>(colvec + 2 * BN);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint64_t* res_bar = tempty_bar + 2;           // [2 groups][NSTG_G]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + NSTG);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;
  const bool has_res = (p.res != nullptr);
  // persistent schedule: a CTA (or CTA pair) walks tiles first, first + stride, ...
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int first_tile = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_stride = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int m_off = static_cast<int>(rank) * BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (OUT_MODE == OUT_BF16) tma_prefetch_desc(&tmOut);
    if (has_res) tma_prefetch_desc(&tmRes);
    for (int i = 0; i < NSTG; ++i) mbar_init(&res_bar[i], 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CTA2 ? 16 : 8);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CTA2) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();          // peer barriers initialised before any remote arrive / 2-SM TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer =====================================
    // (service warps run CONVERGED, an elected lane issues TMA / MMA / commit: divergent lane-0 loops cost ~16 instructions of
    //  uniform-register plumbing per tcgen05.mma — see pm_attn4.cu; -DPM_GEMM_LANE0=1 restores them)
#ifdef PM_GEMM_LANE0
#define GEMM_SERVICE_LANES (lane == 0)
#define GEMM_ONE
#else
#define GEMM_SERVICE_LANES true
#define GEMM_ONE if (elect_one())
#endif
    if (GEMM_SERVICE_LANES) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const TileCoord tc = tile_coord(tile, n_tiles, BN, TILE_M, m_off);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          GEMM_ONE {
            if (CTA2) {
              // both CTAs load their own A rows and their half of the W tile; all bytes are credited to the
              // leader's full barrier, which alone is armed (with the pair's total)
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              tma_load_2d_2sm(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * BK, tc.m0);
              tma_load_2d_2sm(smB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BK, tc.n0 + static_cast<int>(rank) * Cfg::B_ROWS);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              tma_load_2d(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * BK, tc.m0);
              tma_load_2d(smB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BK, tc.n0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =======================================
    if (GEMM_SERVICE_LANES && rank == 0) {        // in a CTA pair only the leader issues MMAs
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      // optional stall accounting (p.debug != nullptr): cycles the issuer spent waiting for accumulators / operands
      long long t_acc = 0, t_opr = 0;
      const long long t_begin = clock64();
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        long long t0 = clock64();
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        t_acc += clock64() - t0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          t0 = clock64();
          mbar_wait(&full_bar[stage], phase);
          t_opr += clock64() - t0;
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smA + stage * Cfg::A_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(smB + stage * Cfg::B_BYTES));
          GEMM_ONE {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in 16 B units
              if (CTA2) umma_ss_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees this smem stage (in both CTAs of a pair) once the MMAs have read it
            if (CTA2) umma_commit_2sm(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
            // accumulator complete -> epilogue (of both CTAs)
            if (kb == k_blocks - 1) {
              if (CTA2) umma_commit_2sm(&tfull_bar[as], 3); else umma_commit(&tfull_bar[as]);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      if (p.debug != nullptr && lane == 0) {
        long long* d = p.debug + 4 * static_cast<size_t>(blockIdx.x);
        d[0] = t_acc; d[1] = t_opr; d[2] = clock64() - t_begin; d[3] = k_blocks;
      }
    }
  } else {
    // =============================== epilogue (warps 2..9) ============================
    // Two groups of four warps (one warp per TMEM lane quarter each).  Group g drains the 64-column output
    // chunks c = g, g+2, ... of every tile, so each SM sub-partition always has two epilogue warps whose
    // TMEM / shared / global latencies overlap.  Each group owns NSTG_G staging buffers, two named barriers
    // and a leader thread that issues the TMA stores.
    const int ew = warp - 2;                    // 0..7
    const int grp = ew >> 2;                    // chunk parity this warp works on
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int et = ew * 32 + lane;              // 0..255 epilogue thread id
    const int row_in_tile = q * 32 + lane;
    const bool leader = ((ew & 3) == 0 && lane == 0);
    const uint32_t tmem_lane = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t bar_free = 2 + 2 * grp, bar_full = 3 + 2 * grp;     // named barriers of this group (128 threads)
    uint8_t* stage_base = smC + grp * NSTG_G * STG_BYTES;
    const uint32_t tempty_leader = CTA2 ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;   // leader CTA's tempty_bar[0]

    const int my_tiles = first_tile < total_tiles ? (total_tiles - first_tile + tile_stride - 1) / tile_stride : 0;
    constexpr int CPG0 = (CHUNKS + 1) / 2, CPG1 = CHUNKS / 2;      // chunks per tile of group 0 / 1
    const int cpg = grp == 0 ? CPG0 : CPG1;
    const int group_chunks = my_tiles * cpg;
    // residual prefetch by TMA into the staging buffer the result will be stored from
    auto issue_res = [&](int g) {
      const int ti = g / cpg, c = grp + 2 * (g % cpg);
      const TileCoord rc = tile_coord(first_tile + ti * tile_stride, n_tiles, BN, TILE_M, m_off);
      uint64_t* bar = &res_bar[grp * NSTG_G + g % NSTG_G];
      mbar_arrive_expect_tx(bar, STG_BYTES);
      tma_load_2d(stage_base + (g % NSTG_G) * STG_BYTES, &tmRes, bar, (SWIGLU ? rc.n0 / 2 : rc.n0) + c * 64,
                  p.res_mod > 0 ? rc.m0 % p.res_mod : rc.m0);
    };
    if (OUT_MODE == OUT_BF16 && has_res && leader && group_chunks > 0) issue_res(0);

    // per-row LayerNorm statistics, prefetched one tile ahead so their global-load latency is off the critical path
    float4 spf[4];
    auto prefetch_stats = [&](int tile_) {
      const int r_ = tile_coord(tile_, n_tiles, BN, TILE_M, m_off).m0 + row_in_tile;
      if (p.stats != nullptr && r_ < p.M) {
        if (p.stats_raw > 0) {
          const float4* sp = reinterpret_cast<const float4*>(p.stats + 2 * static_cast<size_t>(p.stats_raw) * r_);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (2 * i < p.stats_raw) spf[i] = sp[i];
        } else {
          const float2 t2 = *reinterpret_cast<const float2*>(p.stats + 2 * static_cast<size_t>(r_));
          spf[0] = make_float4(t2.x, t2.y, 0.0f, 0.0f);
        }
      }
    };
    if (first_tile < total_tiles) prefetch_stats(first_tile);

    int as = 0;
    uint32_t aphase = 0;
    int gl = 0;   // chunks stored so far by this group (staging buffer = gl % NSTG_G)
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
      const TileCoord tc = tile_coord(tile, n_tiles, BN, TILE_M, m_off);
      const int row = tc.m0 + row_in_tile;
      const bool row_ok = row < p.M;
      const int out_n0 = SWIGLU ? tc.n0 / 2 : tc.n0;

      // per-tile column vectors -> smem (all 8 warps)
      named_bar_sync(1, 256);
      for (int i = et; i < BN; i += 256) {
        const int col = tc.n0 + i;
        colvec[i] = (p.bias != nullptr && col < p.N) ? p.bias[col] : 0.0f;
        colvec[BN + i] = (p.colsum != nullptr && col < p.N) ? p.colsum[col] : 0.0f;
      }
      // LayerNorm statistics of this row were prefetched one tile ahead (spf); turn them into (mean, rstd)
      float mu = 0.0f, rstd = 1.0f;
      if (p.stats != nullptr && row_ok) {
        if (p.stats_raw > 0) {
          // p.stats_raw partial (sum, sum of squares) pairs per row, written by the epilogue of the GEMM that
          // produced A (one pair per N-tile and epilogue group); summed here in a fixed order -> deterministic
          float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (2 * i < p.stats_raw) {
              s1 += spf[i].x + spf[i].z;
              s2 += spf[i].y + spf[i].w;
            }
          }
          const float inv_k = 1.0f / static_cast<float>(p.K);
          mu = s1 * inv_k;
          const float var = fmaxf(s2 * inv_k - mu * mu, 0.0f);
          rstd = rsqrtf(var + p.ln_eps);
        } else {
          mu = spf[0].x;
          rstd = spf[0].y;
        }
      }
      if (tile + tile_stride < total_tiles) prefetch_stats(tile + tile_stride);
      float2 row_s1 = f2(0.0f, 0.0f), row_s2 = f2(0.0f, 0.0f);   // statistics of this thread's output columns (even | odd), for the next LayerNorm
      const float nrmu = -rstd * mu;        // LN fold: rstd * (acc - mu * colsum) + bias == acc * rstd + (nrmu * colsum + bias)
      const float* pos_row = nullptr;
      if (p.pos != nullptr && row_ok) pos_row = p.pos + static_cast<size_t>(row % p.pos_rows) * p.ld_pos;
      // ONE epilogue warp polls the accumulator's mbarrier, the other seven sleep in the named barrier that follows: a polling
      // warp costs issue slots and power for as long as the main loop of the tile runs (PM_GEMM_EPI_POLL_ALL=1: all eight poll)
#ifdef PM_GEMM_EPI_POLL_ALL
      named_bar_sync(1, 256);
      mbar_wait(&tfull_bar[as], aphase);
#else
      if (ew == 0) mbar_wait(&tfull_bar[as], aphase);
      named_bar_sync(1, 256);
#endif
      tc_fence_after();
      const uint32_t tacc = tmem_base + tmem_lane + as * BN;

#pragma unroll 1
      for (int c = grp; c < CHUNKS; c += 2) {
        const bool last_chunk = (c + 2 >= CHUNKS);
        uint8_t* stg = nullptr;
        if (OUT_MODE == OUT_BF16) {
          if (has_res) {
            // chunk gl-1's store has left its buffer: prefetch the residual of chunk gl+1 into it (a whole chunk
            // ahead of its use); this chunk's residual was requested one chunk ago
            if (leader) {
              tma_store_wait_read<0>();
              if (gl + 1 < group_chunks) issue_res(gl + 1);
            }
          } else {
            // staging buffer free? (its previous TMA store has finished reading shared memory)
            if (leader) tma_store_wait_read<NSTG_G - 1>();
            named_bar_sync(bar_free, 128);
          }
          stg = stage_base + (gl % NSTG_G) * STG_BYTES + row_in_tile * 128;
        }
        constexpr int HALVES = (OUT_COLS >= 64) ? 2 : 1;   // BN=32 tiles have a single 32-col half
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          const int oc = c * 64 + h * 32;                   // output column inside the tile
          float2 v2[16];                                   // the 32 output values of this thread, as 16 packed pairs
          float* const v = reinterpret_cast<float*>(v2);
          const float2 rstd2 = f2(rstd, rstd), nrmu2 = f2(nrmu, nrmu);
          if (SWIGLU) {
            uint32_t rg[32], rv[32];
            tmem_ld_x32(tacc + oc, rg);
            tmem_ld_x32(tacc + BN / 2 + oc, rv);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bg = *reinterpret_cast<const float4*>(&colvec[oc + j]);
              const float4 cg = *reinterpret_cast<const float4*>(&colvec[BN + oc + j]);
              const float4 bv = *reinterpret_cast<const float4*>(&colvec[BN / 2 + oc + j]);
              const float4 cv = *reinterpret_cast<const float4*>(&colvec[BN + BN / 2 + oc + j]);
              // LN fold: acc * rstd + (nrmu * colsum + bias), gate and value columns
              const float2 a0 = __ffma2_rn(f2u(rg[j], rg[j + 1]), rstd2, __ffma2_rn(nrmu2, f2(cg.x, cg.y), f2(bg.x, bg.y)));
              const float2 a1 = __ffma2_rn(f2u(rg[j + 2], rg[j + 3]), rstd2, __ffma2_rn(nrmu2, f2(cg.z, cg.w), f2(bg.z, bg.w)));
              const float2 b0 = __ffma2_rn(f2u(rv[j], rv[j + 1]), rstd2, __ffma2_rn(nrmu2, f2(cv.x, cv.y), f2(bv.x, bv.y)));
              const float2 b1 = __ffma2_rn(f2u(rv[j + 2], rv[j + 3]), rstd2, __ffma2_rn(nrmu2, f2(cv.z, cv.w), f2(bv.z, bv.w)));
              v2[j >> 1] = silu_mul2(a0, b0);
              v2[(j >> 1) + 1] = silu_mul2(a1, b1);
            }
          } else {
            uint32_t ra[32];
            tmem_ld_x32(tacc + oc, ra);
            tmem_ld_wait();
            if (p.stats != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = *reinterpret_cast<const float4*>(&colvec[oc + j]);
                const float4 cc = *reinterpret_cast<const float4*>(&colvec[BN + oc + j]);
                v2[j >> 1] = __ffma2_rn(f2u(ra[j], ra[j + 1]), rstd2, __ffma2_rn(nrmu2, f2(cc.x, cc.y), f2(bb.x, bb.y)));
                v2[(j >> 1) + 1] = __ffma2_rn(f2u(ra[j + 2], ra[j + 3]), rstd2, __ffma2_rn(nrmu2, f2(cc.z, cc.w), f2(bb.z, bb.w)));
              }
            } else {
              // no LayerNorm to fold (rstd = 1, mu = 0): acc + bias
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = *reinterpret_cast<const float4*>(&colvec[oc + j]);
                v2[j >> 1] = __fadd2_rn(f2u(ra[j], ra[j + 1]), f2(bb.x, bb.y));
                v2[(j >> 1) + 1] = __fadd2_rn(f2u(ra[j + 2], ra[j + 3]), f2(bb.z, bb.w));
              }
            }
            if (pos_row != nullptr) {
              const int col0 = tc.n0 + oc;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (col0 + j < p.N) {
                  const float4 pe = *reinterpret_cast<const float4*>(pos_row + col0 + j);
                  v[j] += pe.x; v[j + 1] += pe.y; v[j + 2] += pe.z; v[j + 3] += pe.w;
                }
              }
            }
          }
          if (last_chunk && h == HALVES - 1) {
            // all TMEM reads of this accumulator stage by this warp are done -> hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CTA2) mbar_arrive_cluster(tempty_leader + as * 8);   // the leader's MMA thread owns the accumulators of both CTAs
              else mbar_arrive(&tempty_bar[as]);
            }
          }

          if (OUT_MODE == OUT_BF16) {
            if (has_res && h == 0) mbar_wait(&res_bar[grp * NSTG_G + gl % NSTG_G], (gl / NSTG_G) & 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2* x2 = &v2[j * 4];
              uint4* slot = reinterpret_cast<uint4*>(stg + (((h * 4 + j) ^ (row_in_tile & 7)) << 4));
              if (has_res) {
                const uint4 r = *slot;
                x2[0] = __fadd2_rn(x2[0], f2(bf16lo_to_f32(r.x), bf16hi_to_f32(r.x)));
                x2[1] = __fadd2_rn(x2[1], f2(bf16lo_to_f32(r.y), bf16hi_to_f32(r.y)));
                x2[2] = __fadd2_rn(x2[2], f2(bf16lo_to_f32(r.z), bf16hi_to_f32(r.z)));
                x2[3] = __fadd2_rn(x2[3], f2(bf16lo_to_f32(r.w), bf16hi_to_f32(r.w)));
              }
              uint4 o;
              o.x = pack_bf16x2(x2[0].x, x2[0].y);
              o.y = pack_bf16x2(x2[1].x, x2[1].y);
              o.z = pack_bf16x2(x2[2].x, x2[2].y);
              o.w = pack_bf16x2(x2[3].x, x2[3].y);
              *slot = o;
              if (!SWIGLU && p.stats_out != nullptr) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  row_s1 = __fadd2_rn(row_s1, x2[e]);
                  row_s2 = __ffma2_rn(x2[e], x2[e], row_s2);
                }
              }
            }
          } else if (OUT_MODE == OUT_F32) {
            if (row_ok) {
              float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ld_out + tc.n0 + oc;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (tc.n0 + oc + j < p.N) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
            }
          } else if (OUT_MODE == OUT_UNPATCH_U8) {
            // column = (p1 * 8 + p2) * 3 + ch -> uint8 NHWC pixel out[b, h*8 + p1, w*8 + p2, ch] = the reference's
            // `restore` (reconstruct.py:11-16) of the clamped reconstruction, operation by operation in fp32:
            // x = clamp(v, -1, 1); t = (x + 1) * 0.5; u = uint8(trunc(255 * t)).  Four consecutive columns are four
            // consecutive bytes of one pixel row of the patch (24 % 4 == 0), so they leave as one 32-bit store.
            if (row_ok) {
              const int G = p.grid;
              const int tok = row % (G * G), bimg = row / (G * G);
              const int th = tok / G, tw = tok % G;
              const int S = G * 8;
              uint8_t* img = reinterpret_cast<uint8_t*>(p.out) + (static_cast<size_t>(bimg) * S + th * 8) * S * 3 + tw * 24;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int col = tc.n0 + oc + j;
                if (col < p.N) {
                  uint32_t w = 0;
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float x = fminf(fmaxf(v[j + e], -1.0f), 1.0f);
                    const float t = __fmul_rn(__fadd_rn(x, 1.0f), 0.5f);
                    w |= static_cast<uint32_t>(__float2int_rz(__fmul_rn(255.0f, t))) << (8 * e);
                  }
                  const int p1 = col / 24, r = col - p1 * 24;
                  *reinterpret_cast<uint32_t*>(img + static_cast<size_t>(p1) * S * 3 + r) = w;
                }
              }
            }
          } else {  // OUT_UNPATCH: column = (ch * 8 + p1) * 8 + p2  ->  out[b, ch, h*8 + p1, w*8 + p2]
            // The caller permutes the rows of W (and bias / colsum) from the reference's (p1 p2 c) order
            // (layers.py:150) to (c p1 p2): 8 consecutive columns are then 8 consecutive pixels of one image row, each
            // thread stores whole 32-byte sectors and a warp (32 consecutive tokens) writes 1 KB contiguous runs.
            // (Scattered 4-byte stores in the reference order ran this GEMM at 108 TFLOP/s, 1.2 % of the step.)
            if (row_ok) {
              const int C = p.channels, G = p.grid;                 // G tokens per image side, patch = 8
              const int tok = row % (G * G), bimg = row / (G * G);
              const int th = tok / G, tw = tok % G;
              const int S = G * 8;                                  // image side
              float* img = reinterpret_cast<float*>(p.out) + static_cast<size_t>(bimg) * C * S * S +
                           static_cast<size_t>(th * 8) * S + tw * 8;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int col = tc.n0 + oc + j;
                if (col < p.N) {
                  const int ch = col >> 6, p1 = (col >> 3) & 7, p2 = col & 7;
                  float4 o;
                  o.x = fminf(fmaxf(v[j], -1.0f), 1.0f);
                  o.y = fminf(fmaxf(v[j + 1], -1.0f), 1.0f);
                  o.z = fminf(fmaxf(v[j + 2], -1.0f), 1.0f);
                  o.w = fminf(fmaxf(v[j + 3], -1.0f), 1.0f);
                  __stcs(reinterpret_cast<float4*>(img + (static_cast<size_t>(ch) * S + p1) * S + p2), o);
                }
              }
            }
          }
        }

        if (OUT_MODE == OUT_BF16) {
          fence_proxy_async_smem();
          named_bar_sync(bar_full, 128);
          if (leader) {
            tma_store_2d(&tmOut, stage_base + (gl % NSTG_G) * STG_BYTES, out_n0 + c * 64, tc.m0);
            tma_store_commit();
          }
          ++gl;
        }
      }
      if (CHUNKS <= grp) {
        // this group has no chunk in such a narrow tile: it still has to release the accumulator stage
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTA2) mbar_arrive_cluster(tempty_leader + as * 8);
          else mbar_arrive(&tempty_bar[as]);
        }
      }
      if (OUT_MODE == OUT_BF16 && !SWIGLU && p.stats_out != nullptr && row_ok) {
        // slot (n_tile, group) of this row: [M, 2 * n_tiles, 2] — no atomics, no zero-fill, fixed summation order
        const int slot = (tile % n_tiles) * 2 + grp;
        *reinterpret_cast<float2*>(p.stats_out + (static_cast<size_t>(row) * (2 * n_tiles) + slot) * 2) = make_float2(row_s1.x + row_s1.y, row_s2.x + row_s2.y);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (OUT_MODE == OUT_BF16 && leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();          // no CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------
static int g_num_sms = 0;
int pm_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int OUT_MODE, bool SWIGLU, bool CTA2>
static int launch_gemm(const GemmParams& p_in, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTA2, SWIGLU>;
  static_assert(Cfg::STAGES >= 2, "not enough shared memory for a pipeline");
  GemmParams p = p_in;
  CUtensorMap tmA, tmB, tmOut, tmRes;
  int rc;
  if ((rc = pm_make_tmap_2d(&tmA, p.a, 2, p.M, p.K, p.lda, BM, BK)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_2d(&tmB, p.w, 2, p.N, p.K, p.ldw, Cfg::B_ROWS, BK)) != PM_OK) return rc;
  const int out_cols = SWIGLU ? p.N / 2 : p.N;
  if (OUT_MODE == OUT_BF16) {
    if ((rc = pm_make_tmap_2d(&tmOut, p.out, 2, p.M, out_cols, p.ld_out, BM, 64)) != PM_OK) return rc;
  } else {
    tmOut = tmA;
  }
  p.N_out = out_cols;
  if (p.res != nullptr) {
    if (OUT_MODE != OUT_BF16) return PM_ERR_INVALID;
    if (p.res_mod < 0 || (p.res_mod % BM) != 0) return PM_ERR_INVALID;     // a 128-row tile must not wrap around
    if ((rc = pm_make_tmap_2d(&tmRes, p.res, 2, p.res_mod > 0 ? p.res_mod : p.M, out_cols, p.ld_res, BM, 64)) != PM_OK) return rc;
  } else {
    tmRes = tmA;
  }
  auto kern = gemm_kernel<BN, OUT_MODE, SWIGLU, CTA2>;
  static bool attr_done[PM_MAX_DEVICES] = {};
  if ((rc = pm_ensure_dyn_smem(kern, Cfg::SMEM_BYTES, attr_done)) != 0) return rc;
  constexpr int TILE_M = CTA2 ? 2 * BM : BM;
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M, n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int units = CTA2 ? pm_num_sms() / 2 : pm_num_sms();       // CTAs or CTA pairs that fit the machine
  int grid = tiles < units ? tiles : units;
  if (p.max_ctas > 0 && grid > p.max_ctas) grid = p.max_ctas;
  if (!CTA2) {
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmOut, tmRes, p);
    return static_cast<int>(cudaGetLastError());
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmRes, p));
}

int pm_gemm_launch(const GemmParams& p, int bn, int out_mode, int swiglu, cudaStream_t stream) {
  if (p.a == nullptr || p.w == nullptr || p.out == nullptr || p.M <= 0 || p.N <= 0 || p.K <= 0)
    return PM_ERR_INVALID;
  // CTA pairs (256 x 256 tiles) whenever there is enough work to fill the machine with them
  const bool pair = p.cta_pair != 0 && bn == 256 && p.M >= 256;
  if (swiglu) {
    if (out_mode != OUT_BF16 || bn != 256 || (p.N % 256) != 0) return PM_ERR_INVALID;
    return pair ? launch_gemm<256, OUT_BF16, true, true>(p, stream) : launch_gemm<256, OUT_BF16, true, false>(p, stream);
  }
  if (out_mode == OUT_BF16) {
    switch (bn) {
      case 256: return pair ? launch_gemm<256, OUT_BF16, false, true>(p, stream) : launch_gemm<256, OUT_BF16, false, false>(p, stream);
      case 128: return launch_gemm<128, OUT_BF16, false, false>(p, stream);
      case 64:  return launch_gemm<64, OUT_BF16, false, false>(p, stream);
      default:  return PM_ERR_INVALID;
    }
  }
  if (out_mode == OUT_F32) {
    switch (bn) {
      case 32:  return launch_gemm<32, OUT_F32, false, false>(p, stream);
      case 64:  return launch_gemm<64, OUT_F32, false, false>(p, stream);
      case 128: return launch_gemm<128, OUT_F32, false, false>(p, stream);
      case 256: return pair ? launch_gemm<256, OUT_F32, false, true>(p, stream) : launch_gemm<256, OUT_F32, false, false>(p, stream);
      default:  return PM_ERR_INVALID;
    }
  }
  if (out_mode == OUT_UNPATCH) {
    if (p.patch != 8 || p.N != 64 * p.channels || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return PM_ERR_INVALID;
    switch (bn) {
      case 64:  return launch_gemm<64, OUT_UNPATCH, false, false>(p, stream);
      case 192: return launch_gemm<192, OUT_UNPATCH, false, false>(p, stream);
      default:  return PM_ERR_INVALID;
    }
  }
  if (out_mode == OUT_UNPATCH_U8) {
    if (p.patch != 8 || p.channels != 3 || p.N != 192 || (reinterpret_cast<uintptr_t>(p.out) & 3) != 0) return PM_ERR_INVALID;
    switch (bn) {
      case 64:  return launch_gemm<64, OUT_UNPATCH_U8, false, false>(p, stream);
      case 192: return launch_gemm<192, OUT_UNPATCH_U8, false, false>(p, stream);
      default:  return PM_ERR_INVALID;
    }
  }
  return PM_ERR_INVALID;
}

}  // namespace pm
