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
// Structure per CTA (192 threads, 1 CTA / SM, grid = min(#tiles, #SMs)):
//   warp 0 lane 0 : TMA producer  — A and W tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1 lane 0 : MMA issuer    — tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM,
//                                   two accumulator stages so the epilogue overlaps the next tile
//   warps 2..5    : epilogue      — tcgen05.ld -> registers -> fused math -> swizzled smem
//                                   staging -> TMA store (bf16 outputs); residual tiles are
//                                   TMA-prefetched into the same staging buffers.
#include "pm_common.cuh"
#include "pm_kernels.h"

namespace pm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NSTG = 4;                       // epilogue staging buffers (128 rows x 128 B)
constexpr int STG_BYTES = BM * 128;           // 16 KB
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // 227 KB total: staging + column vectors + barriers leave this much for the ring
  static constexpr int RING_BUDGET = 232448 - NSTG * STG_BYTES - 2 * BN * 4 - 1024 - 1024;
  static constexpr int STAGES_RAW = RING_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + NSTG * STG_BYTES + 2 * BN * 4 + 1024;
};

struct TileCoord {
  int m0, n0;
};

__device__ __forceinline__ TileCoord tile_coord(int tile, int n_tiles, int bn) {
  TileCoord t;
  t.m0 = (tile / n_tiles) * BM;
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

template <int BN, int OUT_MODE, bool SWIGLU>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
            const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int OUT_COLS = SWIGLU ? BN / 2 : BN;           // output columns per tile
  constexpr int CHUNKS = (OUT_COLS + 63) / 64;             // 64-column output chunks per tile
  static_assert(!SWIGLU || (BN % 128 == 0), "SwiGLU tiles need BN multiple of 128");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smC = smem + STAGES * Cfg::STAGE_BYTES;
  float* colvec = reinterpret_cast<float*>(smC + NSTG * STG_BYTES);   // [2][BN]: bias, colsum
  uint64_t* bars = reinterpret_cast<uint64_t*>(colvec + 2 * BN);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint64_t* res_bar = tempty_bar + 2;           // [NSTG]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + NSTG);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;
  const bool has_res = (p.res != nullptr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (OUT_MODE == OUT_BF16) tma_prefetch_desc(&tmOut);
    if (has_res) tma_prefetch_desc(&tmRes);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);   // one arrival per epilogue warp
    }
    for (int i = 0; i < NSTG; ++i) mbar_init(&res_bar[i], 1);
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

  if (warp == 0) {
    // =============================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(tile, n_tiles, BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * BK, tc.m0);
          tma_load_2d(smB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BK, tc.n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smA + stage * Cfg::A_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(smB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in 16 B units
            umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees this smem stage once the MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);        // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue (warps 2..5) ============================
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int et = (warp - 2) * 32 + lane;      // 0..127 epilogue thread id
    const int row_in_tile = q * 32 + lane;
    const bool leader = (warp == 2 && lane == 0);
    const uint32_t tmem_lane = static_cast<uint32_t>(q * 32) << 16;

    const int my_tiles = (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int total_chunks = my_tiles * CHUNKS;

    // residual prefetch: chunk g lives in staging buffer g % NSTG
    auto issue_res = [&](int g) {
      const int ti = g / CHUNKS, c = g % CHUNKS;
      const TileCoord tc = tile_coord(blockIdx.x + ti * gridDim.x, n_tiles, BN);
      const int b = g % NSTG;
      mbar_arrive_expect_tx(&res_bar[b], STG_BYTES);
      tma_load_2d(smC + b * STG_BYTES, &tmRes, &res_bar[b], (SWIGLU ? tc.n0 / 2 : tc.n0) + c * 64, tc.m0);
    };
    if (OUT_MODE == OUT_BF16 && has_res && leader) {
      for (int g = 0; g < NSTG && g < total_chunks; ++g) issue_res(g);
    }

    int as = 0;
    uint32_t aphase = 0;
    int g = 0;   // running chunk counter
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = tile_coord(tile, n_tiles, BN);
      const int row = tc.m0 + row_in_tile;
      const bool row_ok = row < p.M;

      // per-tile column vectors -> smem
      named_bar_sync(1, 128);
      for (int i = et; i < BN; i += 128) {
        const int col = tc.n0 + i;
        colvec[i] = (p.bias != nullptr && col < p.N) ? p.bias[col] : 0.0f;
        colvec[BN + i] = (p.colsum != nullptr && col < p.N) ? p.colsum[col] : 0.0f;
      }
      float mu = 0.0f, rstd = 1.0f;
      if (p.stats != nullptr && row_ok) {
        const float2 st = *reinterpret_cast<const float2*>(p.stats + 2 * static_cast<size_t>(row));
        mu = st.x;
        rstd = st.y;
      }
      const float nrmu = -rstd * mu;     // LN fold: rstd * (acc - mu * colsum) + bias == acc * rstd + (nrmu * colsum + bias)
      const float* pos_row = nullptr;
      if (p.pos != nullptr && row_ok) pos_row = p.pos + static_cast<size_t>(row % p.pos_rows) * p.ld_pos;
      named_bar_sync(1, 128);

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + tmem_lane + as * BN;

#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c, ++g) {
        // ---- 64 output values of this thread's row for this chunk ----
        float v[64];
        constexpr int HALVES = (OUT_COLS >= 64) ? 2 : 1;   // BN=32 tiles have a single 32-col half
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          const int oc = c * 64 + h * 32;                   // output column inside the tile
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
              const float bgs[4] = {bg.x, bg.y, bg.z, bg.w}, cgs[4] = {cg.x, cg.y, cg.z, cg.w};
              const float bvs[4] = {bv.x, bv.y, bv.z, bv.w}, cvs[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float a = fmaf(__uint_as_float(rg[j + t]), rstd, fmaf(nrmu, cgs[t], bgs[t]));
                const float b = fmaf(__uint_as_float(rv[j + t]), rstd, fmaf(nrmu, cvs[t], bvs[t]));
                v[h * 32 + j + t] = silu_f(a) * b;
              }
            }
          } else {
            uint32_t ra[32];
            tmem_ld_x32(tacc + oc, ra);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(&colvec[oc + j]);
              const float4 cc = *reinterpret_cast<const float4*>(&colvec[BN + oc + j]);
              const float bbs[4] = {bb.x, bb.y, bb.z, bb.w}, ccs[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
              for (int t = 0; t < 4; ++t)
                v[h * 32 + j + t] = fmaf(__uint_as_float(ra[j + t]), rstd, fmaf(nrmu, ccs[t], bbs[t]));
            }
            if (pos_row != nullptr) {
              const int col0 = tc.n0 + oc;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (col0 + j < p.N) {
                  const float4 pe = *reinterpret_cast<const float4*>(pos_row + col0 + j);
                  v[h * 32 + j] += pe.x; v[h * 32 + j + 1] += pe.y;
                  v[h * 32 + j + 2] += pe.z; v[h * 32 + j + 3] += pe.w;
                }
              }
            }
          }
        }
        if (c == CHUNKS - 1) {
          // all TMEM reads of this accumulator stage are done -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }

        if (OUT_MODE == OUT_BF16) {
          const int b = g % NSTG;
          uint8_t* stg = smC + b * STG_BYTES + row_in_tile * 128;
          if (has_res) {
            mbar_wait(&res_bar[b], (g / NSTG) & 1);
          } else {
            if (leader) tma_store_wait_read<NSTG - 1>();
            named_bar_sync(2, 128);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4* slot = reinterpret_cast<uint4*>(stg + ((j ^ (row_in_tile & 7)) << 4));
            float* x = &v[j * 8];
            if (has_res) {
              const uint4 r = *slot;
              x[0] += bf16lo_to_f32(r.x); x[1] += bf16hi_to_f32(r.x);
              x[2] += bf16lo_to_f32(r.y); x[3] += bf16hi_to_f32(r.y);
              x[4] += bf16lo_to_f32(r.z); x[5] += bf16hi_to_f32(r.z);
              x[6] += bf16lo_to_f32(r.w); x[7] += bf16hi_to_f32(r.w);
            }
            uint4 o;
            o.x = pack_bf16x2(x[0], x[1]);
            o.y = pack_bf16x2(x[2], x[3]);
            o.z = pack_bf16x2(x[4], x[5]);
            o.w = pack_bf16x2(x[6], x[7]);
            *slot = o;
          }
          fence_proxy_async_smem();
          named_bar_sync(3, 128);
          if (leader) {
            tma_store_2d(&tmOut, smC + b * STG_BYTES, (SWIGLU ? tc.n0 / 2 : tc.n0) + c * 64, tc.m0);
            tma_store_commit();
            if (has_res) {
              // buffer of the previous chunk becomes free once its store has read smem;
              // refill it with the residual tile NSTG chunks ahead
              if (g >= 1 && g - 1 + NSTG < total_chunks) {
                tma_store_wait_read<1>();
                issue_res(g - 1 + NSTG);
              }
            }
          }
        } else if (OUT_MODE == OUT_F32) {
          if (row_ok) {
            float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ld_out + tc.n0 + c * 64;
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              if (j < OUT_COLS && tc.n0 + c * 64 + j < p.N) {
                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
            }
          }
        } else {  // OUT_UNPATCH: column = (p1 * P + p2) * C + ch  ->  out[b, ch, h*P + p1, w*P + p2]
          if (row_ok) {
            const int P = p.patch, C = p.channels, G = p.grid;   // G tokens per image side
            const int tok = row % (G * G), bimg = row / (G * G);
            const int th = tok / G, tw = tok % G;
            const int S = G * P;                                  // image side
            float* img = reinterpret_cast<float*>(p.out) + static_cast<size_t>(bimg) * C * S * S;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const int col = tc.n0 + c * 64 + j;
              if (j < OUT_COLS && col < p.N) {
                const int ch = col % C, pp = col / C;
                const int p1 = pp / P, p2 = pp % P;
                const float x = fminf(fmaxf(v[j], -1.0f), 1.0f);
                img[(static_cast<size_t>(ch) * S + th * P + p1) * S + tw * P + p2] = x;
              }
            }
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (OUT_MODE == OUT_BF16 && leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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

template <int BN, int OUT_MODE, bool SWIGLU>
static int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static_assert(Cfg::STAGES >= 2, "not enough shared memory for a pipeline");
  CUtensorMap tmA, tmB, tmOut, tmRes;
  int rc;
  if ((rc = pm_make_tmap_2d(&tmA, p.a, 2, p.M, p.K, p.lda, BM, BK)) != PM_OK) return rc;
  if ((rc = pm_make_tmap_2d(&tmB, p.w, 2, p.N, p.K, p.ldw, BN, BK)) != PM_OK) return rc;
  const int out_cols = SWIGLU ? p.N / 2 : p.N;
  if (OUT_MODE == OUT_BF16) {
    if ((rc = pm_make_tmap_2d(&tmOut, p.out, 2, p.M, out_cols, p.ld_out, BM, 64)) != PM_OK) return rc;
  } else {
    tmOut = tmA;
  }
  if (p.res != nullptr) {
    if (OUT_MODE != OUT_BF16) return PM_ERR_INVALID;
    if ((rc = pm_make_tmap_2d(&tmRes, p.res, 2, p.M, out_cols, p.ld_res, BM, 64)) != PM_OK) return rc;
  } else {
    tmRes = tmA;
  }
  auto kern = gemm_kernel<BN, OUT_MODE, SWIGLU>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  int grid = tiles < pm_num_sms() ? tiles : pm_num_sms();
  if (p.max_ctas > 0 && grid > p.max_ctas) grid = p.max_ctas;
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmOut, tmRes, p);
  return static_cast<int>(cudaGetLastError());
}

int pm_gemm_launch(const GemmParams& p, int bn, int out_mode, int swiglu, cudaStream_t stream) {
  if (p.a == nullptr || p.w == nullptr || p.out == nullptr || p.M <= 0 || p.N <= 0 || p.K <= 0)
    return PM_ERR_INVALID;
  if (swiglu) {
    if (out_mode != OUT_BF16 || bn != 256 || (p.N % 256) != 0) return PM_ERR_INVALID;
    return launch_gemm<256, OUT_BF16, true>(p, stream);
  }
  if (out_mode == OUT_BF16) {
    switch (bn) {
      case 256: return launch_gemm<256, OUT_BF16, false>(p, stream);
      case 128: return launch_gemm<128, OUT_BF16, false>(p, stream);
      case 64:  return launch_gemm<64, OUT_BF16, false>(p, stream);
      default:  return PM_ERR_INVALID;
    }
  }
  if (out_mode == OUT_F32) {
    switch (bn) {
      case 32:  return launch_gemm<32, OUT_F32, false>(p, stream);
      case 64:  return launch_gemm<64, OUT_F32, false>(p, stream);
      case 128: return launch_gemm<128, OUT_F32, false>(p, stream);
      case 256: return launch_gemm<256, OUT_F32, false>(p, stream);
      default:  return PM_ERR_INVALID;
    }
  }
  if (out_mode == OUT_UNPATCH) {
    switch (bn) {
      case 64:  return launch_gemm<64, OUT_UNPATCH, false>(p, stream);
      case 192: return launch_gemm<192, OUT_UNPATCH, false>(p, stream);
      default:  return PM_ERR_INVALID;
    }
  }
  return PM_ERR_INVALID;
}

}  // namespace pm
