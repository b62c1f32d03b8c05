// pm_vq.cu — factorised, l2-normalised vector quantizer on tcgen05 (sm_100a), e_dim = 32.
//
// Replaces VectorQuantizer.forward / decode_from_indice (reference stage1/quantize.py:18-44):
//     zn = l2norm(z);  en = l2norm(E);  d = |zn|^2 + |en|^2 - 2 zn.en^T;  idx = argmin_j d
//     z_q = l2norm(E[idx]);  loss = (1 + beta) * mean((z_q - zn)^2);  out = zn + (z_q - zn)
// Both operands are unit vectors, so argmin_j d == argmax_j zn.en (first index on ties, as
// torch.argmin); the [M, n_e] distance matrix is never materialised.
//
// Exactness: the contraction runs on bf16 tensor cores with BOTH operands split into
// hi + lo bf16 halves, and all four partial products accumulated in fp32 TMEM:
//     A row (K = 128) = [ z_hi | z_hi | z_lo | z_lo ]     (two 128B-swizzled K slabs)
//     B row (K = 64)  = [ e_hi | e_lo ]                   (ONE slab, reused for both A slabs)
// which reproduces the fp32 dot product to ~1e-7 (bf16 x bf16 products are exact in fp32).
//
// vq_main_kernel: work item = (256-row tile, codebook split).  352 threads:
//   warp 0 lane 0 : TMA producer — streams the packed codebook [n_e, 64] bf16 (1 MB for 8192
//                   codes) through a 6-stage smem ring of 128-code tiles
//   warps 1, 10   : MMA issuers  — one thread per 128-row half: 8 k-steps of 128x128x16 per code tile into
//                   double-buffered TMEM accumulators (4 x 128 columns)
//   warps 2..9    : one thread per latent row: normalise z, write the split A tile into swizzled
//                   smem, then drain the accumulators with a running (max, first index) pair.
// With splits == 1 the same threads finish the row in place (gather, straight-through output,
// squared-error and usage-histogram reduction); otherwise vq_finalize_kernel merges the splits.
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <stdlib.h>

namespace pm {

constexpr int VQ_D = 32;                 // e_dim
constexpr int VQ_BM = 256;               // rows per work item (two UMMA M=128 halves)
constexpr int VQ_BN = 128;               // codes per tile
constexpr int VQ_BSTAGES = 6;
constexpr int VQ_A_SLAB = 128 * 128;     // 16 KB: 128 rows x 64 bf16
constexpr int VQ_B_BYTES = VQ_BN * 128;  // 16 KB
constexpr int VQ_THREADS = 352;            // TMA warp, MMA warp (row half 0), 8 drain warps, MMA warp (row half 1)
constexpr int VQ_SMEM_BYTES = 1024 + 4 * VQ_A_SLAB + VQ_BSTAGES * VQ_B_BYTES + 512;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// normalise one 32-d row held in registers: x / max(||x||, 1e-12)   (quantize.py:5-6)
__device__ __forceinline__ void l2norm32(float (&x)[VQ_D]) {
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < VQ_D; ++i) ss = fmaf(x[i], x[i], ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int i = 0; i < VQ_D; ++i) x[i] *= inv;
}

__device__ __forceinline__ void load_row32(const float* __restrict__ src, float (&x)[VQ_D]) {
#pragma unroll
  for (int i = 0; i < VQ_D; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(src + i);
    x[i] = t.x; x[i + 1] = t.y; x[i + 2] = t.z; x[i + 3] = t.w;
  }
}

// ----------------------------------------------------------------------------------------------
// codebook prep: en = l2norm(E) (fp32) and the packed tensor-core operand [e_hi | e_lo] (bf16)
// (the reference re-normalises the whole codebook on every forward, quantize.py:21)
// ----------------------------------------------------------------------------------------------
__global__ void vq_codebook_prep_kernel(const float* __restrict__ E, int n_e, float* __restrict__ en,
                                        __nv_bfloat16* __restrict__ packed) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_e) return;
  float x[VQ_D];
  load_row32(E + static_cast<size_t>(j) * VQ_D, x);
  l2norm32(x);
#pragma unroll
  for (int i = 0; i < VQ_D; i += 4)
    *reinterpret_cast<float4*>(en + static_cast<size_t>(j) * VQ_D + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < VQ_D; i += 2) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[i]), h1 = __float2bfloat16_rn(x[i + 1]);
    hi[i >> 1] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
    lo[i >> 1] = pack_bf16x2(x[i] - __bfloat162float(h0), x[i + 1] - __bfloat162float(h1));
  }
  uint4* dst = reinterpret_cast<uint4*>(packed + static_cast<size_t>(j) * 64);
#pragma unroll
  for (int c = 0; c < 4; ++c) dst[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
#pragma unroll
  for (int c = 0; c < 4; ++c) dst[4 + c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
}

// ----------------------------------------------------------------------------------------------
// per-row finish: gather, straight-through value, loss / histogram partials  (quantize.py:29-36)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float vq_finish_row(const VqParams& p, int row, int idx) {
  float zn[VQ_D], e[VQ_D];
  load_row32(p.z + static_cast<size_t>(row) * p.ldz, zn);
  l2norm32(zn);
  load_row32(p.en + static_cast<size_t>(idx) * VQ_D, e);   // == l2norm(E[idx])
  float sse = 0.0f;
  float out[VQ_D];
#pragma unroll
  for (int i = 0; i < VQ_D; ++i) {
    const float d = e[i] - zn[i];
    sse = fmaf(d, d, sse);
    out[i] = zn[i] + d;                                     // z + (z_q - z).detach(), forward value
  }
  if (p.zq != nullptr) {
    float* dst = p.zq + static_cast<size_t>(row) * VQ_D;
#pragma unroll
    for (int i = 0; i < VQ_D; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(out[i], out[i + 1], out[i + 2], out[i + 3]);
  }
  if (p.zq_split != nullptr) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < VQ_D; i += 2) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(out[i]), h1 = __float2bfloat16_rn(out[i + 1]);
      hi[i >> 1] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
      lo[i >> 1] = pack_bf16x2(out[i] - __bfloat162float(h0), out[i + 1] - __bfloat162float(h1));
    }
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.zq_split) + static_cast<size_t>(row) * 64);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[4 + c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
  }
  if (p.idx != nullptr) p.idx[row] = static_cast<long long>(idx);
  if (p.hist != nullptr) atomicAdd(p.hist + idx, 1ull);
  return sse;
}

// ----------------------------------------------------------------------------------------------
// vq_exact4_kernel: the round-1 kernel (4 bf16 partial products = fp32-exact scores, 4x the algorithmic MMA work).
// Kept selectable (PM_VQ_MODE=4) as the A/B reference of vq_main_kernel below.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_exact4_kernel(const __grid_constant__ CUtensorMap tmB, const VqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                                 // [rowhalf 2][slab 2][16 KB]
  uint8_t* smB = smem + 4 * VQ_A_SLAB;                 // [VQ_BSTAGES][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + VQ_BSTAGES * VQ_B_BYTES);
  uint64_t* b_full = bars;                             // [VQ_BSTAGES]
  uint64_t* b_empty = bars + VQ_BSTAGES;               // [VQ_BSTAGES]
  uint64_t* t_full = bars + 2 * VQ_BSTAGES;            // [2 stages][2 row halves]
  uint64_t* t_empty = t_full + 4;                      // [2 stages][2 row halves]
  uint64_t* a_full = t_empty + 4;                      // [1]  A tile written (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int row_tiles = (p.M + VQ_BM - 1) / VQ_BM;
  const int items = row_tiles * p.splits;
  const int codes_per_split = p.n_e / p.splits;        // host guarantees divisibility by VQ_BN... or tail masked
  const int ntiles = (codes_per_split + VQ_BN - 1) / VQ_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < VQ_BSTAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 2);       // one commit per MMA issuer (row half)
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);       // one arrival per drain warp of that row half
    }
    mbar_init(a_full, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Service warps run CONVERGED and an elected lane issues TMA / MMA / commit (PM_VQ_LANE0=1: the old divergent lane-0 loops, in
  // which every tcgen05.mma costs ~16 instructions of uniform-register plumbing — see pm_attn4.cu)
#ifdef PM_VQ_LANE0
#define VQ_SERVICE_LANES (lane == 0)
#define VQ_ONE
#else
#define VQ_SERVICE_LANES true
#define VQ_ONE if (elect_one())
#endif
  if (warp == 0) {
    if (VQ_SERVICE_LANES) {
      int st = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int split = item % p.splits;
        const int code0 = split * codes_per_split;
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&b_empty[st], ph ^ 1);
          VQ_ONE {
            mbar_arrive_expect_tx(&b_full[st], VQ_B_BYTES);
            tma_load_2d(smB + st * VQ_B_BYTES, &tmB, &b_full[st], 0, code0 + t * VQ_BN);
          }
          if (++st == VQ_BSTAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // two MMA issuer threads, one per 128-row half: the 128x128x16 MMAs are short (64 tensor cycles), a single
    // issuing thread cannot keep the pipe fed
    if (VQ_SERVICE_LANES) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, VQ_BN, 0, 0);
      const int rh = (warp == 1) ? 0 : 1;
      int st = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      uint32_t item_ph = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        mbar_wait(a_full, item_ph);          // A tile of this item is in smem
        item_ph ^= 1;
        tc_fence_after();
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&t_empty[as * 2 + rh], aph ^ 1);
          mbar_wait(&b_full[st], ph);
          tc_fence_after();
          const uint64_t db = umma_desc_sw128(smem_u32(smB + st * VQ_B_BYTES));
          const uint32_t tacc = tmem_base + (as * 2 + rh) * VQ_BN;
          VQ_ONE {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              // A k-step k: slab k/4 (z_hi|z_hi , z_lo|z_lo), 32 B step inside; B: [e_hi|e_lo] k%4
              const uint64_t da = umma_desc_sw128(smem_u32(smA + (rh * 2 + (k >> 2)) * VQ_A_SLAB)) + 2 * (k & 3);
              umma_ss(tacc, da, db + 2 * (k & 3), idesc, k != 0 ? 1u : 0u);
            }
            umma_commit(&b_empty[st]);
            umma_commit(&t_full[as * 2 + rh]);
          }
          if (++st == VQ_BSTAGES) { st = 0; ph ^= 1; }
          if (++as == 2) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp >= 2 && warp < 10) {
    const int q = warp & 3;
    const int rh = (warp - 2) >> 2;                    // row half 0/1
    const int r_in_half = q * 32 + lane;               // 0..127
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    int as = 0;
    uint32_t aph = 0;
    double sse_acc = 0.0;

    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int rt = item / p.splits, split = item % p.splits;
      const int row = rt * VQ_BM + rh * 128 + r_in_half;
      const bool row_ok = row < p.M;
      const int code0 = split * codes_per_split;

      // ---- A tile: normalise, split hi/lo, write [hi|hi] / [lo|lo] into the swizzled slabs ----
      // (all MMAs of the previous item have completed: its last t_full was waited on below)
      {
        float x[VQ_D];
        if (row_ok) {
          load_row32(p.z + static_cast<size_t>(row) * p.ldz, x);
          l2norm32(x);
        } else {
#pragma unroll
          for (int i = 0; i < VQ_D; ++i) x[i] = 0.0f;
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < VQ_D; i += 2) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(x[i]), h1 = __float2bfloat16_rn(x[i + 1]);
          hi[i >> 1] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
          lo[i >> 1] = pack_bf16x2(x[i] - __bfloat162float(h0), x[i + 1] - __bfloat162float(h1));
        }
        uint8_t* a_hi = smA + (rh * 2 + 0) * VQ_A_SLAB + r_in_half * 128;
        uint8_t* a_lo = smA + (rh * 2 + 1) * VQ_A_SLAB + r_in_half * 128;
        const int sw = r_in_half & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int s = (c & 3) * 4;
          *reinterpret_cast<uint4*>(a_hi + ((c ^ sw) << 4)) = make_uint4(hi[s], hi[s + 1], hi[s + 2], hi[s + 3]);
          *reinterpret_cast<uint4*>(a_lo + ((c ^ sw) << 4)) = make_uint4(lo[s], lo[s + 1], lo[s + 2], lo[s + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full);
      }

      // ---- drain: running (max, first index) over this split's codes ----
      float best = -INFINITY;
      int bidx = code0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&t_full[as * 2 + rh], aph);
        tc_fence_after();
        const uint32_t tacc = tmem_base + lane_off + (as * 2 + rh) * VQ_BN;
        const int cbase = code0 + t * VQ_BN;
        const bool tail = (cbase + VQ_BN > code0 + codes_per_split) || (cbase + VQ_BN > p.n_e);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t r[32];
          tmem_ld_x32(tacc + cc * 32, r);
          tmem_ld_wait();
          if (cc == 3) {
            // last TMEM read of this accumulator stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[as * 2 + rh]);
          }
          if (tail) {
            // codes beyond this split / the codebook (TMA zero-filled rows) must never win
            const int lim = min(code0 + codes_per_split, p.n_e);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cbase + cc * 32 + i >= lim) r[i] = 0xff800000u;   // -inf
          }
          // chunk maximum with a short dependency chain (3-input max tree); the running (max, first index) pair is
          // only touched when this chunk beats it — rare after the first few tiles — so the 8192-long compare/select
          // chain of a naive scan never forms
          float m8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            m8[i] = fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])),
                          fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
          const float cm = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
          if (cm > best) {          // strict: an equal value in a later chunk never replaces an earlier index (torch.argmin rule)
            best = cm;
            int first = 31;
#pragma unroll
            for (int i = 30; i >= 0; --i)
              if (__uint_as_float(r[i]) == cm) first = i;
            bidx = cbase + cc * 32 + first;
          }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }

      if (p.splits == 1) {
        float sse = 0.0f;
        if (row_ok) sse = vq_finish_row(p, row, bidx);
        sse_acc += static_cast<double>(warp_sum(sse));
      } else if (row_ok) {
        p.cand_val[static_cast<size_t>(split) * p.M + row] = best;
        p.cand_idx[static_cast<size_t>(split) * p.M + row] = bidx;
      }
    }
    if (p.splits == 1 && lane == 0 && p.sse != nullptr && sse_acc != 0.0) atomicAdd(p.sse, sse_acc);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------------
// vq_main_kernel (round 2): 1x the algorithmic MMA work + exact re-scoring of a proven candidate set.
//
// The tensor cores contract only the bf16 HIGH halves: a_j = z_hi . e_hi,j (K = 32, two 128x128x16 MMAs per
// 128-code tile instead of eight; the codebook streams through shared memory as 64-byte rows, half the bytes).  With unit
// vectors and round-to-nearest bf16 (|x - x_hi| <= 2^-8 |x|)
//     |z.e_j - a_j| <= 2^-8 (2 + 2^-8) sum_k |z_k||e_jk| + (fp32 accumulation) <= eps = 0.007846     (Cauchy-Schwarz)
// so the exact arg-max j* satisfies a_j* >= max_j a_j - 2 eps.  The drain threads (one per latent row) keep the running
// approximate maximum `best` and push every code whose score is within VQ_DELTA >= 2 eps of the running maximum when it
// goes by — (score, index) packed into one word: the low 13 mantissa bits carry the index — onto a 32-entry ring in
// shared memory (about eight pushes per row for random data).  After the scan, the entries still within VQ_DELTA (+ the
// packing error) of the FINAL maximum (1-2 per row) are re-scored against the fp32 normalised codebook (fp32 FMA
// chains, ~1e-7) and the largest wins, the lowest index among scores within VQ_TIE of it — the reference's torch.argmin
// over d = |zn|^2 + |en|^2 - 2 zn.en (quantize.py:22-27) cannot tell such scores apart either.  A row that pushed more
// than 32 codes (only codebooks with dozens of near-duplicates of a row's best code, or a zero latent) is re-done by a
// brute-force scan of the whole split (warp-cooperative), so the result never depends on the ring capacity.
//
// Structure as vq_exact4_kernel (TMA ring of code tiles, two MMA issuer threads, 8 drain warps), plus:
//   * every CTA starts its pass over the codebook at a different tile (see vq_tile_rot);
//   * the A tile is double-buffered so the next item's MMAs start while the rows of this item are resolved;
//   * accumulators are read 64 columns at a time, software-pipelined (tcgen05.ld of chunk g+1 in flight while chunk g is
//     reduced); the reduction is a 3-input-max tree (0.56 ALU operations per score) down to 8-code group maxima, and
//     only groups within reach call the (not inlined) push routine (the first version inlined ~500 instructions of push
//     code per chunk and thrashed the instruction cache: 365 us per call).
// ----------------------------------------------------------------------------------------------
#ifdef PM_VQ_DEBUG
// bring-up counters (scripts/vq_debug.py; built with PM_NVCC_EXTRA=-DPM_VQ_DEBUG): 0 events, 1 pushes, 2 flagged rows, 3 -,
// 4 resolve rounds, 5 scan cycles, 6 resolve cycles, 7 brute-force cycles, 8 issuer waits, 9 -, 10 drain t_full wait,
// 11 issuer MMA issue, 12 issuer commits
__device__ unsigned long long g_vq_dbg[16];
#define VQ_DBG_DECL unsigned long long dbg_loc[16] = {};
#define VQ_DBG_ADD(i, v) dbg_loc[i] += static_cast<unsigned long long>(v)
#define VQ_DBG_FLUSH                                   \
  for (int di = 0; di < 16; ++di)                      \
    if (dbg_loc[di] != 0ull) atomicAdd(&g_vq_dbg[di], dbg_loc[di]);
#else
#define VQ_DBG_DECL
#define VQ_DBG_ADD(i, v)
#define VQ_DBG_FLUSH
#endif
constexpr float VQ_DELTA = 0.0159f;           // >= 2 eps (see above)
constexpr float VQ_PACK_ERR = 0.0011f;        // a ring entry keeps 10 mantissa bits of a |score| <= 1.0001: error < 2^-10
constexpr float VQ_TIE = 2.4e-7f;             // scores this close are one tie class: the lowest index wins
constexpr int VQ_RING = 32;
constexpr uint32_t VQ_GID_MASK = 0x1FFFu;     // 13 bits of code index: n_e <= 8192
constexpr int VQ2_BSTAGES = 8;
constexpr int VQ2_B_BYTES = VQ_BN * 64;       // 8 KB: 128 codes x 32 bf16 (the high halves only), 64B-swizzled rows
constexpr int VQ2_SMEM_BYTES = 1024 + 4 * VQ_A_SLAB + VQ2_BSTAGES * VQ2_B_BYTES + VQ_RING * VQ_BM * 4 + 512;

// Tile order of an item: every CTA starts its pass over the codebook at a different tile and wraps around.  All 148 CTAs run
// in lock-step otherwise and ask the L2 for the same 128 lines at the same time (measured: the MMA issuers then wait
// for operands most of the time); the arg-max does not depend on the order (ties are resolved by index, not by arrival).
__device__ __forceinline__ int vq_tile_rot(int ntiles) {
  return static_cast<int>((static_cast<long long>(blockIdx.x) * ntiles) / gridDim.x);
}

// shared-memory accesses by 32-bit shared::cta address (pointers re-derived through integer arithmetic compile to generic LD / ST)
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

// Push the codes of one 8-code group that are within reach of the running maximum onto the row's ring (entry = 19 high
// bits of the score | 13 bits of code index).  Deliberately NOT inlined: the scan loop calls it from 16 places and is
// entered by few lanes at a time; inlined, the push code (16 x ~60 instructions) pushed the loop out of the instruction cache.
__device__ __noinline__ int vq_push8(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7, float thr,
                                     uint32_t code, uint32_t ring_addr, int n) {
#define VQ_PUSH1(V, U)                                                                                 \
  if ((V) >= thr) {                                                                                   \
    sts_u32(ring_addr + (static_cast<uint32_t>(n) & (VQ_RING - 1)) * (VQ_BM * 4),                      \
            (__float_as_uint(V) & ~VQ_GID_MASK) | (code + (U)));                                       \
    ++n;                                                                                              \
  }
  VQ_PUSH1(v0, 0u) VQ_PUSH1(v1, 1u) VQ_PUSH1(v2, 2u) VQ_PUSH1(v3, 3u)
  VQ_PUSH1(v4, 4u) VQ_PUSH1(v5, 5u) VQ_PUSH1(v6, 6u) VQ_PUSH1(v7, 7u)
#undef VQ_PUSH1
  return n;
}

__device__ __forceinline__ float dot32(const float (&a)[VQ_D], const float (&b)[VQ_D]) {
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
  for (int i = 0; i < VQ_D; i += 4) {
    s0 = fmaf(a[i], b[i], s0);
    s1 = fmaf(a[i + 1], b[i + 1], s1);
    s2 = fmaf(a[i + 2], b[i + 2], s2);
    s3 = fmaf(a[i + 3], b[i + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// running (score, index) under the tie rule: a larger score wins; scores within VQ_TIE of the maximum keep the lowest index
__device__ __forceinline__ void vq_take(float d, int i, float& bd, int& bi) {
  if (d > bd) {
    if (d - bd > VQ_TIE || i < bi) bi = i;
    bd = d;
  } else if (bd - d <= VQ_TIE && i < bi) {
    bi = i;
  }
}

// per-row finish with the normalised row already in registers
__device__ __forceinline__ float vq_finish_row_zn(const VqParams& p, int row, int idx, const float (&zn)[VQ_D]) {
  float e[VQ_D];
  load_row32(p.en + static_cast<size_t>(idx) * VQ_D, e);   // == l2norm(E[idx])
  float sse = 0.0f;
  float out[VQ_D];
#pragma unroll
  for (int i = 0; i < VQ_D; ++i) {
    const float d = e[i] - zn[i];
    sse = fmaf(d, d, sse);
    out[i] = zn[i] + d;                                     // z + (z_q - z).detach(), forward value
  }
  if (p.zq != nullptr) {
    float* dst = p.zq + static_cast<size_t>(row) * VQ_D;
#pragma unroll
    for (int i = 0; i < VQ_D; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(out[i], out[i + 1], out[i + 2], out[i + 3]);
  }
  if (p.zq_split != nullptr) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < VQ_D; i += 2) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(out[i]), h1 = __float2bfloat16_rn(out[i + 1]);
      hi[i >> 1] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
      lo[i >> 1] = pack_bf16x2(out[i] - __bfloat162float(h0), out[i + 1] - __bfloat162float(h1));
    }
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.zq_split) + static_cast<size_t>(row) * 64);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[4 + c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
  }
  if (p.idx != nullptr) p.idx[row] = static_cast<long long>(idx);
  if (p.hist != nullptr) atomicAdd(p.hist + idx, 1ull);
  return sse;
}

__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_main_kernel(const __grid_constant__ CUtensorMap tmB, const VqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                                 // [buffer 2][rowhalf 2][16 KB]  (z_hi in the first 64 B of each 128 B row)
  uint8_t* smB = smem + 4 * VQ_A_SLAB;                 // [VQ2_BSTAGES][8 KB]
  uint32_t* ring = reinterpret_cast<uint32_t*>(smB + VQ2_BSTAGES * VQ2_B_BYTES);   // [VQ_RING][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + VQ_RING * VQ_BM);
  uint64_t* b_full = bars;                             // [VQ2_BSTAGES]
  uint64_t* b_empty = bars + VQ2_BSTAGES;              // [VQ2_BSTAGES]
  uint64_t* t_full = bars + 2 * VQ2_BSTAGES;           // [2 stages][2 row halves]
  uint64_t* t_empty = t_full + 4;                      // [2 stages][2 row halves]
  uint64_t* a_full = t_empty + 4;                      // [1]  A tile written (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  VQ_DBG_DECL

  const int row_tiles = (p.M + VQ_BM - 1) / VQ_BM;
  const int items = row_tiles * p.splits;
  const int codes_per_split = p.n_e / p.splits;
  const int ntiles = (codes_per_split + VQ_BN - 1) / VQ_BN;
  const int rot = vq_tile_rot(ntiles);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < VQ2_BSTAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 2);       // one commit per MMA issuer (row half)
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);       // one arrival per drain warp of that row half
    }
    mbar_init(a_full, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int split = item % p.splits;
        const int code0 = split * codes_per_split;
        for (int t = 0; t < ntiles; ++t) {
          int tt = t + rot;
          if (tt >= ntiles) tt -= ntiles;
          mbar_wait(&b_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&b_full[st], VQ2_B_BYTES);
          tma_load_2d(smB + st * VQ2_B_BYTES, &tmB, &b_full[st], 0, code0 + tt * VQ_BN);
          if (++st == VQ2_BSTAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, VQ_BN, 0, 0);
      const int rh = (warp == 1) ? 0 : 1;
      int st = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      uint32_t item_ph = 0;
      int abuf = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        mbar_wait(a_full, item_ph);          // A tile of this item is in smem
        item_ph ^= 1;
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(smA + (abuf * 2 + rh) * VQ_A_SLAB));
        for (int t = 0; t < ntiles; ++t) {
#ifdef PM_VQ_DEBUG
          const long long w0 = clock64();
#endif
          // both barriers polled together: a try_wait costs ~90 cycles even when the phase has already completed
          {
            uint64_t* const be = &t_empty[as * 2 + rh];
            uint64_t* const bf = &b_full[st];
            uint32_t ok_e = mbar_try_wait(be, aph ^ 1);
            uint32_t ok_f = mbar_try_wait(bf, ph);
            if (!(ok_e && ok_f)) {
              if (!ok_e) mbar_wait(be, aph ^ 1);
              if (!ok_f) mbar_wait(bf, ph);
            }
          }
#ifdef PM_VQ_DEBUG
          const long long w2 = clock64();
          VQ_DBG_ADD(8, w2 - w0);
#endif
          tc_fence_after();
          const uint64_t db = umma_desc_sw64(smem_u32(smB + st * VQ2_B_BYTES));
          const uint32_t tacc = tmem_base + (as * 2 + rh) * VQ_BN;
          // z_hi . e_hi: K = 32 = two k-steps of 16 (32 B each)
          umma_ss(tacc, da, db, idesc, 0u);
          umma_ss(tacc, da + 2, db + 2, idesc, 1u);
#ifdef PM_VQ_DEBUG
          const long long w3 = clock64();
#endif
          umma_commit(&b_empty[st]);
          umma_commit(&t_full[as * 2 + rh]);
#ifdef PM_VQ_DEBUG
          VQ_DBG_ADD(11, w3 - w2);
          VQ_DBG_ADD(12, clock64() - w3);
#endif
          if (++st == VQ2_BSTAGES) { st = 0; ph ^= 1; }
          if (++as == 2) { as = 0; aph ^= 1; }
        }
        abuf ^= 1;
      }
    }
  } else if (warp >= 2 && warp < 10) {
    const int q = warp & 3;
    const int rh = (warp - 2) >> 2;                    // row half 0/1
    const int r_in_half = q * 32 + lane;               // 0..127
    const int r_in_tile = rh * 128 + r_in_half;        // 0..255
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t my_ring = smem_u32(ring + r_in_tile);      // slot s at my_ring + s * 1024 (bytes)
    int as = 0;
    uint32_t aph = 0;
    int abuf = 0;
    double sse_acc = 0.0;

    // A tile of an item: normalise the row, keep the bf16 high halves, 128B-swizzled K-major rows
    auto write_a = [&](int item, int buf) {
      const int rt = item / p.splits;
      const int row = rt * VQ_BM + r_in_tile;
      float x[VQ_D];
      if (row < p.M) {
        load_row32(p.z + static_cast<size_t>(row) * p.ldz, x);
        l2norm32(x);
      } else {
#pragma unroll
        for (int i = 0; i < VQ_D; ++i) x[i] = 0.0f;
      }
      uint32_t hi[16];
#pragma unroll
      for (int i = 0; i < VQ_D; i += 2) hi[i >> 1] = pack_bf16x2(x[i], x[i + 1]);
      uint8_t* a_hi = smA + (buf * 2 + rh) * VQ_A_SLAB + r_in_half * 128;
      const int sw = r_in_half & 7;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(a_hi + ((c ^ sw) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
    };

    if (static_cast<int>(blockIdx.x) < items) write_a(blockIdx.x, 0);

    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int rt = item / p.splits, split = item % p.splits;
      const int row = rt * VQ_BM + r_in_tile;
      const bool row_ok = row < p.M;
      const int code0 = split * codes_per_split;
      const int lim = min(code0 + codes_per_split, p.n_e);

      // ---- scan: running approximate maximum + ring of near-maximum codes ----
      float best = -INFINITY;
      int n = 0;                       // codes pushed so far
      uint32_t r0[64], r1[64];

      // 64 scores (codes cb .. cb+63) -> 8 group maxima -> event test -> pushes
      auto reduce_chunk = [&](uint32_t (&v)[64], int cb, bool tail) {
        if (tail) {
          // codes beyond this split / the codebook (TMA zero-filled rows) must never win
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (cb + i >= lim) v[i] = 0xff800000u;   // -inf
        }
        float g8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float a = fmaxf(__uint_as_float(v[8 * k]), fmaxf(__uint_as_float(v[8 * k + 1]), __uint_as_float(v[8 * k + 2])));
          const float b = fmaxf(__uint_as_float(v[8 * k + 3]), fmaxf(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5])));
          g8[k] = fmaxf(fmaxf(a, b), fmaxf(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7])));
        }
        const float cm = fmaxf(fmaxf(fmaxf(g8[0], g8[1]), fmaxf(g8[2], g8[3])), fmaxf(fmaxf(g8[4], g8[5]), fmaxf(g8[6], g8[7])));
#if defined(PM_VQ_EXP) && (PM_VQ_EXP == 1 || PM_VQ_EXP == 6)
        best = fmaxf(best, cm);
        const bool ev = false;
#else
        const bool ev = row_ok && (cm >= best - VQ_DELTA) && (cm > -INFINITY);
#endif
        if (__any_sync(0xffffffffu, ev)) {
          if (ev) {
            best = fmaxf(best, cm);
            VQ_DBG_ADD(0, 1);
          }
          const float thr = ev ? best - VQ_DELTA : INFINITY;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (__any_sync(0xffffffffu, g8[k] >= thr))
              n = vq_push8(__uint_as_float(v[8 * k]), __uint_as_float(v[8 * k + 1]), __uint_as_float(v[8 * k + 2]),
                           __uint_as_float(v[8 * k + 3]), __uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5]),
                           __uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7]), thr, static_cast<uint32_t>(cb + 8 * k),
                           my_ring, n);
          }
        }
      };

#ifdef PM_VQ_DEBUG
      const long long dbg_t0 = clock64();
#endif
      // software pipeline over the flattened chunk stream: the tcgen05.ld of chunk g+1 is in flight while chunk g is reduced
      mbar_wait(&t_full[as * 2 + rh], aph);
      tc_fence_after();
      tmem_ld_x64(tmem_base + lane_off + (as * 2 + rh) * VQ_BN, r0);
      tmem_ld_wait();
      reg_fence64(r0);
      for (int t = 0; t < ntiles; ++t) {
        const uint32_t tacc = tmem_base + lane_off + (as * 2 + rh) * VQ_BN;
        int tt = t + rot;
        if (tt >= ntiles) tt -= ntiles;
        const int cbase = code0 + tt * VQ_BN;
        const bool tail = cbase + VQ_BN > lim;
        const int as_next = as ^ 1;
        const uint32_t aph_next = (as == 1) ? (aph ^ 1) : aph;
        // chunk 0 (in r0) | load chunk 1
#if defined(PM_VQ_EXP) && PM_VQ_EXP == 6
        // (bring-up experiment: half of the TMEM reads)
#else
        tmem_ld_x64(tacc + 64, r1);
#endif
#ifdef PM_VQ_DEBUG
        const long long ta = clock64();
#endif
        reduce_chunk(r0, cbase, tail);
#ifdef PM_VQ_DEBUG
        const long long tb = clock64();
#endif
        tmem_ld_wait();
        reg_fence64(r1);
#ifdef PM_VQ_DEBUG
        const long long tc = clock64();
        VQ_DBG_ADD(13, tb - ta);
        VQ_DBG_ADD(14, tc - tb);
#endif
        // last TMEM read of this accumulator stage is done
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[as * 2 + rh]);
        // chunk 1 (in r1) | load chunk 0 of the next tile
        if (t + 1 < ntiles) {
#ifdef PM_VQ_DEBUG
          const long long d0 = clock64();
#endif
          mbar_wait(&t_full[as_next * 2 + rh], aph_next);
#ifdef PM_VQ_DEBUG
          VQ_DBG_ADD(10, clock64() - d0);
#endif
          tc_fence_after();
          tmem_ld_x64(tmem_base + lane_off + (as_next * 2 + rh) * VQ_BN, r0);
        }
        reduce_chunk(r1, cbase + 64, tail);
        tmem_ld_wait();
        reg_fence64(r0);
        as = as_next;
        aph = aph_next;
      }
#ifdef PM_VQ_DEBUG
      const long long dbg_t1 = clock64();
      VQ_DBG_ADD(5, dbg_t1 - dbg_t0);
      VQ_DBG_ADD(1, n);
#endif

      // ---- the next item's A tile first: its MMAs run while this item's rows are resolved ----
      const int next_item = item + gridDim.x;
      if (next_item < items) write_a(next_item, abuf ^ 1);
      abuf ^= 1;

      // ---- resolve: re-score the ring entries that can still be the arg-max ----
      float zn[VQ_D];
      if (row_ok) {
        load_row32(p.z + static_cast<size_t>(row) * p.ldz, zn);
        l2norm32(zn);
      } else {
#pragma unroll
        for (int i = 0; i < VQ_D; ++i) zn[i] = 0.0f;
      }
      const float thr = best - (VQ_DELTA + VQ_PACK_ERR);
      const int cnt = n < VQ_RING ? n : VQ_RING;
      const bool flagged = row_ok && n > VQ_RING;      // the ring wrapped: an entry that matters may be gone
#ifdef PM_VQ_DEBUG
      if (flagged) VQ_DBG_ADD(2, 1);
      const long long dbg_t2 = clock64();
#endif
      float bestd = -INFINITY;
      int bidx = 0x7fffffff;
      int pos = (!row_ok || flagged) ? cnt : 0;
#if defined(PM_VQ_EXP) && PM_VQ_EXP == 2
      pos = cnt;
#endif
      while (true) {
        uint32_t ent = 0;
        bool take = false;
        while (pos < cnt) {          // next ring entry that can still hold the arg-max
          const uint32_t e = lds_u32(my_ring + pos * (VQ_BM * 4));
          ++pos;
          if (__uint_as_float(e & ~VQ_GID_MASK) >= thr) {
            ent = e;
            take = true;
            break;
          }
        }
        if (!__any_sync(0xffffffffu, take)) break;
        VQ_DBG_ADD(4, 1);
        if (take) {
          const int ci = static_cast<int>(ent & VQ_GID_MASK);
          float e[VQ_D];
          load_row32(p.en + static_cast<size_t>(ci) * VQ_D, e);
          vq_take(dot32(zn, e), ci, bestd, bidx);
        }
      }
#ifdef PM_VQ_DEBUG
      const long long dbg_t3 = clock64();
      VQ_DBG_ADD(6, dbg_t3 - dbg_t2);
#endif
      // rows whose ring wrapped: brute-force scan of the split, one row at a time by the whole warp
      // (lane l takes codes code0 + l, code0 + l + 32, ...)
      unsigned fl = __ballot_sync(0xffffffffu, flagged);
      while (fl != 0u) {
        const int src = __ffs(fl) - 1;
        fl &= fl - 1u;
        float zs[VQ_D];
#pragma unroll
        for (int k = 0; k < VQ_D; ++k) zs[k] = __shfl_sync(0xffffffffu, zn[k], src);
        float bd = -INFINITY;
        int bi = 0x7fffffff;
        for (int j = code0 + lane; j < lim; j += 32) {
          float e[VQ_D];
          load_row32(p.en + static_cast<size_t>(j) * VQ_D, e);
          vq_take(dot32(zs, e), j, bd, bi);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, bd, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          vq_take(od, oi, bd, bi);
        }
        if (lane == src) { bestd = bd; bidx = bi; }
      }
#ifdef PM_VQ_DEBUG
      VQ_DBG_ADD(7, clock64() - dbg_t3);
#endif
#if defined(PM_VQ_EXP) && (PM_VQ_EXP == 1 || PM_VQ_EXP == 2 || PM_VQ_EXP == 6)
      if (bidx == 0x7fffffff) bidx = row % p.n_e;
#endif
      if (bidx == 0x7fffffff) bidx = code0;     // no comparable score at all (NaN latents): first code, never out of bounds

      if (p.splits == 1) {
        float sse = 0.0f;
        if (row_ok) sse = vq_finish_row_zn(p, row, bidx, zn);
        sse_acc += static_cast<double>(warp_sum(sse));
      } else if (row_ok) {
        p.cand_val[static_cast<size_t>(split) * p.M + row] = bestd;
        p.cand_idx[static_cast<size_t>(split) * p.M + row] = bidx;
      }
    }
    if (p.splits == 1 && lane == 0 && p.sse != nullptr && sse_acc != 0.0) atomicAdd(p.sse, sse_acc);
  }

  VQ_DBG_FLUSH
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// merge the per-split candidates (max value, ties -> lower index) and finish the rows
__global__ void vq_finalize_kernel(const VqParams p) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  float sse = 0.0f;
  if (row < p.M) {
    float best = p.cand_val[row];
    int bidx = p.cand_idx[row];
    for (int s = 1; s < p.splits; ++s) {
      const float v = p.cand_val[static_cast<size_t>(s) * p.M + row];
      const int i = p.cand_idx[static_cast<size_t>(s) * p.M + row];
      if (v > best) {
        if (v - best > VQ_TIE || i < bidx) bidx = i;
        best = v;
      } else if (best - v <= VQ_TIE && i < bidx) {
        bidx = i;
      }
    }
    sse = vq_finish_row(p, row, bidx);
  }
  sse = warp_sum(sse);
  if ((threadIdx.x & 31) == 0 && p.sse != nullptr && sse != 0.0f) atomicAdd(p.sse, static_cast<double>(sse));
}

// decode_from_indice (quantize.py:40-44): z_q = l2norm(E[idx]); also the bf16 hi/lo split
__global__ void vq_gather_kernel(const long long* __restrict__ idx, int M, int n_rows_table,
                                 const float* __restrict__ table, int normalize,
                                 float* __restrict__ out, __nv_bfloat16* __restrict__ out_split) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  long long i = idx[row];
  if (i < 0) i = 0;
  if (i >= n_rows_table) i = n_rows_table - 1;
  float x[VQ_D];
  load_row32(table + static_cast<size_t>(i) * VQ_D, x);
  if (normalize) l2norm32(x);
  if (out != nullptr) {
#pragma unroll
    for (int k = 0; k < VQ_D; k += 4)
      *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * VQ_D + k) = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
  }
  if (out_split != nullptr) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int k = 0; k < VQ_D; k += 2) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x[k]), h1 = __float2bfloat16_rn(x[k + 1]);
      hi[k >> 1] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
      lo[k >> 1] = pack_bf16x2(x[k] - __bfloat162float(h0), x[k + 1] - __bfloat162float(h1));
    }
    uint4* dst = reinterpret_cast<uint4*>(out_split + static_cast<size_t>(row) * 64);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[4 + c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
  }
}

// fp32 [M, 32] -> bf16 [M, 64] = [hi | lo]  (decoder input when the caller passes its own z)
__global__ void split_rows32_kernel(const float* __restrict__ src, int64_t ld, int M, __nv_bfloat16* __restrict__ out_split) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float x[VQ_D];
  load_row32(src + static_cast<size_t>(row) * ld, x);
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int k = 0; k < VQ_D; k += 2) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[k]), h1 = __float2bfloat16_rn(x[k + 1]);
    hi[k >> 1] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
    lo[k >> 1] = pack_bf16x2(x[k] - __bfloat162float(h0), x[k + 1] - __bfloat162float(h1));
  }
  uint4* dst = reinterpret_cast<uint4*>(out_split + static_cast<size_t>(row) * 64);
#pragma unroll
  for (int c = 0; c < 4; ++c) dst[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
#pragma unroll
  for (int c = 0; c < 4; ++c) dst[4 + c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
}

// ----------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------
int pm_vq_codebook_prep_launch(const float* E, int n_e, float* en, void* packed, cudaStream_t stream) {
  if (E == nullptr || en == nullptr || packed == nullptr || n_e <= 0) return PM_ERR_INVALID;
  vq_codebook_prep_kernel<<<(n_e + 127) / 128, 128, 0, stream>>>(E, n_e, en, reinterpret_cast<__nv_bfloat16*>(packed));
  return static_cast<int>(cudaGetLastError());
}

int pm_vq_launch(const VqParams& p_in, cudaStream_t stream) {
  VqParams p = p_in;
  if (p.z == nullptr || p.en == nullptr || p.packed == nullptr || p.M <= 0 || p.n_e <= 0) return PM_ERR_INVALID;
  if (p.e_dim != VQ_D || (p.ldz % 4) != 0) return PM_ERR_INVALID;
  if (p.splits <= 0) {
    // split the codebook over several CTAs per row tile only when the row tiles alone cannot fill the machine
    // (the merge costs a second kernel and every split re-normalises its rows): measured at M = 65536,
    // 256 row tiles: 117 us unsplit vs 138 us with 4 splits
    const int row_tiles = (p.M + VQ_BM - 1) / VQ_BM;
    int s = 1;
    while (row_tiles * s < pm_num_sms() && (p.n_e / (s * 2)) >= 8 * VQ_BN && (p.n_e % (s * 2 * VQ_BN)) == 0 && s < 8) s *= 2;
    p.splits = s;
  }
  if (p.n_e % p.splits != 0) return PM_ERR_INVALID;
  if (p.splits > 1 && ((p.n_e / p.splits) % VQ_BN != 0 || p.cand_val == nullptr || p.cand_idx == nullptr)) return PM_ERR_INVALID;
  int rc;
  static bool attr_done[PM_MAX_DEVICES] = {}, attr_done4[PM_MAX_DEVICES] = {};
  // PM_VQ_MODE=1: the 1x-MMA kernel with exact re-scoring (vq_main_kernel); default (4): the four-term kernel.  Measured on
  // BASELINE configs[1] (profiles/r02_vq.md): four-term 116 us, 1x-MMA 133-162 us — its drain (one fp32 max per score on the
  // half-rate ALU pipe, ~400 cycles per 128-code tile for 8 warps, plus the push path) is slower than the 2 MMAs it waits for.
  static int env_mode = 0;
  if (env_mode == 0) {
    const char* env = getenv("PM_VQ_MODE");
    env_mode = (env != nullptr && env[0] == '1') ? 1 : 4;
  }
  // the 1x kernel packs code indices into 13 bits of its ring entries: larger codebooks take the four-term kernel
  const int mode = (p.n_e > static_cast<int>(VQ_GID_MASK + 1u)) ? 4 : env_mode;
  const int items = ((p.M + VQ_BM - 1) / VQ_BM) * p.splits;
  const int grid = items < pm_num_sms() ? items : pm_num_sms();
  if (mode == 4) {
    CUtensorMap tmB;
    rc = pm_make_tmap_2d(&tmB, p.packed, 2, p.n_e, 64, 64, VQ_BN, 64);
    if (rc != PM_OK) return rc;
    if (const int rc2 = pm_ensure_dyn_smem(vq_exact4_kernel, VQ_SMEM_BYTES, attr_done4)) return rc2;
    vq_exact4_kernel<<<grid, VQ_THREADS, VQ_SMEM_BYTES, stream>>>(tmB, p);
  } else {
    // the high halves only: a [128 codes, 32 bf16] box out of the [n_e, 64] packed rows (64 of every 128 bytes are read)
    CUtensorMap tmBh;
    rc = pm_make_tmap_2d_sw64(&tmBh, p.packed, 2, p.n_e, 32, 64, VQ_BN, 32);
    if (rc != PM_OK) return rc;
    if (const int rc2 = pm_ensure_dyn_smem(vq_main_kernel, VQ2_SMEM_BYTES, attr_done)) return rc2;
    vq_main_kernel<<<grid, VQ_THREADS, VQ2_SMEM_BYTES, stream>>>(tmBh, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (p.splits > 1) {
    vq_finalize_kernel<<<(p.M + 127) / 128, 128, 0, stream>>>(p);
    e = cudaGetLastError();
  }
  return static_cast<int>(e);
}

#ifdef PM_VQ_DEBUG
extern "C" int pm_debug_vq_counters(unsigned long long* out16, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_vq_dbg, sizeof(g_vq_dbg));
  if (e == cudaSuccess && reset) {
    unsigned long long z[16] = {};
    e = cudaMemcpyToSymbol(g_vq_dbg, z, sizeof(z));
  }
  return static_cast<int>(e);
}
#endif

int pm_vq_gather_launch(const long long* idx, int M, int n_rows, const float* table, int normalize,
                        float* out, void* out_split, cudaStream_t stream) {
  if (idx == nullptr || table == nullptr || M <= 0 || n_rows <= 0) return PM_ERR_INVALID;
  vq_gather_kernel<<<(M + 127) / 128, 128, 0, stream>>>(idx, M, n_rows, table, normalize, out,
                                                          reinterpret_cast<__nv_bfloat16*>(out_split));
  return static_cast<int>(cudaGetLastError());
}

int pm_split_rows32_launch(const float* src, int64_t ld, int M, void* out_split, cudaStream_t stream) {
  if (src == nullptr || out_split == nullptr || M <= 0 || (ld % 4) != 0) return PM_ERR_INVALID;
  split_rows32_kernel<<<(M + 127) / 128, 128, 0, stream>>>(src, ld, M, reinterpret_cast<__nv_bfloat16*>(out_split));
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
