// pm_maskgit.cu — the HBM-bound tail of a MaskGIT step (sm_100a), one pass over the logits.
//
// Replaces, for one call of Pipeline.sample (reference generate.py:159-181):
//   top_k (generate.py:33-37)            keep the k largest logits per token, -inf elsewhere
//   gumbel_sample (generate.py:40-46)    argmax(filtered / max(T, 1e-10) - log(-log(u)))
//   fill + confidence (generate.py:166-173)  ids = where(ids == mask, pred, ids);
//                                        score = 1 - softmax(logits)[pred]; unmasked -> -1e5
//   re-mask (generate.py:175-179)        ids.scatter(topk(scores, k).indices, mask_id)
// The reference reads the [B, N, 8192] fp32 logits >= 4 times and draws a same-sized uniform tensor;
// here one warp streams each 32 KB row once (online max / sum-exp + per-lane top-k lists merged with
// warp shuffles) and draws only the k uniforms it needs (Philox4x32-10 keyed on (seed, row, index)),
// or reads them from an injected noise tensor for parity tests.
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <stdlib.h>

namespace pm {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (index, row, offset, 0), key = seed
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t row, uint32_t index, uint32_t offset) {
  uint32_t c0 = index, c1 = row, c2 = offset, c3 = 0x9e3779b9u;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return static_cast<float>(c0 >> 8) * (1.0f / 16777216.0f);   // [0, 1) with 24 random bits, like torch's uniform_
}

// Warp-wide max of a float through ONE redux.sync on an order-preserving integer image of the value (instead of five
// dependent shuffle + max levels): the selection loops below are latency chains, not throughput work.
__device__ __forceinline__ int f32_ordered(float x) {
  const int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float warp_max_f32(float x) {
  const int m = __reduce_max_sync(0xffffffffu, f32_ordered(x));
  return __int_as_float(m ^ ((m >> 31) & 0x7fffffff));
}

template <int KMAX>
struct TopList {
  float v[KMAX];
  int i[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int t = 0; t < KMAX; ++t) { v[t] = -INFINITY; i[t] = 0x7fffffff; }
  }
  // insert (x, xi) into the descending list; ties keep the lower index first
  __device__ __forceinline__ void insert(float x, int xi) {
#pragma unroll
    for (int t = 0; t < KMAX; ++t) {
      const bool gt = (x > v[t]) || (x == v[t] && xi < i[t]);
      const float tv = v[t]; const int ti = i[t];
      v[t] = gt ? x : tv;  i[t] = gt ? xi : ti;
      x = gt ? tv : x;     xi = gt ? ti : xi;
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int t = 0; t + 1 < KMAX; ++t) { v[t] = v[t + 1]; i[t] = i[t + 1]; }
    v[KMAX - 1] = -INFINITY; i[KMAX - 1] = 0x7fffffff;
  }
};

template <int KMAX>
__global__ void __launch_bounds__(256)
maskgit_sample_kernel(const MaskgitParams p_in) {
  // per-step scalars from device memory when the step runs inside a CUDA graph (see pm_step_scalars in the C-ABI header)
  MaskgitParams p = p_in;
  if (p_in.step_tab != nullptr) {
    const StepScalars sc = p_in.step_tab[*p_in.step_idx];
    p.temperature = sc.temperature;
    p.seed = sc.seed;
    p.offset = sc.offset;
  }
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= p.M) return;
  const float* x = p.logits + static_cast<size_t>(row) * p.ld;
  const int k = p.topk;

  TopList<KMAX> top;
  top.init();
  float m = -INFINITY, s = 0.0f;      // lane-local online softmax state
  const int nvec = p.V >> 2;
  for (int c0 = lane; c0 < nvec; c0 += 32 * 4) {
    // 4 independent 16-byte loads in flight per lane
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 32;
      q[u] = (c < nvec) ? __ldcs(reinterpret_cast<const float4*>(x) + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    float cm = m;
#pragma unroll
    for (int u = 0; u < 4; ++u) cm = fmaxf(cm, fmaxf(fmaxf(q[u].x, q[u].y), fmaxf(q[u].z, q[u].w)));
    if (cm > m) { s *= __expf(m - cm); m = cm; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int base = (c0 + u * 32) * 4;
      const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        s += __expf(e[t] - m);
        if (e[t] > top.v[KMAX - 1] && k > 0) top.insert(e[t], base + t);   // rare after the first few chunks
      }
    }
  }
  // warp-wide softmax statistics
  float mw = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  float sw = (m == -INFINITY) ? 0.0f : s * __expf(m - mw);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, o);

  // merge the per-lane lists: k rounds of warp arg-max (ties -> lower index); lane r keeps winner r
  float my_v = -INFINITY;
  int my_i = 0x7fffffff;
  for (int r = 0; r < k; ++r) {
    float bv = top.v[0];
    int bi = top.i[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (top.i[0] == bi && top.v[0] == bv) top.pop();
    if (lane == r) { my_v = bv; my_i = bi; }
  }

  // gumbel-perturbed arg-max over the k survivors (generate.py:45-46)
  float score = -INFINITY;
  if (lane < k && my_i != 0x7fffffff) {
    float u;
    if (p.noise != nullptr) u = p.noise[static_cast<size_t>(row) * p.ld_noise + my_i];
    else u = philox_uniform(p.seed, static_cast<uint32_t>(row), static_cast<uint32_t>(my_i), static_cast<uint32_t>(p.offset));
    const float inner = -logf(fmaxf(u, 1e-20f));
    const float g = -logf(fmaxf(inner, 1e-20f));
    score = my_v / fmaxf(p.temperature, 1e-10f) + g;
  }
  float bs = score;
  int bi = (lane < k) ? my_i : 0x7fffffff;
  float bl = my_v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const float ol = __shfl_xor_sync(0xffffffffu, bl, o);
    if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; bl = ol; }
  }
  if (lane == 0) {
    const long long pred = static_cast<long long>(bi);
    const float prob = __expf(bl - mw) / sw;                 // softmax(logits)[pred], unfiltered, T = 1
    if (p.pred_ids != nullptr) p.pred_ids[row] = pred;
    bool is_mask = true;
    if (p.ids != nullptr) {
      is_mask = (p.ids[row] == p.mask_id);
      if (is_mask) p.ids[row] = pred;                        // fill the mask, keep the unmasked
    }
    if (p.scores != nullptr) p.scores[row] = is_mask ? (1.0f - prob) : -1e5f;
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory staged variant (V % 128 == 0, row <= 64 KB): each warp pulls its whole logit row into shared
// memory with ONE 1-D bulk TMA copy (cp.async.bulk + mbarrier), then makes two cheap passes over it:
//   pass 1: row max and, per lane, the two largest values (branch-free max/min network) -> tau = k-th largest of
//           the 64 lane leaders, a lower bound of the row's k-th largest value
//   pass 2: sum of exp(x - max) and the few candidates x >= tau (warp-aggregated append), from which the exact
//           top-k (ties -> lower index) is selected
// Streaming insertion sort (the kernel above) diverges on almost every element; this one is HBM-bound.
// ---------------------------------------------------------------------------------------------
constexpr int MG_CAND = 128;     // candidate capacity per row (expected: k .. ~2k entries)

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
maskgit_sample_smem_kernel(const MaskgitParams p_in) {
  // per-step scalars from device memory when the step runs inside a CUDA graph (see pm_step_scalars in the C-ABI header)
  MaskgitParams p = p_in;
  if (p_in.step_tab != nullptr) {
    const StepScalars sc = p_in.step_tab[*p_in.step_idx];
    p.temperature = sc.temperature;
    p.seed = sc.seed;
    p.offset = sc.offset;
  }
  extern __shared__ __align__(128) uint8_t mg_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_bytes = p.V * 4;
  float* buf = reinterpret_cast<float*>(mg_smem + static_cast<size_t>(warp) * row_bytes);
  uint8_t* tail = mg_smem + static_cast<size_t>(WARPS) * row_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail) + warp;
  float* cand_v = reinterpret_cast<float*>(tail + 64) + warp * MG_CAND;
  int* cand_i = reinterpret_cast<int*>(tail + 64 + WARPS * MG_CAND * 4) + warp * MG_CAND;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int k = p.topk;
  const int nvec = p.V >> 2;
  uint32_t phase = 0;
  const int row_stride = gridDim.x * WARPS;
  for (int row = blockIdx.x * WARPS + warp; row < p.M; row += row_stride) {
    // ---- one bulk copy of the row: global -> shared ----
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, static_cast<uint32_t>(row_bytes));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(p.logits + static_cast<size_t>(row) * p.ld)),
                   "r"(static_cast<uint32_t>(row_bytes)), "r"(smem_u32(bar))
                   : "memory");
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    const float4* b4 = reinterpret_cast<const float4*>(buf);
    // ---- pass 1: lane-local two largest values (4 independent max/min chains for ILP: few warps per SM) ----
    float t0 = -INFINITY, t1 = -INFINITY;
    {
      float u0[4], u1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { u0[u] = -INFINITY; u1[u] = -INFINITY; }
      for (int c = lane; c < nvec; c += 128) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = (c + 32 * u < nvec) ? b4[c + 32 * u] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            u1[u] = fmaxf(u1[u], fminf(u0[u], e[t]));
            u0[u] = fmaxf(u0[u], e[t]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        t1 = fmaxf(t1, fminf(t0, u0[u]));
        t0 = fmaxf(t0, u0[u]);
        t1 = fmaxf(t1, fminf(t0, u1[u]));       // u1[u] <= u0[u] <= t0: only t1 can change
      }
    }
    float mw = t0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    // tau = k-th largest of the 64 lane leaders (k <= 32): k rounds of warp max with removal
    float a0 = t0, a1 = t1, tau = -INFINITY;
    for (int r = 0; r < k; ++r) {
      float bv = a0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bv = fmaxf(bv, __shfl_xor_sync(0xffffffffu, bv, o));
      tau = bv;
      const unsigned who = __ballot_sync(0xffffffffu, a0 == bv);
      if (lane == __ffs(who) - 1) { a0 = a1; a1 = -INFINITY; }      // pop from exactly one lane
    }
    // ---- pass 2: softmax denominator (4 independent partial sums) and candidates >= tau ----
    float ssum = 0.0f;
    int ncand = 0;                                                    // warp-uniform
    {
      float ps[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const float nm = -mw * 1.4426950408889634f;
      for (int c = lane; c < nvec; c += 128) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = (c + 32 * u < nvec) ? b4[c + 32 * u] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        bool any = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float ex;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(e[t], 1.4426950408889634f, nm)));
            ps[u] += ex;
            any |= (e[t] >= tau);
          }
        }
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const bool hit = e[t] >= tau;
              const unsigned m = __ballot_sync(0xffffffffu, hit);
              if (hit) {
                const int pos = ncand + __popc(m & ((1u << lane) - 1u));
                if (pos < MG_CAND) { cand_v[pos] = e[t]; cand_i[pos] = (c + 32 * u) * 4 + t; }
              }
              ncand += __popc(m);
            }
          }
        }
      }
      ssum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    __syncwarp();
    ncand = min(ncand, MG_CAND);
    // ---- exact top-k among the candidates (value desc, index asc); lane r keeps winner r ----
    float cv[MG_CAND / 32];
    int ci[MG_CAND / 32];
#pragma unroll
    for (int u = 0; u < MG_CAND / 32; ++u) {
      const int j = u * 32 + lane;
      cv[u] = j < ncand ? cand_v[j] : -INFINITY;
      ci[u] = j < ncand ? cand_i[j] : 0x7fffffff;
    }
    float my_v = -INFINITY;
    int my_i = 0x7fffffff;
    for (int r = 0; r < k; ++r) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int u = 0; u < MG_CAND / 32; ++u)
        if (cv[u] > bv || (cv[u] == bv && ci[u] < bi)) { bv = cv[u]; bi = ci[u]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
#pragma unroll
      for (int u = 0; u < MG_CAND / 32; ++u)
        if (ci[u] == bi) { cv[u] = -INFINITY; ci[u] = 0x7fffffff; }   // remove the winner (indices are unique)
      if (lane == r) { my_v = bv; my_i = bi; }
    }
    // ---- gumbel-perturbed arg-max over the k survivors (generate.py:45-46) ----
    float score = -INFINITY;
    if (lane < k && my_i != 0x7fffffff) {
      float u;
      if (p.noise != nullptr) u = p.noise[static_cast<size_t>(row) * p.ld_noise + my_i];
      else u = philox_uniform(p.seed, static_cast<uint32_t>(row), static_cast<uint32_t>(my_i), static_cast<uint32_t>(p.offset));
      const float inner = -logf(fmaxf(u, 1e-20f));
      const float g = -logf(fmaxf(inner, 1e-20f));
      score = my_v / fmaxf(p.temperature, 1e-10f) + g;
    }
    float bs = score;
    int bi = (lane < k) ? my_i : 0x7fffffff;
    float bl = my_v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const float ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; bl = ol; }
    }
    if (lane == 0) {
      const long long pred = static_cast<long long>(bi);
      const float prob = __expf(bl - mw) / ssum;                 // softmax(logits)[pred], unfiltered, T = 1
      if (p.pred_ids != nullptr) p.pred_ids[row] = pred;
      bool is_mask = true;
      if (p.ids != nullptr) {
        is_mask = (p.ids[row] == p.mask_id);
        if (is_mask) p.ids[row] = pred;                          // fill the mask, keep the unmasked
      }
      if (p.scores != nullptr) p.scores[row] = is_mask ? (1.0f - prob) : -1e5f;
    }
    __syncwarp();      // all lanes done with buf / candidates before the next bulk copy overwrites them
  }
}

// ---------------------------------------------------------------------------------------------
// Block-per-row variant of the staged kernel (V % 512 == 0, row <= 64 KB): FOUR warps share one row buffer, so an SM
// holds 6 rows in flight with 24 warps instead of 6 — the one-warp-per-row kernel above is latency-bound on its two
// passes (3.0 TB/s = 47 % of the HBM copy peak at V = 8192), not on HBM.  Per row:
//   TMA bulk copy -> pass 1 (per-thread two largest, per-warp k-th largest leader; tau = the best of the four
//   lower bounds) -> pass 2 (sum exp, candidates >= tau into per-warp lists, warp-aggregated appends in a fixed
//   order) -> warp 0 selects the exact top-k (value desc, index asc), applies the gumbel arg-max and writes.
// If a candidate list overflows (rows with thousands of equal values) the block falls back to k rounds of a
// block-wide arg-max over the staged row: results never depend on the capacity.
// Measured on [65536, 8192] (scripts/sample_variants.py, L2 flushed): this kernel 0.539 ms (4.0 TB/s; 0.598 ms before the
// warp maxima went through redux.sync), warp-per-row
// staged 0.707 ms, streaming insertion 1.80 ms; a register-resident variant (256 threads holding the row in
// registers, next row prefetched under the selection) was tried and measured SLOWER (0.695 ms) — the per-row critical
// path (shuffle-serial tau / top-k selection), not the staging, is what limits these kernels.
// Round 2: in-kernel phase timers (cycles per row and block, six blocks per SM): load 2,000 | pass 1 + tau 1,800 | pass 2 4,100 |
// selection 4,300.  Moving the selection to a FIFTH warp with double-buffered candidate lists (named-barrier hand-over, ids[row]
// fetched ahead, workers already staging the next row; bit-identical results, tests green) measured 0.641 ms against 0.540 ms
// for this kernel at six blocks per SM (0.727 ms at five): the phases stretch when more warps run — dropped (profiles/r02_membound.txt).
// Eight warps per row buffer (half the per-thread work of both passes): 0.85 ms — 67 registers x 256 threads leave three resident
// blocks instead of six; a single polling warp per block (the others asleep in the block barrier) changed neither variant.
// ---------------------------------------------------------------------------------------------
constexpr int MGB_WARPS = 4;
constexpr int MGB_CAND = 64;      // per warp

__global__ void __launch_bounds__(MGB_WARPS * 32)
maskgit_sample_block_kernel(const MaskgitParams p_in) {
  // per-step scalars from device memory when the step runs inside a CUDA graph (see pm_step_scalars in the C-ABI header)
  MaskgitParams p = p_in;
  if (p_in.step_tab != nullptr) {
    const StepScalars sc = p_in.step_tab[*p_in.step_idx];
    p.temperature = sc.temperature;
    p.seed = sc.seed;
    p.offset = sc.offset;
  }
  extern __shared__ __align__(128) uint8_t mg_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_bytes = p.V * 4;
  float* buf = reinterpret_cast<float*>(mg_smem);
  uint8_t* tail = mg_smem + row_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  float* sh_f = reinterpret_cast<float*>(tail + 16);                 // [0..3] warp max, [4..7] warp tau, [8..11] warp sum
  int* sh_n = reinterpret_cast<int*>(tail + 64);                     // [0..3] candidates per warp
  float* cand_v = reinterpret_cast<float*>(tail + 128);              // [4][MGB_CAND]
  int* cand_i = reinterpret_cast<int*>(tail + 128 + MGB_WARPS * MGB_CAND * 4);
  float* sel_v = reinterpret_cast<float*>(tail + 128 + MGB_WARPS * MGB_CAND * 8);          // [32] slow-path winners
  int* sel_i = reinterpret_cast<int*>(tail + 128 + MGB_WARPS * MGB_CAND * 8 + 128);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int k = p.topk;
  const int nvec = p.V >> 2;                                         // multiple of 128
  const float4* b4 = reinterpret_cast<const float4*>(buf);
  uint32_t phase = 0;
  auto fetch_row = [&](int r) {                                      // one bulk copy of a logit row: global -> shared
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(row_bytes));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(p.logits + static_cast<size_t>(r) * p.ld)),
                 "r"(static_cast<uint32_t>(row_bytes)), "r"(smem_u32(bar))
                 : "memory");
  };
  for (int row = blockIdx.x; row < p.M; row += gridDim.x) {
    if (threadIdx.x == 0) fetch_row(row);
    // ids[row] is only needed at the very end of the row, by one lane: fetched now, its global-memory latency used to sit at
    // the end of the serial selection tail
    long long id_row = 0;
    if (threadIdx.x == 0 && p.ids != nullptr) id_row = p.ids[row];
    mbar_wait(bar, phase);
    phase ^= 1;
    // ---- pass 1: per-thread maximum (four independent 3-input max chains) ----
    // Round 2: the thread's SECOND largest value is no longer tracked (three half-rate FMNMX per element, 41 % of the ALU pipe):
    // the k-th largest of a warp's 32 thread maxima is still a lower bound of the row's k-th largest value (the maxima are 32
    // distinct elements), only a slightly looser one — a few more candidates reach the lists of pass 2 (capacity 64 per warp).
    float t0;
    {
      float u0[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) u0[u] = -INFINITY;
      for (int c = threadIdx.x; c < nvec; c += MGB_WARPS * 32 * 4) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = b4[c + MGB_WARPS * 32 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          u0[u] = fmaxf(fmaxf(u0[u], q[u].x), q[u].y);
          u0[u] = fmaxf(fmaxf(u0[u], q[u].z), q[u].w);
        }
      }
      t0 = fmaxf(fmaxf(u0[0], u0[1]), fmaxf(u0[2], u0[3]));
    }
    const float mw = warp_max_f32(t0);
    // k-th largest of this warp's 32 thread maxima: a lower bound of the row's k-th largest value
    float a0 = t0, tau_w = -INFINITY;
    for (int r = 0; r < k; ++r) {
      const float bv = warp_max_f32(a0);
      tau_w = bv;
      const unsigned who = __ballot_sync(0xffffffffu, a0 == bv);
      if (lane == __ffs(who) - 1) a0 = -INFINITY;
    }
    if (lane == 0) {
      sh_f[warp] = mw;
      sh_f[4 + warp] = tau_w;
    }
    __syncthreads();
    const float m_row = fmaxf(fmaxf(sh_f[0], sh_f[1]), fmaxf(sh_f[2], sh_f[3]));
    const float tau = fmaxf(fmaxf(sh_f[4], sh_f[5]), fmaxf(sh_f[6], sh_f[7]));
    // ---- pass 2: softmax denominator and candidates >= tau ----
    float2 ps2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) ps2[u] = make_float2(0.0f, 0.0f);
    int ncand = 0;                                                    // warp-uniform
    float* cv_w = cand_v + warp * MGB_CAND;
    int* ci_w = cand_i + warp * MGB_CAND;
    const float nm = -m_row * 1.4426950408889634f;
    for (int c = threadIdx.x; c < nvec; c += MGB_WARPS * 32 * 4) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = b4[c + MGB_WARPS * 32 * u];
      // packed f32x2 arithmetic for the exponent argument and the partial sums; one compare per 16 elements (group maximum)
      float gmax = -INFINITY;
      const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f), nm2 = make_float2(nm, nm);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 a01 = __ffma2_rn(make_float2(q[u].x, q[u].y), l2e, nm2);
        const float2 a23 = __ffma2_rn(make_float2(q[u].z, q[u].w), l2e, nm2);
        float2 e01, e23;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e01.x) : "f"(a01.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e01.y) : "f"(a01.y));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e23.x) : "f"(a23.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e23.y) : "f"(a23.y));
        ps2[u] = __fadd2_rn(ps2[u], __fadd2_rn(e01, e23));
        gmax = fmaxf(fmaxf(gmax, q[u].x), q[u].y);
        gmax = fmaxf(fmaxf(gmax, q[u].z), q[u].w);
      }
      const bool any = gmax >= tau;
      if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const bool hit = e[t] >= tau;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) {
              const int pos = ncand + __popc(m & ((1u << lane) - 1u));
              if (pos < MGB_CAND) { cv_w[pos] = e[t]; ci_w[pos] = (c + MGB_WARPS * 32 * u) * 4 + t; }
            }
            ncand += __popc(m);
          }
        }
      }
    }
    float ssum = ((ps2[0].x + ps2[0].y) + (ps2[1].x + ps2[1].y)) + ((ps2[2].x + ps2[2].y) + (ps2[3].x + ps2[3].y));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    if (lane == 0) {
      sh_f[8 + warp] = ssum;
      sh_n[warp] = ncand;
    }
    __syncthreads();
    const bool overflow = sh_n[0] > MGB_CAND || sh_n[1] > MGB_CAND || sh_n[2] > MGB_CAND || sh_n[3] > MGB_CAND;
    if (overflow) {
      // degenerate row (masses of equal values): k rounds of a block-wide arg-max over the staged row, each round
      // taking the successor of the previous pick in (value desc, index asc) order — capacity-independent, few registers
      float pv = INFINITY;
      int pi = -1;
      for (int r = 0; r < k; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = threadIdx.x; c < nvec; c += MGB_WARPS * 32) {
          const float4 qq = b4[c];
          const float e[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int idx = c * 4 + t;
            const bool after = e[t] < pv || (e[t] == pv && idx > pi);
            if (after && (e[t] > bv || (e[t] == bv && idx < bi))) { bv = e[t]; bi = idx; }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { cand_v[warp] = bv; cand_i[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
          float fv = cand_v[0];
          int fi = cand_i[0];
          for (int w = 1; w < MGB_WARPS; ++w)
            if (cand_v[w] > fv || (cand_v[w] == fv && cand_i[w] < fi)) { fv = cand_v[w]; fi = cand_i[w]; }
          sel_v[r] = fv;
          sel_i[r] = fi;
        }
        __syncthreads();
        pv = sel_v[r];
        pi = sel_i[r];
      }
    }
    if (warp == 0) {
      const float s_row = (sh_f[8] + sh_f[9]) + (sh_f[10] + sh_f[11]);       // fixed order: deterministic
      float my_v = -INFINITY;
      int my_i = 0x7fffffff;
      if (!overflow) {
        // exact top-k among <= 256 candidates (value desc, index asc); lane r keeps winner r
        float cv[2 * MGB_WARPS];
        int ci[2 * MGB_WARPS];
#pragma unroll
        for (int w = 0; w < MGB_WARPS; ++w)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = h * 32 + lane;
            const bool ok = j < sh_n[w];
            cv[w * 2 + h] = ok ? cand_v[w * MGB_CAND + j] : -INFINITY;
            ci[w * 2 + h] = ok ? cand_i[w * MGB_CAND + j] : 0x7fffffff;
          }
        for (int r = 0; r < k; ++r) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
#pragma unroll
          for (int u = 0; u < 2 * MGB_WARPS; ++u)
            if (cv[u] > bv || (cv[u] == bv && ci[u] < bi)) { bv = cv[u]; bi = ci[u]; }
          const float wv = warp_max_f32(bv);                                           // best value of the warp ...
          const int wi = __reduce_min_sync(0xffffffffu, bv == wv ? bi : 0x7fffffff);      // ... at its lowest index
#pragma unroll
          for (int u = 0; u < 2 * MGB_WARPS; ++u)
            if (ci[u] == wi) { cv[u] = -INFINITY; ci[u] = 0x7fffffff; }
          if (lane == r) { my_v = wv; my_i = wi; }
        }
      } else if (lane < k) {
        my_v = sel_v[lane];
        my_i = sel_i[lane];
      }
      // ---- gumbel-perturbed arg-max over the k survivors (generate.py:45-46) ----
      float score = -INFINITY;
      if (lane < k && my_i != 0x7fffffff) {
        float u;
        if (p.noise != nullptr) u = p.noise[static_cast<size_t>(row) * p.ld_noise + my_i];
        else u = philox_uniform(p.seed, static_cast<uint32_t>(row), static_cast<uint32_t>(my_i), static_cast<uint32_t>(p.offset));
        const float inner = -logf(fmaxf(u, 1e-20f));
        const float g = -logf(fmaxf(inner, 1e-20f));
        score = my_v / fmaxf(p.temperature, 1e-10f) + g;
      }
      const float bs = warp_max_f32(score);                                            // best perturbed score ...
      const int bi = __reduce_min_sync(0xffffffffu, (lane < k && score == bs) ? my_i : 0x7fffffff);   // ... ties -> lower index
      const unsigned owner = __ballot_sync(0xffffffffu, lane < k && my_i == bi);
      const float bl = __shfl_sync(0xffffffffu, my_v, owner != 0 ? __ffs(owner) - 1 : 0);             // its logit
      if (lane == 0) {
        const long long pred = static_cast<long long>(bi);
        const float prob = __expf(bl - m_row) / s_row;           // softmax(logits)[pred], unfiltered, T = 1
        if (p.pred_ids != nullptr) p.pred_ids[row] = pred;
        bool is_mask = true;
        if (p.ids != nullptr) {
          is_mask = (id_row == p.mask_id);
          if (is_mask) p.ids[row] = pred;
        }
        if (p.scores != nullptr) p.scores[row] = is_mask ? (1.0f - prob) : -1e5f;
      }
    }
    // everyone is done with buf / lists before the next bulk copy overwrites them.  (Issuing the next row's copy
    // before the selection and dropping this sync was measured: 0.68 ms instead of 0.54 — slower.)
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// re-mask: per image, the k highest scores get mask_id (ties: lower token index first).
// rank_i = #{j : s_j > s_i  or (s_j == s_i and j < i)};  token i is re-masked iff rank_i < k.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
maskgit_remask_kernel(const float* __restrict__ scores, long long* __restrict__ ids, int N, int k, long long mask_id,
                      const StepScalars* __restrict__ step_tab, int* __restrict__ step_idx, int* __restrict__ ticket) {
  extern __shared__ float sm_scores[];
  if (step_tab != nullptr) k = step_tab[*step_idx].k;      // every block reads the step index before it takes its ticket below
  const int b = blockIdx.x;
  const float* s = scores + static_cast<size_t>(b) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sm_scores[i] = s[i];
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float si = sm_scores[i];
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < N; ++j) {
      const float sj = sm_scores[j];                  // broadcast read
      rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
    }
    if (rank < k) ids[static_cast<size_t>(b) * N + i] = mask_id;
  }
  if (step_tab != nullptr) {
    // last kernel of a MaskGIT step: the block that finishes last advances the step index for the next graph replay
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(ticket, 1) == static_cast<int>(gridDim.x) - 1) {
        *ticket = 0;
        *step_idx += 1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stage-2 TRAINING forward, the two pieces around the transformer:
//
// random_masking (generate.py:78-108): per image, the len_keep tokens with the SMALLEST noise keep their latent,
// the others are replaced by mask_token; mask = 1 where replaced.  The reference does argsort(noise) ->
// gather -> cat(mask tokens) -> gather(ids_restore); the result is the same as ranking every token:
//   rank_i = #{j : n_j < n_i or (n_j == n_i and j < i)};  masked_i = rank_i >= len_keep.
// One block per image; noise comes from the caller (parity tests) or Philox keyed on (seed, image, token, offset).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
maskgit_random_mask_kernel(const float* __restrict__ z, int64_t ldz, const float* __restrict__ noise,
                           unsigned long long seed, unsigned long long offset, const float* __restrict__ mask_token,
                           int N, int len_keep, float* __restrict__ mask, float* __restrict__ x_out) {
  extern __shared__ float sm_noise[];                 // [N] noise, then [N] flags
  float* flag = sm_noise + N;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    sm_noise[i] = noise != nullptr ? noise[static_cast<size_t>(b) * N + i]
                                   : philox_uniform(seed, static_cast<uint32_t>(b), static_cast<uint32_t>(i), static_cast<uint32_t>(offset));
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float ni = sm_noise[i];
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < N; ++j) {
      const float nj = sm_noise[j];                   // broadcast read
      rank += (nj < ni || (nj == ni && j < i)) ? 1 : 0;
    }
    const float m = rank >= len_keep ? 1.0f : 0.0f;
    flag[i] = m;
    mask[static_cast<size_t>(b) * N + i] = m;
  }
  __syncthreads();
  if (x_out == nullptr) return;
  // tokens: 8 threads per 32-float row (16 B each), coalesced
  for (int e = threadIdx.x; e < N * 8; e += blockDim.x) {
    const int i = e >> 3, c = (e & 7) * 4;
    const size_t row = static_cast<size_t>(b) * N + i;
    const float4 v = flag[i] != 0.0f ? *reinterpret_cast<const float4*>(mask_token + c)
                                     : *reinterpret_cast<const float4*>(z + row * ldz + c);
    *reinterpret_cast<float4*>(x_out + row * 32 + c) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Masked label-smoothed cross entropy (generate.py:110-123: F.cross_entropy(logit, label, label_smoothing = eps,
// reduction = 'none'), then (loss * masks).sum() / masks.sum()).  Per row
//   loss = (1 - eps) * (lse - x[label]) + eps * (lse - mean(x)) = lse - (1 - eps) * x[label] - eps * mean(x)
// One warp streams one fp32 logit row ONCE (8 x 16-byte loads in flight per lane): online max / sum-exp2 / sum.
// Rows with mask == 0 contribute nothing to the reference's result and are not read at all (at mask_ratio 0.75
// that is a quarter of the 4*M*V bytes).  row_loss[row] = loss * mask; a single-block second kernel reduces
// row_loss and mask in a fixed order in fp64 (deterministic) and forms the scalar.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float* __restrict__ logits, int64_t ld, int M, int V, const long long* __restrict__ label,
               const float* __restrict__ mask, float eps, float* __restrict__ row_loss) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float mk = mask != nullptr ? mask[row] : 1.0f;
  if (mk == 0.0f) {
    if (lane == 0) row_loss[row] = 0.0f;
    return;
  }
  const float4* x4 = reinterpret_cast<const float4*>(logits + static_cast<size_t>(row) * ld);
  const int nvec = V >> 2;
  constexpr float L2E = 1.4426950408889634f;
  float m = -INFINITY, s = 0.0f, t = 0.0f;            // lane-local: running max, sum exp(x - m), sum x
  for (int c0 = lane; c0 < nvec; c0 += 32 * 8) {
    float4 q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + u * 32;
      q[u] = c < nvec ? __ldcs(x4 + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    float cm = m;
#pragma unroll
    for (int u = 0; u < 8; ++u) cm = fmaxf(cm, fmaxf(fmaxf(q[u].x, q[u].y), fmaxf(q[u].z, q[u].w)));
    if (cm > m) {
      s *= exp2f((m - cm) * L2E);                     // m = -inf: s is 0 anyway
      m = cm;
    }
    const float nm = -m * L2E;
    float ps[4] = {0.0f, 0.0f, 0.0f, 0.0f}, pt[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (c0 + u * 32 < nvec) {
        const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float ex;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(e[k], L2E, nm)));
          ps[k] += ex;
          pt[k] += e[k];
        }
      }
    }
    s += (ps[0] + ps[1]) + (ps[2] + ps[3]);
    t += (pt[0] + pt[1]) + (pt[2] + pt[3]);
  }
  float mw = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  float sw = (m == -INFINITY) ? 0.0f : s * exp2f((m - mw) * L2E);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  if (lane == 0) {
    const long long y = label[row];
    const float xy = (y >= 0 && y < V) ? logits[static_cast<size_t>(row) * ld + y] : 0.0f;
    const float lse = mw + logf(sw);
    row_loss[row] = (lse - (1.0f - eps) * xy - eps * (t / static_cast<float>(V))) * mk;
  }
}

__global__ void __launch_bounds__(1024)
ce_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ mask, int M, float* __restrict__ loss_out,
                 double* __restrict__ sums_out) {
  __shared__ double sh_l[1024], sh_m[1024];
  double al = 0.0, am = 0.0;
  for (int i = threadIdx.x; i < M; i += 1024) {       // fixed assignment and order: deterministic
    al += static_cast<double>(row_loss[i]);
    am += mask != nullptr ? static_cast<double>(mask[i]) : 1.0;
  }
  sh_l[threadIdx.x] = al;
  sh_m[threadIdx.x] = am;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh_l[threadIdx.x] += sh_l[threadIdx.x + o];
      sh_m[threadIdx.x] += sh_m[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (loss_out != nullptr) *loss_out = static_cast<float>(sh_l[0] / sh_m[0]);     // masks.sum() == 0 -> nan, as the reference
    if (sums_out != nullptr) {
      sums_out[0] = sh_l[0];
      sums_out[1] = sh_m[0];
    }
  }
}

int pm_maskgit_random_mask_launch(const float* z, int64_t ldz, const float* noise, unsigned long long seed,
                                  unsigned long long offset, const float* mask_token, int B, int N, int len_keep,
                                  float* mask, float* x_out, cudaStream_t stream) {
  if (mask == nullptr || B <= 0 || N <= 0 || N > 12288 || len_keep < 0 || len_keep > N) return PM_ERR_INVALID;
  if (x_out != nullptr && (z == nullptr || mask_token == nullptr || (ldz % 4) != 0)) return PM_ERR_INVALID;
  const int threads = N < 1024 ? ((N + 31) / 32) * 32 : 1024;
  maskgit_random_mask_kernel<<<B, threads, 2 * N * sizeof(float), stream>>>(z, ldz, noise, seed, offset, mask_token, N,
                                                                            len_keep, mask, x_out);
  return static_cast<int>(cudaGetLastError());
}

int pm_ce_label_smooth_launch(const float* logits, int64_t ld, int M, int V, const long long* label, const float* mask,
                              float eps, float* row_loss, float* loss_out, double* sums_out, cudaStream_t stream) {
  if (logits == nullptr || label == nullptr || row_loss == nullptr || M <= 0 || V <= 0 || (V & 3) != 0 || (ld & 3) != 0)
    return PM_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(logits) & 15) != 0) return PM_ERR_INVALID;
  ce_rows_kernel<<<(M + 7) / 8, 256, 0, stream>>>(logits, ld, M, V, label, mask, eps, row_loss);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (loss_out != nullptr || sums_out != nullptr) {
    ce_reduce_kernel<<<1, 1024, 0, stream>>>(row_loss, mask, M, loss_out, sums_out);
    e = cudaGetLastError();
  }
  return static_cast<int>(e);
}

int pm_maskgit_sample_launch(const MaskgitParams& p, cudaStream_t stream) {
  if (p.logits == nullptr || p.M <= 0 || p.V <= 0 || (p.V & 3) != 0 || (p.ld & 3) != 0) return PM_ERR_INVALID;
  if (p.topk < 1 || p.topk > 32 || p.topk > p.V) return PM_ERR_INVALID;
  const int row_bytes = p.V * 4;
  static int variant = -1;       // PM_MASKGIT_VARIANT (tuning aid): 0 = auto (block-per-row staged), 1 = warp-per-row staged, 2 = streaming
  if (variant < 0) {
    const char* env = getenv("PM_MASKGIT_VARIANT");
    variant = env != nullptr ? atoi(env) : 0;
  }
  if (variant == 0 && (p.V % 2048) == 0 && row_bytes <= 65536 && (reinterpret_cast<uintptr_t>(p.logits) & 15) == 0) {
    // block-per-row staged kernel: 4 warps per row buffer, as many blocks per SM as fit (6 for V = 8192)
    const int smem = row_bytes + 128 + MGB_WARPS * MGB_CAND * 8 + 256;
    static bool attr_done_b[PM_MAX_DEVICES] = {};
    if (const int rc = pm_ensure_dyn_smem(maskgit_sample_block_kernel, 65536 + 4096, attr_done_b)) return rc;
    int per_sm = (232448 - 1024) / (smem + 1024);
    if (per_sm > 8) per_sm = 8;
    int blocks = pm_num_sms() * per_sm;
    if (blocks > p.M) blocks = p.M;
    maskgit_sample_block_kernel<<<blocks, MGB_WARPS * 32, smem, stream>>>(p);
    return static_cast<int>(cudaGetLastError());
  }
  if (variant != 2 && (p.V % 128) == 0 && row_bytes <= 65536 && (reinterpret_cast<uintptr_t>(p.logits) & 15) == 0) {
    // shared-memory staged kernel: as many row buffers per SM as fit (6 x 32 KB for V = 8192)
    constexpr int W = 6;
    const int smem = W * row_bytes + 64 + W * MG_CAND * 8;
    if (smem <= 232448 - 1024) {
      static bool attr_done[PM_MAX_DEVICES] = {};
      if (const int rc = pm_ensure_dyn_smem(maskgit_sample_smem_kernel<W>, 232448 - 1024, attr_done)) return rc;
      const int per_sm = (232448 - 1024) / smem;                      // co-resident blocks when rows are short
      int blocks = pm_num_sms() * (per_sm > 4 ? 4 : per_sm);
      const int need = (p.M + W - 1) / W;
      if (blocks > need) blocks = need;
      maskgit_sample_smem_kernel<W><<<blocks, W * 32, smem, stream>>>(p);
      return static_cast<int>(cudaGetLastError());
    }
  }
  const int threads = 256;
  const int blocks = (p.M + 7) / 8;
  if (p.topk <= 8) maskgit_sample_kernel<8><<<blocks, threads, 0, stream>>>(p);
  else maskgit_sample_kernel<32><<<blocks, threads, 0, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}

int pm_maskgit_remask_launch(const float* scores, long long* ids, int B, int N, int k, long long mask_id, const StepScalars* step_tab,
                             int* step_idx, int* ticket, cudaStream_t stream) {
  if (scores == nullptr || ids == nullptr || B <= 0 || N <= 0 || k < 0 || N > 12288) return PM_ERR_INVALID;
  if (step_tab != nullptr && (step_idx == nullptr || ticket == nullptr)) return PM_ERR_INVALID;
  const int threads = N < 1024 ? ((N + 31) / 32) * 32 : 1024;
  maskgit_remask_kernel<<<B, threads, N * sizeof(float), stream>>>(scores, ids, N, k, mask_id, step_tab, step_idx, ticket);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
