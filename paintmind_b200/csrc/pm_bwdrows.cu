// pm_bwdrows.cu — HBM-bound row kernels of the generator BACKWARD path: what the reference
// obtains from autograd for nn.LayerNorm (stage1/layers.py:49,51,89,128), SwiGLU (modules/mlp.py:29-30), the
// softmax row term of attention (modules/attention.py:55-57), the straight-through VectorQuantizer
// (stage1/quantize.py:19,29-36) and the clamp + un-patchify of the decoder output (vqmodel.py:30, layers.py:150).
// One warp per row (or a few lanes per row), 16-byte accesses, fp32 math; reductions over rows in a fixed order.
#include "pm_common.cuh"
#include "pm_kernels.h"

namespace pm {

__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
  v[0] = bf16lo_to_f32(u.x); v[1] = bf16hi_to_f32(u.x);
  v[2] = bf16lo_to_f32(u.y); v[3] = bf16hi_to_f32(u.y);
  v[4] = bf16lo_to_f32(u.z); v[5] = bf16hi_to_f32(u.z);
  v[6] = bf16lo_to_f32(u.w); v[7] = bf16hi_to_f32(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------------------------
// LayerNorm backward.  n = xhat * gamma + beta, xhat = (x - mu) * rstd:
//   g = dn * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres);  dgamma += dn * xhat;  dbeta += dn
// (mu, rstd) are recomputed from the row (it is in registers anyway).  Each block walks rows blockIdx.x * 8 + warp,
// += gridDim.x * 8 and leaves its partial (dgamma | dbeta) in part[blockIdx.x, 2, D]; colreduce_kernel sums the blocks.
// ----------------------------------------------------------------------------------------------
// Round 2: gamma lives in registers for the whole launch (ncu on the forward kernel: re-reading parameters per row saturates the
// L1 data stage, not HBM) and the NEXT row's three vectors per lane (x, dn, dres) are requested before the current row's four
// dependent reductions — dres used to be loaded behind them.
template <int NV>
__global__ void __launch_bounds__(256, NV <= 2 ? 2 : 1)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dn, int64_t lddn, const __nv_bfloat16* __restrict__ x, int64_t ldx,
              const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres, int64_t ldres,
              __nv_bfloat16* __restrict__ dx, int64_t lddx, int M, int D, float eps, float* __restrict__ part) {
  extern __shared__ float red[];                 // [8][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  float ag[NV][8], ab[NV][8], gm[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const bool in = i * 32 + lane < nvec;
    const float4 g0 = in ? __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 8)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 g1 = in ? __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 8 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    gm[i][0] = g0.x; gm[i][1] = g0.y; gm[i][2] = g0.z; gm[i][3] = g0.w;
    gm[i][4] = g1.x; gm[i][5] = g1.y; gm[i][6] = g1.z; gm[i][7] = g1.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      ag[i][k] = 0.f;
      ab[i][k] = 0.f;
    }
  }
  const float inv_d = 1.0f / static_cast<float>(D);
  const int step = gridDim.x * 8;
  uint4 ux[NV], ug[NV], ur[NV];
  auto load_row = [&](int row) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = i * 32 + lane;
      ux[i] = ug[i] = ur[i] = make_uint4(0u, 0u, 0u, 0u);
      if (vi < nvec && row < M) {
        ux[i] = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * ldx + vi * 8);
        ug[i] = *reinterpret_cast<const uint4*>(dn + static_cast<size_t>(row) * lddn + vi * 8);
        if (dres != nullptr) ur[i] = *reinterpret_cast<const uint4*>(dres + static_cast<size_t>(row) * ldres + vi * 8);
      }
    }
  };
  int row = blockIdx.x * 8 + warp;
  load_row(row);
  for (; row < M; row += step) {
    float xv[NV][8], gv[NV][8], rv[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      unpack8(ux[i], xv[i]);
      unpack8(ug[i], gv[i]);
      unpack8(ur[i], rv[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += xv[i][k];              // lanes beyond the row hold zeros
    }
    load_row(row + step);                                      // in flight under this row's reductions
    const float mu = warp_sum(s) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i * 32 + lane < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = xv[i][k] - mu;
          sq = fmaf(d, d, sq);
        }
      }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i * 32 + lane < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (xv[i][k] - mu) * rstd;
          const float dnv = gv[i][k];
          ag[i][k] = fmaf(dnv, xh, ag[i][k]);
          ab[i][k] += dnv;
          const float g = dnv * gm[i][k];
          xv[i][k] = xh;
          gv[i][k] = g;
          c1 += g;
          c2 = fmaf(g, xh, c2);
        }
      }
    c1 = warp_sum(c1) * inv_d;
    c2 = warp_sum(c2) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = i * 32 + lane;
      if (vi < nvec) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = rstd * (gv[i][k] - c1 - xv[i][k] * c2);
        if (dres != nullptr) {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += rv[i][k];
        }
        *reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * lddx + vi * 8) = pack8(o);
      }
    }
  }
  // block reduction of the per-warp (dgamma, dbeta) accumulators, one after the other through the same buffer
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = i * 32 + lane;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) red[warp * D + vi * 8 + k] = which == 0 ? ag[i][k] : ab[i][k];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w * D + c];
      part[(static_cast<size_t>(blockIdx.x) * 2 + which) * D + c] = t;
    }
  }
}

int pm_ln_bwd_blocks(int M) {
  const int want = (M + 7) / 8, cap = 2 * pm_num_sms();       // two resident blocks per SM: one wave
  return want < cap ? want : cap;
}

int pm_ln_bwd_launch(const void* dn, int64_t lddn, const void* x, int64_t ldx, const float* gamma, const void* dres,
                     int64_t ldres, void* dx, int64_t lddx, int M, int D, float eps, float* part, float* dgamma_dbeta,
                     cudaStream_t stream) {
  if (dn == nullptr || x == nullptr || gamma == nullptr || dx == nullptr || part == nullptr || dgamma_dbeta == nullptr) return PM_ERR_INVALID;
  if (M <= 0 || D <= 0 || (D % 8) != 0 || D > 1024 || (lddn % 8) != 0 || (ldx % 8) != 0 || (lddx % 8) != 0 || (ldres % 8) != 0) return PM_ERR_INVALID;
  const int blocks = pm_ln_bwd_blocks(M);
  const int smem = 8 * D * 4;
  auto a = reinterpret_cast<const __nv_bfloat16*>(dn);
  auto b = reinterpret_cast<const __nv_bfloat16*>(x);
  auto c = reinterpret_cast<const __nv_bfloat16*>(dres);
  auto d = reinterpret_cast<__nv_bfloat16*>(dx);
  if (D <= 256) ln_bwd_kernel<1><<<blocks, 256, smem, stream>>>(a, lddn, b, ldx, gamma, c, ldres, d, lddx, M, D, eps, part);
  else if (D <= 512) ln_bwd_kernel<2><<<blocks, 256, smem, stream>>>(a, lddn, b, ldx, gamma, c, ldres, d, lddx, M, D, eps, part);
  else ln_bwd_kernel<4><<<blocks, 256, smem, stream>>>(a, lddn, b, ldx, gamma, c, ldres, d, lddx, M, D, eps, part);
  return pm_colreduce_launch(part, blocks, 2 * D, dgamma_dbeta, 0, stream);
}

// ----------------------------------------------------------------------------------------------
// SwiGLU backward (modules/mlp.py:29-30: hidden = silu(x1) * x2).  x12 is the [M, 2 * hp] output of the w12
// projection in the tile layout of the packed weight (per 256 columns: 128 gate columns, then their 128 value columns);
// dh is the gradient of the hidden [M, hp].  Writes h (operand of the w3 weight gradient) and d12 in x12's layout.
// ----------------------------------------------------------------------------------------------
// Block = a strip of 256 hidden columns x a chunk of rows (warp w takes rows w, w + 8, ... of the chunk, lane l the 8 hidden
// columns 8 l .. 8 l + 7 of the strip: 512-byte contiguous runs per warp and tensor).  Walking rows inside the block lets each
// thread keep the column sums of its 16 outputs, so the bias gradient of w12 (sum over tokens of d12, mlp.py:28) costs no
// second pass over the 1.5 GB d12 tensor: per-block partial sums go to `part[chunk, 2 hp]`, colreduce adds them in order.
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ x12, int64_t ld12, const __nv_bfloat16* __restrict__ dh, int64_t lddh,
                  __nv_bfloat16* __restrict__ h, int64_t ldh, __nv_bfloat16* __restrict__ d12, int64_t ldd12, int M, int hp,
                  int rows_per, float* __restrict__ part) {
  __shared__ float red[8][32][17];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 256 + lane * 8;                        // hidden column
  const bool col_ok = j < hp;
  const int gcol = (j >> 7) * 256 + (j & 127);
  const int r0 = blockIdx.y * rows_per;
  const int r1 = r0 + rows_per < M ? r0 + rows_per : M;
  float sg_[8], sv_[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sg_[k] = 0.f, sv_[k] = 0.f;
  if (col_ok) {
    // two rows per trip: six 16-byte loads in flight per lane (one row at a time ran at 4.2 TB/s instead of 6.6)
    for (long long row = r0 + warp; row < r1; row += 16) {
      const bool two = row + 8 < r1;
      const long long rowb = two ? row + 8 : row;
      const uint4 ug0 = *reinterpret_cast<const uint4*>(x12 + row * ld12 + gcol), uv0 = *reinterpret_cast<const uint4*>(x12 + row * ld12 + gcol + 128);
      const uint4 ud0 = *reinterpret_cast<const uint4*>(dh + row * lddh + j);
      const uint4 ug1 = *reinterpret_cast<const uint4*>(x12 + rowb * ld12 + gcol), uv1 = *reinterpret_cast<const uint4*>(x12 + rowb * ld12 + gcol + 128);
      const uint4 ud1 = *reinterpret_cast<const uint4*>(dh + rowb * lddh + j);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 1 && !two) break;
        const long long rr = half == 0 ? row : rowb;
        float g[8], v[8], d[8];
        unpack8(half == 0 ? ug0 : ug1, g);
        unpack8(half == 0 ? uv0 : uv1, v);
        unpack8(half == 0 ? ud0 : ud1, d);
        float ho[8], dg[8], dv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float sg = 1.0f / (1.0f + __expf(-g[k]));
          const float si = g[k] * sg;
          ho[k] = si * v[k];
          dv[k] = d[k] * si;
          dg[k] = d[k] * v[k] * sg * (1.0f + g[k] * (1.0f - sg));
          sg_[k] += dg[k];
          sv_[k] += dv[k];
        }
        if (h != nullptr) *reinterpret_cast<uint4*>(h + rr * ldh + j) = pack8(ho);
        *reinterpret_cast<uint4*>(d12 + rr * ldd12 + gcol) = pack8(dg);
        *reinterpret_cast<uint4*>(d12 + rr * ldd12 + gcol + 128) = pack8(dv);
      }
    }
  }
  if (part == nullptr) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) red[warp][lane][k] = sg_[k], red[warp][lane][8 + k] = sv_[k];
  __syncthreads();
  // thread (w, l) finishes outputs 2 w and 2 w + 1 of lane l's sixteen
  if (col_ok) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int o = 2 * warp + e;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][lane][o];
      part[static_cast<size_t>(blockIdx.y) * (2 * hp) + gcol + (o < 8 ? o : 128 + o - 8)] = t;
    }
  }
}

int pm_swiglu_bwd_chunks(int M, int hp) {
  const int strips = (hp + 255) / 256;
  int R = (8 * pm_num_sms() + strips - 1) / strips;
  const int max_r = (M + 63) / 64;
  if (R > max_r) R = max_r;
  if (R < 1) R = 1;
  const int rows_per = (M + R - 1) / R;
  return (M + rows_per - 1) / rows_per;
}

int pm_swiglu_bwd_launch(const void* x12, int64_t ld12, const void* dh, int64_t lddh, void* h, int64_t ldh, void* d12,
                         int64_t ldd12, int M, int hp, float* work, float* b12, cudaStream_t stream) {
  if (x12 == nullptr || dh == nullptr || d12 == nullptr || M <= 0 || hp <= 0 || (hp % 128) != 0) return PM_ERR_INVALID;
  if ((ld12 % 8) != 0 || (lddh % 8) != 0 || (ldh % 8) != 0 || (ldd12 % 8) != 0) return PM_ERR_INVALID;
  if ((b12 != nullptr) != (work != nullptr)) return PM_ERR_INVALID;
  const int R = pm_swiglu_bwd_chunks(M, hp);
  const int rows_per = (M + R - 1) / R;
  dim3 grid((hp + 255) / 256, R);
  swiglu_bwd_kernel<<<grid, 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x12), ld12, reinterpret_cast<const __nv_bfloat16*>(dh), lddh,
      reinterpret_cast<__nv_bfloat16*>(h), ldh, reinterpret_cast<__nv_bfloat16*>(d12), ldd12, M, hp, rows_per, work);
  if (b12 != nullptr) return pm_colreduce_launch(work, R, 2 * hp, b12, 0, stream);
  return static_cast<int>(cudaGetLastError());
}

// ----------------------------------------------------------------------------------------------
// nds[b, h, n] = -scale * sum_d dO[b, n, h*64 + d] * O[b, n, h*64 + d]   (the softmax-backward row term, pre-scaled and
// negated so that the attention backward kernel applies it with one packed FMA) and nlse = -lse.
// One warp per token, 8 columns per lane per pass; a head = 8 consecutive lanes.
// ----------------------------------------------------------------------------------------------
template <bool O32>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const void* __restrict__ o_, int64_t ldo, int64_t bso, const __nv_bfloat16* __restrict__ dO, int64_t lddo,
                  int64_t bsdo, int B, int H, int N, const float* __restrict__ lse, float scale, float* __restrict__ nds,
                  float* __restrict__ nlse, int64_t delta_ld) {
  const long long tok = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= static_cast<long long>(B) * N) return;
  const int b = static_cast<int>(tok / N), n = static_cast<int>(tok % N);
  const __nv_bfloat16* drow = dO + b * bsdo + static_cast<int64_t>(n) * lddo;
  for (int c0 = 0; c0 < H * 64; c0 += 256) {
    const int c = c0 + lane * 8;
    float s = 0.f;
    if (c < H * 64) {
      float a[8], d[8];
      if (O32) {
        const float* orow = reinterpret_cast<const float*>(o_) + b * bso + static_cast<int64_t>(n) * ldo + c;
        const float4 a0 = *reinterpret_cast<const float4*>(orow), a1 = *reinterpret_cast<const float4*>(orow + 4);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      } else {
        unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(o_) + b * bso + static_cast<int64_t>(n) * ldo + c), a);
      }
      unpack8(*reinterpret_cast<const uint4*>(drow + c), d);
#pragma unroll
      for (int k = 0; k < 8; ++k) s = fmaf(a[k], d[k], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0 && c < H * 64) {
      const size_t at = (static_cast<size_t>(b) * H + (c >> 6)) * delta_ld + n;
      nds[at] = -scale * s;
      nlse[at] = -lse[at];
    }
  }
}

int pm_attn_delta_launch(const void* o, int o_is_f32, int64_t ldo, int64_t bso, const void* dO, int64_t lddo, int64_t bsdo, int B, int H,
                         int N, const float* lse, float scale, float* nds, float* nlse, int64_t delta_ld, cudaStream_t stream) {
  if (o == nullptr || dO == nullptr || nds == nullptr || nlse == nullptr || lse == nullptr || B <= 0 || H <= 0 || N <= 0) return PM_ERR_INVALID;
  if ((ldo % 8) != 0 || (lddo % 8) != 0 || (bso % 8) != 0 || (bsdo % 8) != 0) return PM_ERR_INVALID;
  const long long threads = static_cast<long long>(B) * N * 32;
  const unsigned blocks = static_cast<unsigned>((threads + 255) / 256);
  if (o_is_f32)
    attn_delta_kernel<true><<<blocks, 256, 0, stream>>>(o, ldo, bso, reinterpret_cast<const __nv_bfloat16*>(dO), lddo, bsdo, B, H, N, lse, scale, nds, nlse, delta_ld);
  else
    attn_delta_kernel<false><<<blocks, 256, 0, stream>>>(o, ldo, bso, reinterpret_cast<const __nv_bfloat16*>(dO), lddo, bsdo, B, H, N, lse, scale, nds, nlse, delta_ld);
  return static_cast<int>(cudaGetLastError());
}

// ----------------------------------------------------------------------------------------------
// VectorQuantizer backward (stage1/quantize.py:19,29-36), e_dim = 32, 8 lanes per row (float4 each):
//   zn = l2norm(z); en = l2norm(E[idx]); loss = beta * mean((sg(en) - zn)^2) + mean((en - sg(zn))^2); out = zn + sg(en - zn)
//   dzn = d_out + beta * c * (zn - en),  c = 2 * d_loss / (M * 32);   dz = (dzn - zn (zn . dzn)) / max(|z|, eps)
//   den = c * (en - zn);  dE[idx] += (den - en (en . den)) / max(|E[idx]|, eps)     (fp32 atomics)
// dz leaves as fp32 and as the [hi | lo] bf16 split the prev_quant dgrad / wgrad GEMMs consume.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__global__ void __launch_bounds__(256)
vq_bwd_kernel(const float* __restrict__ z, int64_t ldz, const long long* __restrict__ idx, const float* __restrict__ E,
              const float* __restrict__ d_out, int64_t ldd, const float* __restrict__ d_loss, float beta, long long M,
              float* __restrict__ dz, __nv_bfloat16* __restrict__ dz_split, float* __restrict__ dE) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long row = t >> 3;
  const int l = static_cast<int>(t & 7);
  const bool ok = row < M;
  const long long r = ok ? row : 0;
  const long long id = idx[r];
  const float4 zv = *reinterpret_cast<const float4*>(z + r * ldz + l * 4);
  const float4 ev = *reinterpret_cast<const float4*>(E + id * 32 + l * 4);
  float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
  if (d_out != nullptr) go = *reinterpret_cast<const float4*>(d_out + r * ldd + l * 4);
  const float c = 2.0f * (d_loss != nullptr ? *d_loss : 0.0f) / (static_cast<float>(M) * 32.0f);
  const float nz = fmaxf(sqrtf(sum8(dot4(zv, zv))), 1e-12f), ne = fmaxf(sqrtf(sum8(dot4(ev, ev))), 1e-12f);
  const float iz = 1.0f / nz, ie = 1.0f / ne;
  const float4 zn = make_float4(zv.x * iz, zv.y * iz, zv.z * iz, zv.w * iz);
  const float4 en = make_float4(ev.x * ie, ev.y * ie, ev.z * ie, ev.w * ie);
  const float bc = beta * c;
  const float4 dzn = make_float4(go.x + bc * (zn.x - en.x), go.y + bc * (zn.y - en.y), go.z + bc * (zn.z - en.z), go.w + bc * (zn.w - en.w));
  const float pz = sum8(dot4(zn, dzn));
  const float4 o = make_float4((dzn.x - zn.x * pz) * iz, (dzn.y - zn.y * pz) * iz, (dzn.z - zn.z * pz) * iz, (dzn.w - zn.w * pz) * iz);
  const float4 den = make_float4(c * (en.x - zn.x), c * (en.y - zn.y), c * (en.z - zn.z), c * (en.w - zn.w));
  const float pe = sum8(dot4(en, den));
  if (!ok) return;
  if (dz != nullptr) *reinterpret_cast<float4*>(dz + row * 32 + l * 4) = o;
  if (dz_split != nullptr) {
    const float v[4] = {o.x, o.y, o.z, o.w};
    uint32_t hi[2], lo[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      hi[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
      lo[k] = pack_bf16x2(v[2 * k] - bf16lo_to_f32(hi[k]), v[2 * k + 1] - bf16hi_to_f32(hi[k]));
    }
    *reinterpret_cast<uint2*>(dz_split + row * 64 + l * 4) = make_uint2(hi[0], hi[1]);
    *reinterpret_cast<uint2*>(dz_split + row * 64 + 32 + l * 4) = make_uint2(lo[0], lo[1]);
  }
  if (dE != nullptr && c != 0.0f) {
    float* dst = dE + id * 32 + l * 4;
    atomicAdd(dst + 0, (den.x - en.x * pe) * ie);
    atomicAdd(dst + 1, (den.y - en.y * pe) * ie);
    atomicAdd(dst + 2, (den.z - en.z * pe) * ie);
    atomicAdd(dst + 3, (den.w - en.w * pe) * ie);
  }
}

int pm_vq_bwd_launch(const float* z, int64_t ldz, const long long* idx, const float* E, const float* d_out, int64_t ldd,
                     const float* d_loss, float beta, int M, float* dz, void* dz_split, float* dE, cudaStream_t stream) {
  if (z == nullptr || idx == nullptr || E == nullptr || M <= 0 || (ldz % 4) != 0 || (ldd % 4) != 0) return PM_ERR_INVALID;
  const long long threads = static_cast<long long>(M) * 8;
  vq_bwd_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(z, ldz, idx, E, d_out, ldd, d_loss, beta, M, dz,
                                                                               reinterpret_cast<__nv_bfloat16*>(dz_split), dE);
  return static_cast<int>(cudaGetLastError());
}

// ----------------------------------------------------------------------------------------------
// Backward of clamp(-1, 1) + un-patchify (vqmodel.py:30, layers.py:150): the image gradient fp32 NCHW goes back to
// token rows bf16 [B * gh * gw, C * 64] in (c, p1, p2) column order (the order of the permuted `proj` rows the forward
// store uses), zeroed where the reconstruction sits on a clamp bound.
// ----------------------------------------------------------------------------------------------
__global__ void patchify8_masked_kernel(const float* __restrict__ g, const float* __restrict__ rec, __nv_bfloat16* __restrict__ out,
                                        int B, int C, int H, int W) {
  const int gw = W >> 3, gh = H >> 3;
  const long long total = static_cast<long long>(B) * C * H * gw;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int tw = static_cast<int>(t % gw);
  long long r = t / gw;
  const int y = static_cast<int>(r % H); r /= H;
  const int c = static_cast<int>(r % C);
  const int b = static_cast<int>(r / C);
  const size_t off = ((static_cast<size_t>(b) * C + c) * H + y) * W + tw * 8;
  const float4 v0 = *reinterpret_cast<const float4*>(g + off), v1 = *reinterpret_cast<const float4*>(g + off + 4);
  const float4 r0 = *reinterpret_cast<const float4*>(rec + off), r1 = *reinterpret_cast<const float4*>(rec + off + 4);
  auto m = [](float gv, float rv) { return fabsf(rv) < 1.0f ? gv : 0.0f; };
  uint4 o;
  o.x = pack_bf16x2(m(v0.x, r0.x), m(v0.y, r0.y)); o.y = pack_bf16x2(m(v0.z, r0.z), m(v0.w, r0.w));
  o.z = pack_bf16x2(m(v1.x, r1.x), m(v1.y, r1.y)); o.w = pack_bf16x2(m(v1.z, r1.z), m(v1.w, r1.w));
  const int th = y >> 3, kh = y & 7;
  const size_t row = (static_cast<size_t>(b) * gh + th) * gw + tw;
  *reinterpret_cast<uint4*>(out + row * (static_cast<size_t>(C) * 64) + c * 64 + kh * 8) = o;
}

int pm_unpatchify_bwd_launch(const float* g, const float* rec, void* out, int B, int C, int H, int W, cudaStream_t stream) {
  if (g == nullptr || rec == nullptr || out == nullptr || (H % 8) != 0 || (W % 8) != 0 || B <= 0 || C <= 0) return PM_ERR_INVALID;
  const long long total = static_cast<long long>(B) * C * H * (W / 8);
  patchify8_masked_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(g, rec, reinterpret_cast<__nv_bfloat16*>(out), B, C, H, W);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
