// pm_rowops.cu — HBM-bound row kernels of the tokenizer path (sm_100a): patch extraction,
// LayerNorm statistics / LayerNorm.  One warp per row, 16-byte vector accesses, fp32 math.
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <stdlib.h>

namespace pm {

// ----------------------------------------------------------------------------------------------
// im2col for the stride-P patch conv (stage1/layers.py:82-83): img fp32 NCHW -> bf16 [B*gh*gw, C*P*P]
// with K order (c, kh, kw), i.e. the flattened Conv2d weight layout.  P == 8.
// thread = (b, c, y, tw): reads 8 contiguous floats (coalesced along the image row), writes 16 B.
// ----------------------------------------------------------------------------------------------
__global__ void patchify8_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out,
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
  const float* src = img + ((static_cast<size_t>(b) * C + c) * H + y) * W + tw * 8;
  const float4 v0 = *reinterpret_cast<const float4*>(src);
  const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
  uint4 o;
  o.x = pack_bf16x2(v0.x, v0.y); o.y = pack_bf16x2(v0.z, v0.w);
  o.z = pack_bf16x2(v1.x, v1.y); o.w = pack_bf16x2(v1.z, v1.w);
  const int th = y >> 3, kh = y & 7;
  const size_t row = (static_cast<size_t>(b) * gh + th) * gw + tw;
  *reinterpret_cast<uint4*>(out + row * (static_cast<size_t>(C) * 64) + c * 64 + kh * 8) = o;
}

// ----------------------------------------------------------------------------------------------
// Same im2col, fed straight from decoded pixels: uint8 NHWC [B, H, W, 3] (PIL / numpy layout).  Fuses the
// reference's ingest transform (utils/transform.py:17-18: T.ToTensor() = u / 255, T.Normalize(0.5, 0.5) =
// (t - 0.5) / 0.5, both in fp32 — reproduced here operation by operation, so the result is bit-identical to
// patchify8 of the fp32 tensor the reference would have built) into the patch extraction: 3 B/px are read instead
// of 12 B/px and the fp32 NCHW image (4x the size of the pixels) never exists.
// thread = (b, y, tw): reads 8 pixels x 3 channels = 24 contiguous bytes, writes three 16-byte rows.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float norm_u8(uint32_t u) {
  const float t = __fdiv_rn(static_cast<float>(u), 255.0f);          // ToTensor
  return __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);                        // Normalize(mean 0.5, std 0.5)
}

// The two IEEE divisions per byte made the first version compute-bound (2.2 TB/s); the second tabulated the 256 possible
// results in shared memory — 24 two-byte look-ups per thread at random addresses, i.e. ~4-way bank conflicts on each: that table
// was what held the kernel at 62 % of the copy peak.  Now: bf16(fma(u, fp32(2/255), -1)) — it equals bf16 of the reference's chain
// for ALL 256 byte values (tests/test_host_logic.py::test_u8_normalise_formula_is_exact checks the 256 cases), one PRMT + one
// FFMA per byte: the byte goes into the mantissa of 2^23 (0x4B0000uu = 8388608 + u) and the fused multiply-add removes the
// offset again exactly (8388608 * c + 1 = 65794.0078125 is representable).
__device__ __forceinline__ float norm_u8_fast(uint32_t word, int byte) {
  const float f = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u + static_cast<uint32_t>(byte)));     // 8388608 + u
  return __fmaf_rn(f, 0.007843137718737125f, -65794.0078125f);
}
//
// Round 2: one block iteration = one ROW OF PATCHES (8 image rows x W pixels in, gw patch rows of 384 B out — contiguous in
// `out`).  Thread (y, tw) converts 8 pixels into three 16-byte chunks as before, but drops them into a shared-memory image of
// the output (patch pitch padded to 400 B: conflict-free 16-byte stores); the block then streams that image out with
// consecutive threads writing consecutive 16-byte chunks.  The first version wrote its chunks straight to global memory,
// 16 bytes every 384: half-filled sectors, 2.6 TB/s (40 % of the copy peak, profiles/r01_membound.txt).
constexpr int PU8_PITCH = 400;                 // bytes per patch in the staging image (384 + 16)

__global__ void __launch_bounds__(256)
patchify8_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  extern __shared__ __align__(16) uint8_t stage[];          // [gw][PU8_PITCH]
  const int gw = W >> 3, gh = H >> 3;
  const int work = gw * 8;                                   // (y, tw) pairs of one patch row
  const int chunks = gw * 24;                                // 16-byte chunks of its output
  const long long n_rows = static_cast<long long>(B) * gh;
  // (y, tw) of this thread within a patch row; the loads of the NEXT patch row are issued before this one is streamed out
  const bool active = static_cast<int>(threadIdx.x) < work;       // gw <= 32 (W <= 256): one (y, tw) pair per thread
  const int tw0 = threadIdx.x % gw, kh0 = threadIdx.x / gw;
  auto load = [&](long long pr, uint2& w0, uint2& w1, uint2& w2) {
    const int th = static_cast<int>(pr % gh);
    const int b = static_cast<int>(pr / gh);
    const uint2* src = reinterpret_cast<const uint2*>(img + ((static_cast<size_t>(b) * H + th * 8 + kh0) * W + tw0 * 8) * 3);
    w0 = __ldcs(src); w1 = __ldcs(src + 1); w2 = __ldcs(src + 2);
  };
  uint2 w0 = make_uint2(0u, 0u), w1 = w0, w2 = w0;
  long long pr = blockIdx.x;
  if (pr < n_rows && active) load(pr, w0, w1, w2);
  for (; pr < n_rows; pr += gridDim.x) {
    if (active) {
      const uint32_t wd[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
      float v[3][8];
#pragma unroll
      for (int i = 0; i < 24; ++i) v[i % 3][i / 3] = norm_u8_fast(wd[i >> 2], i & 3);
      uint8_t* dst = stage + tw0 * PU8_PITCH + kh0 * 16;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint4 o;
        o.x = pack_bf16x2(v[c][0], v[c][1]); o.y = pack_bf16x2(v[c][2], v[c][3]);
        o.z = pack_bf16x2(v[c][4], v[c][5]); o.w = pack_bf16x2(v[c][6], v[c][7]);
        *reinterpret_cast<uint4*>(dst + c * 128) = o;
      }
    }
    __syncthreads();
    if (pr + gridDim.x < n_rows && active) load(pr + gridDim.x, w0, w1, w2);
    uint4* gdst = reinterpret_cast<uint4*>(out + static_cast<size_t>(pr) * gw * 192);
    for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
      const int tw = i / 24, ch = i - tw * 24;
      __stcs(gdst + i, *reinterpret_cast<const uint4*>(stage + tw * PU8_PITCH + ch * 16));
    }
    __syncthreads();
  }
}

// fp32 NCHW ingest through the same shared-memory output image (C = 3, W <= 256): the direct kernel above stores 16 bytes every
// 384 (half-filled sectors: 78 % of the copy peak).  One block iteration = one row of patches: 24 gw items (c, kh, tw) of 8 floats,
// up to three per thread; the next row's loads are in flight while this one is streamed out.
__global__ void __launch_bounds__(256)
patchify8_stage_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  extern __shared__ __align__(16) uint8_t stage[];          // [gw][PU8_PITCH]
  const int gw = W >> 3, gh = H >> 3;
  const int items = gw * 24;                                 // (c, kh, tw) pieces of one patch row == 16-byte chunks of its output
  const long long n_rows = static_cast<long long>(B) * gh;
  int soff[3];                                               // staging offset of this thread's items (-1: none)
  size_t goff[3];                                            // source offset within the image, relative to the patch row's first line
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int j = threadIdx.x + 256 * r;
    const int tw = j % gw, rest = j / gw;
    const int kh = rest & 7, c = rest >> 3;
    soff[r] = j < items ? tw * PU8_PITCH + c * 128 + kh * 16 : -1;
    goff[r] = (static_cast<size_t>(c) * H + kh) * W + tw * 8;
  }
  float4 a[3][2];
  auto load = [&](long long pr) {
    const int th = static_cast<int>(pr % gh);
    const int b = static_cast<int>(pr / gh);
    const float* base = img + (static_cast<size_t>(b) * 3 * H + th * 8) * W;
#pragma unroll
    for (int r = 0; r < 3; ++r)
      if (soff[r] >= 0) {
        a[r][0] = __ldcs(reinterpret_cast<const float4*>(base + goff[r]));
        a[r][1] = __ldcs(reinterpret_cast<const float4*>(base + goff[r] + 4));
      }
  };
  long long pr = blockIdx.x;
  if (pr < n_rows) load(pr);
  for (; pr < n_rows; pr += gridDim.x) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
      if (soff[r] >= 0) {
        uint4 o;
        o.x = pack_bf16x2(a[r][0].x, a[r][0].y); o.y = pack_bf16x2(a[r][0].z, a[r][0].w);
        o.z = pack_bf16x2(a[r][1].x, a[r][1].y); o.w = pack_bf16x2(a[r][1].z, a[r][1].w);
        *reinterpret_cast<uint4*>(stage + soff[r]) = o;
      }
    __syncthreads();
    if (pr + gridDim.x < n_rows) load(pr + gridDim.x);
    uint4* gdst = reinterpret_cast<uint4*>(out + static_cast<size_t>(pr) * gw * 192);
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
      const int tw = i / 24, ch = i - tw * 24;
      gdst[i] = *reinterpret_cast<const uint4*>(stage + tw * PU8_PITCH + ch * 16);
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------
// LayerNorm over rows of a bf16 [M, D] matrix (D % 8 == 0, D <= 2048), eps inside the sqrt,
// biased variance — nn.LayerNorm semantics (stage1/layers.py:49,51,89,128).
//   MODE 0: stats only  -> stats[row] = (mean, rstd)   (consumed by the LN-folded GEMM epilogue)
//   MODE 1: y = (x - mean) * rstd * gamma + beta  (bf16), optionally also the stats of y
// ----------------------------------------------------------------------------------------------
constexpr int LN_MAX_VEC = 8;   // 8 x (32 lanes x 8 elts) = 2048 columns

// NV = 16-byte vectors per lane actually needed (D <= NV * 256): the row lives in NV * 8 registers per lane, so
// D = 512 compiles to a 2-vector kernel with ~4x the occupancy of the generic 8-vector one (measured: 1.9 -> TB/s).
// Round 2: every warp handles TWO rows at once (2 NV 16-byte loads in flight per lane instead of NV): at D = 512 one row per
// warp left the kernel at 4.0 TB/s, latency-bound on 32 bytes in flight per lane (profiles/r01_membound.txt).
// Round 2, second pass: ncu showed the two-rows-per-warp kernel ISSUE-bound, not latency-bound (issue slots 76 % busy, 371 warp
// instructions per 512-wide row, 1.2 M local loads: 64 registers spilled; four rows per warp changed nothing).  Now: packed
// f32x2 arithmetic in every pass (sum, variance, affine output, statistics of the rounded output: ~7 instead of ~14 instructions
// per element) and three instead of four resident blocks so that the fp32 row stays in registers without spills.
__device__ __forceinline__ void ln_unpack4(const uint4& u, float2 (&f)[4]) {
  f[0] = make_float2(bf16lo_to_f32(u.x), bf16hi_to_f32(u.x));
  f[1] = make_float2(bf16lo_to_f32(u.y), bf16hi_to_f32(u.y));
  f[2] = make_float2(bf16lo_to_f32(u.z), bf16hi_to_f32(u.z));
  f[3] = make_float2(bf16lo_to_f32(u.w), bf16hi_to_f32(u.w));
}

template <int MODE, int NV>
__global__ void __launch_bounds__(256, NV <= 2 ? 3 : 1)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int M, int D, float eps,
                 const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y, int64_t ldy, float* __restrict__ stats) {
  constexpr int R = 2;                           // rows per warp
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = R * warp;
  if (row0 >= M) return;
  const int nrows = M - row0 < R ? M - row0 : R;
  const int nvec = D >> 3;                      // 16-byte vectors per row
  float2 v[R][NV][4];
  {
    uint4 u[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = i * 32 + lane;
        u[r][i] = make_uint4(0u, 0u, 0u, 0u);
        if (vi < nvec && r < nrows) u[r][i] = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(row0 + r) * ldx + vi * 8);
      }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < NV; ++i) ln_unpack4(u[r][i], v[r][i]);
  }
  float mean[R], rstd[R];
  {
    float2 s2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s2[r] = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) s2[r] = __fadd2_rn(s2[r], v[r][i][k]);       // lanes beyond the row hold zeros
    }
    float sum[R];
#pragma unroll
    for (int r = 0; r < R; ++r) sum[r] = s2[r].x + s2[r].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
    float sq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] = sum[r] / static_cast<float>(D);
      const float2 nm = make_float2(-mean[r], -mean[r]);
      float2 q2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (i * 32 + lane < nvec) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 d = __fadd2_rn(v[r][i][k], nm);
            q2 = __ffma2_rn(d, d, q2);
          }
        }
      }
      sq[r] = q2.x + q2.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], o);
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(sq[r] / static_cast<float>(D) + eps);
  }
  if (MODE == 0) {
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r < nrows) *reinterpret_cast<float2*>(stats + 2 * static_cast<size_t>(row0 + r)) = make_float2(mean[r], rstd[r]);
    }
    return;
  }
  float2 ys2[R];
#pragma unroll
  for (int r = 0; r < R; ++r) ys2[r] = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = i * 32 + lane;
    if (vi < nvec) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + vi * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + vi * 8 + 4);
      const float2 g[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
      const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float2 nm = make_float2(-mean[r], -mean[r]);
        const float2 rs = make_float2(rstd[r], rstd[r]);
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // ((x - mean) * rstd) * gamma + beta, each step rounded as before
          const float2 a = __ffma2_rn(__fmul2_rn(__fadd2_rn(v[r][i][k], nm), rs), g[k], bb[k]);
          pk[k] = pack_bf16x2(a.x, a.y);
          // statistics of the ROUNDED output (what the next GEMM will actually read); v is reused to hold it
          v[r][i][k] = make_float2(bf16lo_to_f32(pk[k]), bf16hi_to_f32(pk[k]));
          ys2[r] = __fadd2_rn(ys2[r], v[r][i][k]);
        }
        if (r < nrows) *reinterpret_cast<uint4*>(y + static_cast<size_t>(row0 + r) * ldy + vi * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  if (stats != nullptr) {
    float ysum[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ysum[r] = ys2[r].x + ys2[r].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) ysum[r] += __shfl_xor_sync(0xffffffffu, ysum[r], o);
    float ysq[R], ymean[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ymean[r] = ysum[r] / static_cast<float>(D);
      const float2 nm = make_float2(-ymean[r], -ymean[r]);
      float2 q2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (i * 32 + lane < nvec) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 d = __fadd2_rn(v[r][i][k], nm);
            q2 = __ffma2_rn(d, d, q2);
          }
        }
      }
      ysq[r] = q2.x + q2.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) ysq[r] += __shfl_xor_sync(0xffffffffu, ysq[r], o);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r < nrows)
          *reinterpret_cast<float2*>(stats + 2 * static_cast<size_t>(row0 + r)) = make_float2(ymean[r], rsqrtf(ysq[r] / static_cast<float>(D) + eps));
    }
  }
}

// Fast paths for D = 256 / 512 / 1024 (compile-time D, no tail guards, divisions by D fold into exact multiplications).
// Tried first: HALF a warp per row (the two rows of a warp share every shuffle; 330 executed instructions per two rows against 740
// for the generic kernel) — 0.117 ms one-shot, 0.111 ms persistent with prefetch: no better than the generic 0.119-0.121 ms.
// ncu on the variants above: L1/TEX throughput 94 % — every output element drags 8 bytes of gamma / beta through the L1 data
// stage next to its 2 + 2 bytes of row data.  This kernel is PERSISTENT (a fixed grid walks the rows) so that a lane keeps the
// gamma / beta of ITS columns in registers for the whole launch, and it requests the next rows' vectors before it works on the
// current ones.  One warp per row (R rows per trip), compile-time D.
template <int MODE, int D>
__global__ void __launch_bounds__(256, 2)
layernorm_reg_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int M, float eps,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y, int64_t ldy, float* __restrict__ stats) {
  constexpr int NV = D / 256;                    // 16-byte vectors per lane
  constexpr int R = D <= 512 ? 2 : 1;            // rows per trip
  constexpr float inv_d = 1.0f / static_cast<float>(D);      // exact: D is a power of two
  const int lane = threadIdx.x & 31;
  const int nwarps = static_cast<int>(gridDim.x * blockDim.x) >> 5;
  int row0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R;
  if (row0 >= M) return;
  float2 g[NV][4], bb[NV][4];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      g[i][0] = make_float2(g0.x, g0.y); g[i][1] = make_float2(g0.z, g0.w); g[i][2] = make_float2(g1.x, g1.y); g[i][3] = make_float2(g1.z, g1.w);
      bb[i][0] = make_float2(b0.x, b0.y); bb[i][1] = make_float2(b0.z, b0.w); bb[i][2] = make_float2(b1.x, b1.y); bb[i][3] = make_float2(b1.z, b1.w);
    }
  }
  auto load_rows = [&](uint4 (&u)[R][NV], int r0) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        u[r][i] = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + r < M) u[r][i] = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r0 + r) * ldx + (i * 32 + lane) * 8);
      }
  };
  uint4 u[R][NV];
  load_rows(u, row0);
  for (; row0 < M; row0 += nwarps * R) {
    float2 v[R][NV][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < NV; ++i) ln_unpack4(u[r][i], v[r][i]);
    if (row0 + nwarps * R < M) load_rows(u, row0 + nwarps * R);       // next trip's rows: in flight under this trip's arithmetic
    float mean[R], rstd[R];
    {
      float sum[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float2 s2 = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) s2 = __fadd2_rn(s2, v[r][i][k]);
        sum[r] = s2.x + s2.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < R; ++r) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
      float sq[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        mean[r] = sum[r] * inv_d;
        const float2 nm = make_float2(-mean[r], -mean[r]);
        float2 q2 = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 d = __fadd2_rn(v[r][i][k], nm);
            q2 = __ffma2_rn(d, d, q2);
          }
        sq[r] = q2.x + q2.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < R; ++r) sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], o);
#pragma unroll
      for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(sq[r] * inv_d + eps);
    }
    if (MODE == 0) {
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (row0 + r < M) *reinterpret_cast<float2*>(stats + 2 * static_cast<size_t>(row0 + r)) = make_float2(mean[r], rstd[r]);
      }
      continue;
    }
    float ysum[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float2 nm = make_float2(-mean[r], -mean[r]);
      const float2 rs = make_float2(rstd[r], rstd[r]);
      float2 ys2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = __ffma2_rn(__fmul2_rn(__fadd2_rn(v[r][i][k], nm), rs), g[i][k], bb[i][k]);
          pk[k] = pack_bf16x2(a.x, a.y);
          v[r][i][k] = make_float2(bf16lo_to_f32(pk[k]), bf16hi_to_f32(pk[k]));   // the ROUNDED output: what the next GEMM reads
          ys2 = __fadd2_rn(ys2, v[r][i][k]);
        }
        if (row0 + r < M)
          *reinterpret_cast<uint4*>(y + static_cast<size_t>(row0 + r) * ldy + (i * 32 + lane) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      ysum[r] = ys2.x + ys2.y;
    }
    if (stats != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < R; ++r) ysum[r] += __shfl_xor_sync(0xffffffffu, ysum[r], o);
      float ysq[R], ymean[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        ymean[r] = ysum[r] * inv_d;
        const float2 nm = make_float2(-ymean[r], -ymean[r]);
        float2 q2 = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 d = __fadd2_rn(v[r][i][k], nm);
            q2 = __ffma2_rn(d, d, q2);
          }
        ysq[r] = q2.x + q2.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < R; ++r) ysq[r] += __shfl_xor_sync(0xffffffffu, ysq[r], o);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (row0 + r < M)
            *reinterpret_cast<float2*>(stats + 2 * static_cast<size_t>(row0 + r)) = make_float2(ymean[r], rsqrtf(ysq[r] * inv_d + eps));
      }
    }
  }
}

// fp32 -> bf16 (round to nearest even), n % 8 == 0, 16-byte aligned; used for the text context
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
  const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
  reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}

int pm_cast_launch(const float* src, void* dst, long long n, cudaStream_t stream) {
  if (src == nullptr || dst == nullptr || n <= 0 || (n & 7) != 0) return PM_ERR_INVALID;
  const long long n8 = n >> 3;
  cast_f32_bf16_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n8);
  return static_cast<int>(cudaGetLastError());
}

int pm_patchify_launch(const float* img, void* out, int B, int C, int H, int W, int P, cudaStream_t stream) {
  if (img == nullptr || out == nullptr || P != 8 || (H % 8) != 0 || (W % 8) != 0 || B <= 0 || C <= 0) return PM_ERR_INVALID;
  static int direct = -1;                                         // PM_PATCHIFY_DIRECT=1: the direct kernel for every shape (A/B aid)
  if (direct < 0) direct = getenv("PM_PATCHIFY_DIRECT") != nullptr ? 1 : 0;
  if (!direct && C == 3 && W <= 256) {
    long long nb = static_cast<long long>(B) * (H / 8);           // one patch row per block iteration
    const long long cap = static_cast<long long>(pm_num_sms()) * 8;
    if (nb > cap) nb = cap;
    patchify8_stage_kernel<<<static_cast<unsigned>(nb), 256, (W / 8) * PU8_PITCH, stream>>>(img, reinterpret_cast<__nv_bfloat16*>(out), B, H, W);
    return static_cast<int>(cudaGetLastError());
  }
  const long long total = static_cast<long long>(B) * C * H * (W / 8);
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  patchify8_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(img, reinterpret_cast<__nv_bfloat16*>(out), B, C, H, W);
  return static_cast<int>(cudaGetLastError());
}

int pm_patchify_u8_launch(const uint8_t* img, void* out, int B, int H, int W, cudaStream_t stream) {
  if (img == nullptr || out == nullptr || (H % 8) != 0 || (W % 8) != 0 || B <= 0) return PM_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(img) & 7) != 0) return PM_ERR_INVALID;
  const int threads = 256;                                        // == table size
  long long blocks = static_cast<long long>(B) * (H / 8);         // one patch row per block iteration
  const long long cap = static_cast<long long>(pm_num_sms()) * 8; // grid-stride: amortise the table build
  if (blocks > cap) blocks = cap;
  const int smem = (W / 8) * PU8_PITCH;
  if (W > 256) return PM_ERR_INVALID;                             // one (image row, patch) pair per thread of a 256-thread block
  patchify8_u8_kernel<<<static_cast<unsigned>(blocks), threads, smem, stream>>>(img, reinterpret_cast<__nv_bfloat16*>(out), B, H, W);
  return static_cast<int>(cudaGetLastError());
}

template <int NV>
static void launch_ln(const void* x, int64_t ldx, int M, int D, float eps, const float* gamma, const float* beta, void* y,
                      int64_t ldy, float* stats, cudaStream_t stream) {
  const int threads = 256;                       // 8 warps x 2 rows per block
  const int blocks = (M + 15) / 16;
  if (y == nullptr)
    layernorm_kernel<0, NV><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, D, eps,
                                                            nullptr, nullptr, nullptr, 0, stats);
  else
    layernorm_kernel<1, NV><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, D, eps,
                                                            gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), ldy, stats);
}

template <int D>
static void launch_ln_reg(const void* x, int64_t ldx, int M, float eps, const float* gamma, const float* beta, void* y, int64_t ldy,
                          float* stats, cudaStream_t stream) {
  constexpr int R = D <= 512 ? 2 : 1;
  const int want = (M + 8 * R - 1) / (8 * R);    // 8 warps x R rows per block and trip
  static int per_sm = -1;                        // PM_LN_BLOCKS=<blocks per SM> (tuning aid)
  if (per_sm < 0) {
    const char* env = getenv("PM_LN_BLOCKS");
    per_sm = env != nullptr ? atoi(env) : 2;
  }
  const int cap = pm_num_sms() * per_sm;
  const int blocks = want < cap ? want : cap;
  if (y == nullptr)
    layernorm_reg_kernel<0, D><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, eps, nullptr, nullptr,
                                                           nullptr, 0, stats);
  else
    layernorm_reg_kernel<1, D><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, eps, gamma, beta,
                                                           reinterpret_cast<__nv_bfloat16*>(y), ldy, stats);
}

int pm_layernorm_launch(const void* x, int64_t ldx, int M, int D, float eps, const float* gamma,
                        const float* beta, void* y, int64_t ldy, float* stats, cudaStream_t stream) {
  if (x == nullptr || M <= 0 || D <= 0 || (D % 8) != 0 || D > LN_MAX_VEC * 256 || (ldx % 8) != 0) return PM_ERR_INVALID;
  if (y == nullptr && stats == nullptr) return PM_ERR_INVALID;
  if (y != nullptr && (gamma == nullptr || beta == nullptr || (ldy % 8) != 0)) return PM_ERR_INVALID;
  static int generic = -1;                       // PM_LN_GENERIC=1: the generic kernel for every D (A/B aid)
  if (generic < 0) generic = getenv("PM_LN_GENERIC") != nullptr ? 1 : 0;
  if (!generic && (D == 256 || D == 512 || D == 1024)) {
    if (D == 256) launch_ln_reg<256>(x, ldx, M, eps, gamma, beta, y, ldy, stats, stream);
    else if (D == 512) launch_ln_reg<512>(x, ldx, M, eps, gamma, beta, y, ldy, stats, stream);
    else launch_ln_reg<1024>(x, ldx, M, eps, gamma, beta, y, ldy, stats, stream);
    return static_cast<int>(cudaGetLastError());
  }
  if (D <= 256) launch_ln<1>(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, stream);
  else if (D <= 512) launch_ln<2>(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, stream);
  else if (D <= 1024) launch_ln<4>(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, stream);
  else launch_ln<LN_MAX_VEC>(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, stream);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
