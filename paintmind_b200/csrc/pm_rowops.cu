// pm_rowops.cu — HBM-bound row kernels of the tokenizer path (sm_100a): patch extraction,
// LayerNorm statistics / LayerNorm.  One warp per row, 16-byte vector accesses, fp32 math.
#include "pm_common.cuh"
#include "pm_kernels.h"

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

// The 256 possible results are tabulated once per block (bf16 bits in shared memory): the two IEEE divisions per
// byte made the first version compute-bound (2.2 TB/s); with the table the kernel is a byte shuffle.
__global__ void __launch_bounds__(256)
patchify8_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  __shared__ uint16_t lut[256];
  {
    const __nv_bfloat16 h = __float2bfloat16_rn(norm_u8(threadIdx.x));
    lut[threadIdx.x] = *reinterpret_cast<const uint16_t*>(&h);
  }
  __syncthreads();
  const int gw = W >> 3, gh = H >> 3;
  const long long total = static_cast<long long>(B) * H * gw;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int tw = static_cast<int>(t % gw);
    const long long r = t / gw;
    const int y = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    const uint2* src = reinterpret_cast<const uint2*>(img + ((static_cast<size_t>(b) * H + y) * W + tw * 8) * 3);
    const uint2 w0 = __ldcs(src), w1 = __ldcs(src + 1), w2 = __ldcs(src + 2);
    const uint32_t wd[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
    uint32_t v[3][8];
#pragma unroll
    for (int i = 0; i < 24; ++i) v[i % 3][i / 3] = lut[(wd[i >> 2] >> ((i & 3) * 8)) & 0xffu];
    const int th = y >> 3, kh = y & 7;
    const size_t row = (static_cast<size_t>(b) * gh + th) * gw + tw;
    __nv_bfloat16* dst = out + row * 192 + kh * 8;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint4 o;
      o.x = v[c][0] | (v[c][1] << 16); o.y = v[c][2] | (v[c][3] << 16);
      o.z = v[c][4] | (v[c][5] << 16); o.w = v[c][6] | (v[c][7] << 16);
      *reinterpret_cast<uint4*>(dst + c * 64) = o;
    }
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
template <int MODE, int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int M, int D, float eps,
                 const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y, int64_t ldy, float* __restrict__ stats) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const __nv_bfloat16* xr = x + static_cast<size_t>(warp) * ldx;
  const int nvec = D >> 3;                      // 16-byte vectors per row
  float v[NV][8];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = i * 32 + lane;
    if (vi < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + vi * 8);
      v[i][0] = bf16lo_to_f32(u.x); v[i][1] = bf16hi_to_f32(u.x);
      v[i][2] = bf16lo_to_f32(u.y); v[i][3] = bf16hi_to_f32(u.y);
      v[i][4] = bf16lo_to_f32(u.z); v[i][5] = bf16hi_to_f32(u.z);
      v[i][6] = bf16lo_to_f32(u.w); v[i][7] = bf16hi_to_f32(u.w);
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += v[i][k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(D);
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = i * 32 + lane;
    if (vi < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = v[i][k] - mean;
        sq = fmaf(d, d, sq);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(D) + eps);
  if (MODE == 0) {
    if (lane == 0) *reinterpret_cast<float2*>(stats + 2 * static_cast<size_t>(warp)) = make_float2(mean, rstd);
    return;
  }
  __nv_bfloat16* yr = y + static_cast<size_t>(warp) * ldy;
  float ysum = 0.0f;
  float yv[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = i * 32 + lane;
    if (vi < nvec) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + vi * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + vi * 8 + 4);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t pk[4];
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const float a0 = (v[i][k] - mean) * rstd * g[k] + bb[k];
        const float a1 = (v[i][k + 1] - mean) * rstd * g[k + 1] + bb[k + 1];
        pk[k >> 1] = pack_bf16x2(a0, a1);
        // statistics of the ROUNDED output (what the next GEMM will actually read)
        yv[i][k] = bf16lo_to_f32(pk[k >> 1]);
        yv[i][k + 1] = bf16hi_to_f32(pk[k >> 1]);
        ysum += yv[i][k] + yv[i][k + 1];
      }
      *reinterpret_cast<uint4*>(yr + vi * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ysum += __shfl_xor_sync(0xffffffffu, ysum, o);
    const float ymean = ysum / static_cast<float>(D);
    float ysq = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = i * 32 + lane;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = yv[i][k] - ymean;
          ysq = fmaf(d, d, ysq);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ysq += __shfl_xor_sync(0xffffffffu, ysq, o);
    if (lane == 0)
      *reinterpret_cast<float2*>(stats + 2 * static_cast<size_t>(warp)) = make_float2(ymean, rsqrtf(ysq / static_cast<float>(D) + eps));
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
  const long long total = static_cast<long long>(B) * C * H * (W / 8);
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  patchify8_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(img, reinterpret_cast<__nv_bfloat16*>(out), B, C, H, W);
  return static_cast<int>(cudaGetLastError());
}

int pm_patchify_u8_launch(const uint8_t* img, void* out, int B, int H, int W, cudaStream_t stream) {
  if (img == nullptr || out == nullptr || (H % 8) != 0 || (W % 8) != 0 || B <= 0) return PM_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(img) & 7) != 0) return PM_ERR_INVALID;
  const long long total = static_cast<long long>(B) * H * (W / 8);
  const int threads = 256;                                        // == table size
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(pm_num_sms()) * 32;   // grid-stride: amortise the table build
  if (blocks > cap) blocks = cap;
  patchify8_u8_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(img, reinterpret_cast<__nv_bfloat16*>(out), B, H, W);
  return static_cast<int>(cudaGetLastError());
}

template <int NV>
static void launch_ln(const void* x, int64_t ldx, int M, int D, float eps, const float* gamma, const float* beta, void* y,
                      int64_t ldy, float* stats, cudaStream_t stream) {
  const int threads = 256;                       // 8 rows per block
  const int blocks = (M + 7) / 8;
  if (y == nullptr)
    layernorm_kernel<0, NV><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, D, eps,
                                                            nullptr, nullptr, nullptr, 0, stats);
  else
    layernorm_kernel<1, NV><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, D, eps,
                                                            gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), ldy, stats);
}

int pm_layernorm_launch(const void* x, int64_t ldx, int M, int D, float eps, const float* gamma,
                        const float* beta, void* y, int64_t ldy, float* stats, cudaStream_t stream) {
  if (x == nullptr || M <= 0 || D <= 0 || (D % 8) != 0 || D > LN_MAX_VEC * 256 || (ldx % 8) != 0) return PM_ERR_INVALID;
  if (y == nullptr && stats == nullptr) return PM_ERR_INVALID;
  if (y != nullptr && (gamma == nullptr || beta == nullptr || (ldy % 8) != 0)) return PM_ERR_INVALID;
  if (D <= 256) launch_ln<1>(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, stream);
  else if (D <= 512) launch_ln<2>(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, stream);
  else if (D <= 1024) launch_ln<4>(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, stream);
  else launch_ln<LN_MAX_VEC>(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, stream);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pm
