// pm_kernels.h — internal (C++) declarations shared by the .cu files of libpaintmind_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

enum : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_UNPATCH = 2, OUT_UNPATCH_U8 = 3 };

struct GemmParams {
  const void* a;
  const void* w;
  void* out;
  const float* bias;
  const float* colsum;
  const float* stats;
  const float* pos;
  const void* res;
  int64_t lda, ldw, ld_out, ld_pos, ld_res;
  int M, N, K;
  int N_out;          // output columns (N, or N/2 with SwiGLU); set by the launcher
  int pos_rows;
  int patch, channels, grid;
  int max_ctas;
  float* stats_out;   // [M, 2] += (sum x, sum x^2) of the OUTPUT rows (OUT_BF16, non-SwiGLU), or nullptr
  int stats_raw;      // 1: `stats` holds raw (sum, sum of squares) over K columns; 0: (mean, rstd)
  float ln_eps;
  long long* debug;   // optional [grid, 4] int64: MMA-issuer stall cycles (accumulator wait, operand wait, total, k_blocks)
  int res_mod;        // > 0: `res` has res_mod rows and row r of the output adds res[r % res_mod] (a broadcast table such as
                      // the position embedding, staged by TMA like a residual); multiple of 128
  int cta_pair;       // 1: use the cta_group::2 kernel (256-row tiles) when the tile shape allows it
};

struct AttnParams {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  int64_t ldq, ldk, ldv, ldo;      // row pitch in elements
  int64_t bsq, bsk, bsv, bso;      // batch stride in elements
  int B, H, Nq, Nk, head_dim;
  float scale_log2;                // softmax scale * log2(e)
  float* lse;                      // optional [B, H, Nq] fp32 out: base-2 log-sum-exp of the scaled scores (training forward)
  int64_t lse_ld;                  // row pitch of lse per (batch, head): >= Nq (pm_attn_bwd wants it rounded up to 128)
  float* o32;                      // optional [B, Nq, ldo32] fp32 copy of O (training forward; feeds delta of the backward)
  int64_t ldo32;
  int prescaled;                   // Q carries scale * log2(e): scale_log2 is 1 (pm_attn3.cu)
};

struct AttnBwdParams {
  const void* q;
  const void* k;
  const void* v;
  const void* dO;
  void* dq;
  void* dk;
  void* dv;
  const float* nlse;               // [B, H, lse_ld]: -lse (lse from the forward, base 2)
  const float* nds;                // [B, H, lse_ld]: -scale * rowsum(dO * O)
  int64_t lse_ld;                  // >= Nq rounded up to 128
  long long* debug;                // optional [#SMs, 8] int64 cycle counters of the last launch (profiling aid), or nullptr
  int64_t ldq, ldk, ldv, lddo, lddq, lddk, lddv;
  int64_t bsq, bsk, bsv, bsdo, bsdq, bsdk, bsdv;
  int B, H, Nq, Nk, head_dim;
  float scale, scale_log2;
};

struct WgradParams {
  const void* dy;                  // bf16 [M, N] (row pitch lddy >= N rounded up to 64)
  const void* x;                   // bf16 [M, K]
  float* work;                     // fp32 [splits, N, K]
  float* out;                      // fp32 [N, K], row pitch ld_out
  int64_t lddy, ldx, ld_out;
  int M, N, K;
  int splits;                      // 0 = auto (pm_wgrad_splits)
  int accumulate;                  // 1: out += result
  float scale;
};

struct VqParams {
  const float* z;                  // [M, 32] raw latents (row pitch ldz)
  int64_t ldz;
  int M;
  const float* en;                 // [n_e, 32] normalised codebook (fp32)
  const void* packed;              // [n_e, 64] bf16 = [e_hi | e_lo]
  int n_e, e_dim;
  int splits;                      // 0 = auto
  float* cand_val;                 // [splits, M] scratch (splits > 1)
  int* cand_idx;
  long long* idx;                  // [M] int64 out
  float* zq;                       // [M, 32] fp32 out (straight-through forward value)
  void* zq_split;                  // [M, 64] bf16 out = [hi | lo] (decoder GEMM operand), optional
  double* sse;                     // += sum (z_q - zn)^2
  unsigned long long* hist;        // [n_e] += usage counts, optional
};

// per-step scalars of the MaskGIT loop in device memory (== pm_step_scalars of the C-ABI header)
struct StepScalars {
  float temperature;
  int k;
  unsigned long long seed, offset;
};

struct MaskgitParams {
  const float* logits;             // [M, V] fp32, row pitch ld
  int64_t ld;
  int M, V, topk;
  float temperature;
  const float* noise;              // optional injected uniforms [M, V] (row pitch ld_noise)
  int64_t ld_noise;
  unsigned long long seed, offset; // Philox key / stream offset when noise == nullptr
  long long* ids;                  // [M] in/out (masked positions filled with the prediction), optional
  long long* pred_ids;             // [M] out
  float* scores;                   // [M] out: 1 - p(pred) at masked positions, -1e5 elsewhere
  long long mask_id;
  const StepScalars* step_tab;     // optional: temperature / seed / offset are read from step_tab[*step_idx] (CUDA-graph replay)
  const int* step_idx;
};

int pm_num_sms();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel, device).  `done` is a
// call-site-owned `static bool[PM_MAX_DEVICES]` (zero-initialised).
constexpr int PM_MAX_DEVICES = 64;
template <typename Kern>
inline int pm_ensure_dyn_smem(Kern kern, int bytes, bool* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (dev < 0 || dev >= PM_MAX_DEVICES) dev = PM_MAX_DEVICES - 1, done[dev] = false;
  if (!done[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    done[dev] = true;
  }
  return 0;
}
int pm_cast_launch(const float* src, void* dst, long long n, cudaStream_t stream);
int pm_maskgit_sample_launch(const MaskgitParams& p, cudaStream_t stream);
int pm_maskgit_remask_launch(const float* scores, long long* ids, int B, int N, int k, long long mask_id, const StepScalars* step_tab,
                             int* step_idx, int* ticket, cudaStream_t stream);
int pm_maskgit_random_mask_launch(const float* z, int64_t ldz, const float* noise, unsigned long long seed,
                                  unsigned long long offset, const float* mask_token, int B, int N, int len_keep,
                                  float* mask, float* x_out, cudaStream_t stream);
int pm_ce_label_smooth_launch(const float* logits, int64_t ld, int M, int V, const long long* label, const float* mask,
                              float eps, float* row_loss, float* loss_out, double* sums_out, cudaStream_t stream);
int pm_attn_launch(const AttnParams& p, cudaStream_t stream);
int pm_attn2_launch(const AttnParams& p, cudaStream_t stream);
int pm_attn3_launch(const AttnParams& p, cudaStream_t stream);
bool pm_attn3_supported(const AttnParams& p);
int pm_attn4_launch(const AttnParams& p, cudaStream_t stream);
bool pm_attn4_supported(const AttnParams& p);
int pm_attn_bwd_launch(const AttnBwdParams& p, cudaStream_t stream);
int pm_attn_delta_launch(const void* o, int o_is_f32, int64_t ldo, int64_t bso, const void* dO, int64_t lddo, int64_t bsdo, int B, int H,
                         int N, const float* lse, float scale, float* nds, float* nlse, int64_t delta_ld, cudaStream_t stream);
int pm_wgrad_splits(int M, int N, int K);
int pm_wgrad_launch(const WgradParams& p, cudaStream_t stream);
int pm_colsum_rows(int M, int N);
int pm_colreduce_launch(const float* partial, int R, int N, float* out, int accumulate, cudaStream_t stream);
int pm_colsum_launch(const void* x, int64_t ld, int M, int N, float* partial, float* out, int accumulate, cudaStream_t stream);
int pm_ln_bwd_blocks(int M);
int pm_ln_bwd_launch(const void* dn, int64_t lddn, const void* x, int64_t ldx, const float* gamma, const void* dres,
                     int64_t ldres, void* dx, int64_t lddx, int M, int D, float eps, float* part, float* dgamma_dbeta,
                     cudaStream_t stream);
int pm_swiglu_bwd_chunks(int M, int hp);
int pm_swiglu_bwd_launch(const void* x12, int64_t ld12, const void* dh, int64_t lddh, void* h, int64_t ldh, void* d12,
                         int64_t ldd12, int M, int hp, float* work, float* b12, cudaStream_t stream);
int pm_vq_bwd_launch(const float* z, int64_t ldz, const long long* idx, const float* E, const float* d_out, int64_t ldd,
                     const float* d_loss, float beta, int M, float* dz, void* dz_split, float* dE, cudaStream_t stream);
int pm_unpatchify_bwd_launch(const float* g, const float* rec, void* out, int B, int C, int H, int W, cudaStream_t stream);
int pm_vq_codebook_prep_launch(const float* E, int n_e, float* en, void* packed, cudaStream_t stream);
int pm_vq_launch(const VqParams& p, cudaStream_t stream);
int pm_vq_gather_launch(const long long* idx, int M, int n_rows, const float* table, int normalize,
                        float* out, void* out_split, cudaStream_t stream);
int pm_split_rows32_launch(const float* src, int64_t ld, int M, void* out_split, cudaStream_t stream);
int pm_patchify_u8_launch(const uint8_t* img, void* out, int B, int H, int W, cudaStream_t stream);
int pm_patchify_launch(const float* img, void* out, int B, int C, int H, int W, int P, cudaStream_t stream);
int pm_layernorm_launch(const void* x, int64_t ldx, int M, int D, float eps, const float* gamma,
                        const float* beta, void* y, int64_t ldy, float* stats, cudaStream_t stream);
int pm_gemm_launch(const GemmParams& p, int bn, int out_mode, int swiglu, cudaStream_t stream);

}  // namespace pm
