/* paintmind_b200.h — C-ABI of libpaintmind_b200.so (sm_100a kernels of the PaintMind tokenizer hot path).
 *
 * The reference (Qiyuan-Ge/PaintMind) is pure Python: its only plug-in seam for this path is the
 * nn.Module surface.  There is therefore no reference FFI to mirror; each entry
 * point below names the reference Python call site(s) it replaces.  The Python host side
 * (paintmind_b200/*.py) keeps the reference's module / state_dict surface and reaches these
 * functions through ctypes (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch allocator); nothing is
 *     allocated, retained or freed by the library;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync;
 *   - return value 0 = ok, <0 = library error (PM_ERR_*), >0 = cudaError_t;
 *   - bf16 tensors are row-major with explicit leading dimensions (in elements); TMA requires
 *     16-byte aligned bases and leading dimensions that are multiples of 8 elements;
 *   - thread-safe per stream; no global state besides cached function attributes.
 */
#ifndef PAINTMIND_B200_H_
#define PAINTMIND_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_OK 0
#define PM_ERR_INVALID (-1)
#define PM_ERR_TENSORMAP (-2)
#define PM_ERR_NO_DRIVER (-3)
#define PM_ERR_ARCH (-4)

/* output modes of pm_gemm_bf16 */
#define PM_OUT_BF16 0    /* bf16 row-major [M, N] (or [M, N/2] with swiglu) via TMA store      */
#define PM_OUT_F32 1     /* fp32 row-major [M, N]                                              */
#define PM_OUT_UNPATCH 2 /* fp32 NCHW image, clamp(-1,1): layers.py:150 + vqmodel.py:30; patch 8; the rows of W
                            (and bias / colsum) must be permuted from the reference's (p1 p2 c) order to
                            (c p1 p2) so that 8 consecutive columns are 8 consecutive pixels               */
#define PM_OUT_UNPATCH_U8 3 /* uint8 NHWC pixels, W rows in the reference's (p1 p2 c) order: + `restore` (reconstruct.py:11-16):
                               u = uint8(255 * ((clamp(v,-1,1) + 1) * 0.5)), patch 8, 3 channels  */

/* Library / device introspection. */
int pm_version(void);                 /* ABI version of this header (1)                        */
int pm_device_check(void);            /* PM_OK iff the current device is sm_100 (B200)         */
const char* pm_error_string(int rc);  /* static string for PM_ERR_* / cudaError_t              */

/* ---------------------------------------------------------------------------------------------
 * Dense projection  D = epilogue(A[M,K] · W[N,K]^T)  on tcgen05 tensor cores.
 * Replaces every nn.Linear / patch Conv2d on the path:
 *   stage1/layers.py:82 (patch embed as GEMM), modules/attention.py:34-41 (to_q/k/v, to_out),
 *   modules/mlp.py:28-31 (w12 + SwiGLU, w3), stage1/vqmodel.py:13-14 (prev/post_quant),
 *   stage1/layers.py:129 (proj), stage2/transformer.py:56,63 (token_proj, to_logits).
 * Epilogue (applied in this order, each optional):
 *   stats/colsum : v = rstd[row] * (v - mu[row] * colsum[col])       (LayerNorm folded in, layers.py:49-58)
 *   bias         : v += bias[col]
 *   pos          : v += pos[(row % pos_rows) * ld_pos + col]         (layers.py:108,146; transformer.py:82)
 *   swiglu       : out[:, j] = silu(v[:, gate j]) * v[:, value j]     (mlp.py:29-30), W rows packed per
 *                  256-row tile as 128 gate rows then 128 value rows (see pm_repack docs in DESIGN.md)
 *   res          : v += res[row, col] (bf16)                          (layers.py:55-56)
 *   stats_out    : row sums / sums of squares of the final values, so that the NEXT projection can fold
 *                  the following LayerNorm without a separate statistics pass over HBM
 * ------------------------------------------------------------------------------------------- */
typedef struct pm_gemm_args {
  const void* a;        /* bf16 [M, K], leading dim lda                                        */
  const void* w;        /* bf16 [N, K], leading dim ldw                                        */
  void* out;            /* see out_mode                                                        */
  const float* bias;    /* [N] or NULL                                                         */
  const float* colsum;  /* [N] or NULL (with stats)                                            */
  const float* stats;   /* [M, 2] = (mean, rstd) per row, or NULL                              */
  const float* pos;     /* [pos_rows, ld_pos] fp32 or NULL                                     */
  const void* res;      /* bf16 [M, N_out], leading dim ld_res, or NULL (PM_OUT_BF16 only)     */
  int64_t lda, ldw, ld_out, ld_pos, ld_res;
  int32_t M, N, K;
  int32_t pos_rows;
  int32_t out_mode;     /* PM_OUT_*                                                            */
  int32_t swiglu;       /* 0/1                                                                 */
  int32_t bn;           /* N tile: 0 = auto, else one of 32/64/128/192/256                     */
  int32_t patch, channels, grid; /* PM_OUT_UNPATCH: patch size (8), channels (3), tokens/side  */
  int32_t max_ctas;     /* 0 = one CTA per SM; >0 caps the persistent grid (tests)             */
  int32_t stats_raw;    /* P > 0: `stats` is [M, P, 2] partial (sum x, sum x^2) pairs over the K columns, as written
                           by `stats_out` of the GEMM that produced A; 0: `stats` is [M, 2] = (mean, rstd)     */
  float ln_eps;         /* LayerNorm eps used with stats_raw (1e-5)                                    */
  float* stats_out;     /* [M, P, 2] fp32 with P = 2 * ceil(N / bn): per-row partial (sum, sum of squares) of the
                           OUTPUT, one pair per N-tile and epilogue warp group (PM_OUT_BF16, no swiglu) or NULL */
  int32_t cta_group;    /* 0 = auto, 1 = one CTA per 128-row tile, 2 = CTA pairs: tcgen05.mma.cta_group::2
                           on 256 x 256 tiles (needs bn == 256)                                        */
  int64_t* debug;       /* optional [grid, 4] int64 stall counters of the MMA issuer (profiling aid) or NULL      */
  int32_t res_mod;      /* 0: `res` is [M, N_out].  > 0 (multiple of 128): `res` is a [res_mod, N_out] bf16 table and
                           output row r adds res[r % res_mod] — the position embedding (layers.py:108,146;
                           transformer.py:82) staged through shared memory by TMA instead of per-thread fp32 loads */
} pm_gemm_args;

int pm_gemm_bf16(const pm_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head attention core, head_dim 64:  O[b, n, h*64:(h+1)*64] = softmax(scale * Q K^T) V.
 * Replaces modules/attention.py:51-58 (CrossAttention) == :84-106 (MemoryEfficientCrossAttention,
 * i.e. xformers.ops.memory_efficient_attention): self-attention (Nq == Nk) and cross-attention on
 * the text context (Nk = 77, no mask — SURVEY.md N6).  Q/K/V/O are bf16, token-major, with head h
 * at columns [h*64, h*64+64) from the given base pointer; ld* = row pitch, bs* = batch stride
 * (elements).  The (b h) n d rearrange of the reference is done by TMA addressing.
 * ------------------------------------------------------------------------------------------- */
typedef struct pm_attn_args {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  int64_t ldq, ldk, ldv, ldo;
  int64_t bsq, bsk, bsv, bso;
  int32_t B, H, Nq, Nk, head_dim; /* head_dim must be 64 */
  float scale;                     /* dim_head ** -0.5 (attention.py:31) */
  float* lse;                      /* optional out [B, H, lse_ld] fp32: log2-sum-exp2 of the scaled score rows, kept by a
                                      training forward for pm_attn_bwd; NULL for inference */
  int64_t lse_ld;                  /* row pitch of lse (0 = Nq); pm_attn_bwd wants Nq rounded up to a multiple of 128 */
  float* o32;                      /* optional out: fp32 copy of O, [B, Nq, ldo32] dense over the batch (training forward: */
  int64_t ldo32;                   /* keeps delta = rowsum(dO * O) of the backward free of O's bf16 rounding); else NULL   */
  int32_t q_prescaled;             /* != 0: Q already carries scale * log2(e) (folded into the to_q weight rows by the caller):
                                      `scale` is ignored, softmax(Q K^T in base 2) is computed by the bias-MMA kernel
                                      (pm_attn3.cu: the row maximum is subtracted by the tensor core) */
  int32_t reserved;
} pm_attn_args;

int pm_attn_fwd(const pm_attn_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Vector quantizer (stage1/quantize.py), e_dim = 32.
 *   pm_vq_codebook_prep : en = l2norm(E) fp32 [n_e,32]; packed = bf16 [n_e,64] = [hi | lo] of en
 *                         (quantize.py:21 — the reference re-normalises E on every forward)
 *   pm_vq_fwd           : VectorQuantizer.forward (quantize.py:18-38): idx = argmin distance
 *                         (first index on ties), zq = zn + (l2norm(E[idx]) - zn), *sse += sum of
 *                         squared differences (loss = (1+beta) * sse / (M*32), quantize.py:33),
 *                         hist[idx] += 1 (codebook usage; new in this build).
 *                         cand_val/cand_idx: scratch [8, M] used when the codebook is split over
 *                         several CTAs per row tile (small M); splits = 0 lets the library choose.
 *   pm_vq_gather        : VectorQuantizer.decode_from_indice (quantize.py:40-44) when normalize=1;
 *                         Pipeline.ids2tokens raw-table gather (generate.py:148-157) when 0.
 *   pm_split_rows32     : fp32 [M,32] -> bf16 [M,64] = [hi | lo], the exact two-term operand the
 *                         post_quant / token_proj GEMMs consume (K = 64 against [W | W]).
 * ------------------------------------------------------------------------------------------- */
typedef struct pm_vq_args {
  const float* z;        /* [M, 32] raw latents, row pitch ldz                                 */
  const float* en;       /* [n_e, 32] from pm_vq_codebook_prep                                 */
  const void* packed;    /* [n_e, 64] bf16 from pm_vq_codebook_prep                            */
  float* cand_val;       /* [8, M] scratch or NULL (then splits must be 1)                     */
  int32_t* cand_idx;     /* [8, M] scratch or NULL                                             */
  int64_t* idx;          /* [M] out                                                            */
  float* zq;             /* [M, 32] out or NULL                                                */
  void* zq_split;        /* [M, 64] bf16 out or NULL                                           */
  double* sse;           /* accumulated (caller zeroes) or NULL                                */
  uint64_t* hist;        /* [n_e] accumulated or NULL                                          */
  int64_t ldz;
  int32_t M, n_e, e_dim, splits;
} pm_vq_args;

int pm_vq_codebook_prep(const float* E, int32_t n_e, int32_t e_dim, float* en, void* packed, void* stream);
int pm_vq_fwd(const pm_vq_args* args, void* stream);
int pm_vq_gather(const int64_t* idx, int32_t M, int32_t n_rows, int32_t e_dim, const float* table,
                 int32_t normalize, float* out, void* out_split, void* stream);
int pm_split_rows32(const float* src, int64_t ld, int32_t M, void* out_split, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HBM-bound row kernels.
 *   pm_patchify8 : im2col of the stride-8 patch conv (stage1/layers.py:82-83): fp32 NCHW image ->
 *                  bf16 [B*(H/8)*(W/8), C*64], K order (c, kh, kw) = flattened Conv2d weight.
 *   pm_layernorm : nn.LayerNorm over bf16 rows (layers.py:49,51,89,128; eps 1e-5).
 *                  y == NULL: only stats[row] = (mean, rstd) (feeds the LN-folded GEMM epilogue);
 *                  y != NULL: y = LN(x) * gamma + beta (bf16) and, if stats != NULL, the stats of y.
 *   pm_patchify8_u8 : the same im2col fed from decoded pixels, uint8 NHWC [B, H, W, 3] (8-byte aligned), with the
 *                  reference's ingest transform fused (utils/transform.py:17-18: ToTensor u/255, Normalize
 *                  (t-0.5)/0.5, evaluated in fp32 exactly as torchvision does) — SURVEY.md §8f row 3.
 *                  W <= 256 (one patch row of the image per block iteration; the reference models are 256 x 256).
 */
/* pm_cast_f32_bf16 : fp32 -> bf16 (n % 8 == 0); the text context entering cross-attention k/v (transformer.py:84-86). */
int pm_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
int pm_patchify8_u8(const uint8_t* img, void* out, int32_t B, int32_t H, int32_t W, void* stream);
int pm_patchify8(const float* img, void* out, int32_t B, int32_t C, int32_t H, int32_t W, void* stream);
int pm_layernorm(const void* x, int64_t ldx, int32_t M, int32_t D, float eps, const float* gamma,
                 const float* beta, void* y, int64_t ldy, float* stats, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MaskGIT step tail (generate.py:159-181), one pass over the fp32 logits [M = B*N, V]:
 *   pm_maskgit_sample : top_k filter (generate.py:33-37) + gumbel_sample (:40-46) + mask fill (:166-168)
 *                       + confidence score 1 - softmax(logits)[pred], -1e5 at unmasked positions (:170-173).
 *                       noise != NULL injects the uniforms the reference would have drawn at [row, index]
 *                       (parity tests); otherwise Philox4x32-10 keyed on (seed, row, index, offset).
 *                       topk in [1, 32].
 *   pm_maskgit_remask : ids.scatter(1, scores.topk(k).indices, mask_id) per image (generate.py:175-179);
 *                       ties at the k-th score resolve to the lower token index.
 * ------------------------------------------------------------------------------------------- */
/* Per-step scalars of the MaskGIT loop (generate.py:186-192: temperature * (1 - step / T), the number of tokens to re-mask,
 * the noise key) in DEVICE memory, one entry per step.  When `step_tab` is given, the kernels read entry *step_idx instead of
 * their by-value arguments, and pm_maskgit_remask_step — the last kernel of a step — advances *step_idx: one captured CUDA
 * graph (gather -> transformer -> sample -> re-mask) then serves every step of generate(), including the RNG stream. */
typedef struct pm_step_scalars {
  float temperature;
  int32_t k;              /* tokens to re-mask after this step */
  uint64_t seed, offset;  /* Philox key / stream offset of this step's gumbel noise */
} pm_step_scalars;

typedef struct pm_maskgit_sample_args {
  const float* logits;
  const float* noise;     /* optional [M, V] uniforms in [0,1) */
  int64_t* ids;           /* [M] in/out, optional */
  int64_t* pred_ids;      /* [M] out */
  float* scores;          /* [M] out */
  int64_t ld, ld_noise;
  int64_t mask_id;
  uint64_t seed, offset;
  int32_t M, V, topk;
  float temperature;
  const pm_step_scalars* step_tab; /* optional (with step_idx): temperature / seed / offset come from step_tab[*step_idx] */
  const int32_t* step_idx;
} pm_maskgit_sample_args;

int pm_maskgit_sample(const pm_maskgit_sample_args* args, void* stream);
int pm_maskgit_remask(const float* scores, int64_t* ids, int32_t B, int32_t N, int32_t k, int64_t mask_id, void* stream);
/* pm_maskgit_remask with k = step_tab[*step_idx].k; afterwards *step_idx += 1 (`ticket`: one zero-initialised int32 of scratch) */
int pm_maskgit_remask_step(const float* scores, int64_t* ids, int32_t B, int32_t N, int64_t mask_id, const pm_step_scalars* step_tab,
                           int32_t* step_idx, int32_t* ticket, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage-2 training forward, the pieces around the transformer (generate.py:78-146):
 *   pm_maskgit_random_mask : Pipeline.random_masking (generate.py:78-108).  Per image the len_keep tokens with the
 *                            smallest noise keep their latent z[b, i, :32], the others become mask_token;
 *                            mask[b, i] = 1 where replaced (ties in the noise: lower token index is kept first).
 *                            noise != NULL injects the uniforms torch.rand(B, N) would have supplied (parity tests);
 *                            otherwise Philox4x32-10 keyed on (seed, image, token, offset).  x_out may be NULL
 *                            (mask only).
 *   pm_ce_label_smooth     : Pipeline.loss (generate.py:110-123): row_loss[i] = mask[i] * cross_entropy(logits[i],
 *                            label[i], label_smoothing) (rows with mask 0 are skipped, not read); loss_out =
 *                            sum(row_loss) / sum(mask) reduced in a fixed order in fp64; sums_out[0:2] = the two sums
 *                            (for a cross-rank all-reduce).  mask == NULL counts every row.
 * ------------------------------------------------------------------------------------------- */
int pm_maskgit_random_mask(const float* z, int64_t ldz, const float* noise, uint64_t seed, uint64_t offset,
                           const float* mask_token, int32_t B, int32_t N, int32_t len_keep, float* mask, float* x_out,
                           void* stream);
int pm_ce_label_smooth(const float* logits, int64_t ld, int32_t M, int32_t V, const int64_t* label, const float* mask,
                       float label_smoothing, float* row_loss, float* loss_out, double* sums_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Generator BACKWARD path.  The reference has no backward code of its own: VQGANTrainer
 * (utils/trainer.py:205-225) calls accelerator.backward(loss) and torch autograd differentiates the modules of the
 * forward path.  These entry points are what a torch.autograd.Function.backward of that path calls
 * (paintmind_b200/train.py); every one of them cites the forward lines it differentiates.
 *
 *   dgrad (dX = dY W) of every nn.Linear is pm_gemm_bf16 itself with the transposed weight as `w`.
 *   pm_wgrad_bf16       : dW[N, K] (+)= dY[M, N]^T X[M, K]  — nn.Linear / patch-conv weight gradients
 *                         (attention.py:34-41, mlp.py:28-31, layers.py:82,129, vqmodel.py:13-14).  dY and X are read
 *                         token-major in place (MN-major UMMA operands), the M dimension is split across CTAs and the
 *                         partial tiles are summed in a fixed order.  lddy / ldx must cover N / K rounded up to 64
 *                         columns.  work: pm_wgrad_workspace_floats(M, N, K) floats.
 *   pm_colsum_bf16      : out[N] (+)= sum_m x[m, N] — bias gradients; position-embedding gradients when x is viewed
 *                         as [B, tokens * D] (layers.py:108,146).  work: pm_colsum_workspace_floats(M, N) floats.
 *   pm_layernorm_bwd    : nn.LayerNorm backward (layers.py:49,51,89,128): dx = LN'(x)^T (dn) (+ dres, the residual
 *                         branch of layers.py:55-56), dgamma_dbeta[2, D] = (sum dn * xhat, sum dn).
 *                         work: pm_layernorm_bwd_workspace_floats(M, D) floats.
 *   pm_swiglu_bwd       : hidden = silu(x1) * x2 backward (mlp.py:29-30) on the tile-interleaved x12 the packed w12
 *                         projection produces; also re-materialises `h` (operand of the w3 weight gradient) and, when
 *                         b12 != NULL, the bias gradient of w12 (column sums of d12, [2 hp] in the packed order) in the
 *                         same pass.  work: pm_swiglu_bwd_workspace_floats(M, hp) floats (with b12) or NULL.
 *   pm_attn_bwd         : backward of softmax(scale Q K^T) V (attention.py:52-57): dq, dk, dv from q, k, v, o, d_o and
 *                         the forward's lse.  delta: [2, B, H, lse_ld] fp32 scratch.  Ragged token counts are masked.
 *   pm_vq_bwd           : VectorQuantizer backward (quantize.py:19,29-36): straight-through estimator + both loss
 *                         terms; dz fp32 [M, 32], dz_split bf16 [M, 64] = [hi | lo], dE[n_e, 32] += (fp32 atomics).
 *                         d_out: gradient of the returned z_q (may be NULL), d_loss: device scalar (may be NULL).
 *   pm_unpatchify8_bwd  : backward of clamp(-1, 1) (vqmodel.py:30) + un-patchify (layers.py:150): d_img fp32 NCHW ->
 *                         bf16 rows [B * tokens, C * 64] in (c p1 p2) order, zero where rec sits on a clamp bound.
 * ------------------------------------------------------------------------------------------- */
typedef struct pm_attn_bwd_args {
  const void* q;
  const void* k;
  const void* v;
  const void* o;      /* forward output: bf16, or the fp32 copy (o32 of pm_attn_fwd) when o_is_f32 */
  const void* d_o;    /* gradient of o */
  const float* lse;   /* [B, H, lse_ld] from pm_attn_fwd */
  float* delta;       /* [2, B, H, lse_ld] scratch (the row term -scale * rowsum(d_o * o), and -lse) */
  void* dq;
  void* dk;
  void* dv;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int64_t bsq, bsk, bsv, bso, bsdo, bsdq, bsdk, bsdv;
  int32_t B, H, Nq, Nk, head_dim;
  float scale;
  int32_t o_is_f32;
  int64_t lse_ld;     /* row pitch of lse / delta: Nq rounded up to a multiple of 128 */
  int64_t* debug;     /* optional [#SMs, 8] int64 cycle counters of the last kernel (profiling aid) or NULL */
} pm_attn_bwd_args;

int pm_attn_bwd(const pm_attn_bwd_args* args, void* stream);
int64_t pm_wgrad_workspace_floats(int32_t M, int32_t N, int32_t K);
int pm_wgrad_bf16(const void* dy, int64_t lddy, const void* x, int64_t ldx, int32_t M, int32_t N, int32_t K, float* work,
                  float* out, int64_t ld_out, int32_t accumulate, void* stream);
int64_t pm_colsum_workspace_floats(int32_t M, int32_t N);
int pm_colsum_bf16(const void* x, int64_t ld, int32_t M, int32_t N, float* work, float* out, int32_t accumulate, void* stream);
int64_t pm_layernorm_bwd_workspace_floats(int32_t M, int32_t D);
int pm_layernorm_bwd(const void* dn, int64_t lddn, const void* x, int64_t ldx, const float* gamma, const void* dres, int64_t ldres,
                     void* dx, int64_t lddx, int32_t M, int32_t D, float eps, float* work, float* dgamma_dbeta, void* stream);
int64_t pm_swiglu_bwd_workspace_floats(int32_t M, int32_t hp);
int pm_swiglu_bwd(const void* x12, int64_t ld12, const void* dh, int64_t lddh, void* h, int64_t ldh, void* d12, int64_t ldd12,
                  int32_t M, int32_t hp, float* work, float* b12, void* stream);
int pm_vq_bwd(const float* z, int64_t ldz, const int64_t* idx, const float* E, int32_t e_dim, const float* d_out, int64_t ldd,
              const float* d_loss, float beta, int32_t M, float* dz, void* dz_split, float* dE, void* stream);
int pm_unpatchify8_bwd(const float* d_img, const float* rec, void* out, int32_t B, int32_t C, int32_t H, int32_t W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PAINTMIND_B200_H_ */
