// pm_api.cu — extern "C" entry points declared in include/paintmind_b200.h.
#include "../../include/paintmind_b200.h"
#include "pm_common.cuh"
#include "pm_kernels.h"

#include <stdlib.h>

using namespace pm;

extern "C" {

int pm_version(void) { return 1; }

int pm_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  return major == 10 ? PM_OK : PM_ERR_ARCH;
}

const char* pm_error_string(int rc) {
  switch (rc) {
    case PM_OK: return "ok";
    case PM_ERR_INVALID: return "invalid argument (null pointer, misaligned or unsupported shape)";
    case PM_ERR_TENSORMAP: return "cuTensorMapEncodeTiled failed";
    case PM_ERR_NO_DRIVER: return "CUDA driver entry point cuTensorMapEncodeTiled not available";
    case PM_ERR_ARCH: return "device is not sm_100 (B200)";
    default: break;
  }
  if (rc > 0) return cudaGetErrorString(static_cast<cudaError_t>(rc));
  return "unknown error";
}

int pm_gemm_bf16(const pm_gemm_args* a, void* stream) {
  if (a == nullptr) return PM_ERR_INVALID;
  GemmParams p;
  p.a = a->a; p.w = a->w; p.out = a->out;
  p.bias = a->bias; p.colsum = a->colsum; p.stats = a->stats; p.pos = a->pos; p.res = a->res;
  p.lda = a->lda; p.ldw = a->ldw; p.ld_out = a->ld_out; p.ld_pos = a->ld_pos; p.ld_res = a->ld_res;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.pos_rows = a->pos_rows > 0 ? a->pos_rows : 1;
  p.patch = a->patch; p.channels = a->channels; p.grid = a->grid;
  p.max_ctas = a->max_ctas;
  p.stats_out = a->stats_out; p.stats_raw = a->stats_raw; p.ln_eps = a->ln_eps > 0.0f ? a->ln_eps : 1e-5f;
  if (p.stats_out != nullptr && (a->out_mode != PM_OUT_BF16 || a->swiglu)) return PM_ERR_INVALID;
  if (p.stats_raw < 0 || p.stats_raw > 8 || (p.stats_raw & 1)) return PM_ERR_INVALID;   // partial pairs come in (tile, group) twos
  // auto: CTA pairs (256 x 256 tiles, half the W traffic per SM) whenever the tile shape allows it.  Measured on B200 at
  // M = 262144, single vs pair: qkv 0.389 / 0.352 ms, out 0.163 / 0.158, w12+SwiGLU 0.652 / 0.602, w3 0.357 / 0.319.
  p.cta_pair = a->cta_group == 1 ? 0 : 1;
  p.debug = reinterpret_cast<long long*>(a->debug);
  p.res_mod = a->res_mod;
  if ((p.colsum == nullptr) != (p.stats == nullptr)) return PM_ERR_INVALID;
  int bn = a->bn;
  if (bn == 0) {
    if (a->swiglu) bn = 256;
    else if (a->out_mode == PM_OUT_UNPATCH || a->out_mode == PM_OUT_UNPATCH_U8) bn = (p.N % 192 == 0) ? 192 : 64;
    else if (p.N % 256 == 0) bn = 256;
    else if (p.N % 128 == 0) bn = 128;
    else if (p.N % 64 == 0 || a->out_mode == PM_OUT_BF16) bn = 64;
    else bn = 32;
  }
  if ((a->out_mode == PM_OUT_UNPATCH || a->out_mode == PM_OUT_UNPATCH_U8) && (p.patch <= 0 || p.channels <= 0 || p.grid <= 0 ||
                                        p.N != p.patch * p.patch * p.channels))
    return PM_ERR_INVALID;
  return pm_gemm_launch(p, bn, a->out_mode, a->swiglu, static_cast<cudaStream_t>(stream));
}

int pm_attn_fwd(const pm_attn_args* a, void* stream) {
  if (a == nullptr) return PM_ERR_INVALID;
  AttnParams p;
  p.q = a->q; p.k = a->k; p.v = a->v; p.o = a->o;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.bsq = a->bsq; p.bsk = a->bsk; p.bsv = a->bsv; p.bso = a->bso;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.head_dim = a->head_dim;
  p.prescaled = a->q_prescaled != 0 ? 1 : 0;
  p.scale_log2 = p.prescaled ? 1.0f : a->scale * 1.4426950408889634f;
  p.lse = a->lse;
  p.lse_ld = a->lse_ld > 0 ? a->lse_ld : a->Nq;
  p.o32 = a->o32;
  p.ldo32 = a->ldo32;
  if (p.o32 != nullptr && ((p.ldo32 % 4) != 0 || (reinterpret_cast<uintptr_t>(p.o32) & 15) != 0)) return PM_ERR_INVALID;
  // PM_ATTN_IMPL=1: the round-1 kernel (8 softmax warps, pm_attn.cu); 2: the 16-softmax-warp kernel (pm_attn2.cu).  A/B aid.
  static int impl = 0;
  if (impl == 0) {
    const char* env = getenv("PM_ATTN_IMPL");
    impl = (env != nullptr && env[0] == '2') ? 2 : 1;
  }
  // pre-scaled queries: the bias-MMA kernel (PM_ATTN_IMPL set = A/B against the older kernels, which take scale_log2 = 1)
  if (p.prescaled && getenv("PM_ATTN_IMPL") == nullptr && pm_attn4_supported(p)) {
    // default: the 16-softmax-warp kernel (pm_attn4.cu; sustained, i.e. power-capped: 0.721 ms at B = 256, H = 8, N = 1024 against
    // 0.734-0.741 for the 8-warp pm_attn3.cu and 0.759 for pm_attn.cu — profiles/r02_attention.md).  PM_ATTN_PRE=3 picks pm_attn3.cu.
    static int pre_impl = 0;
    if (pre_impl == 0) {
      const char* env = getenv("PM_ATTN_PRE");
      pre_impl = (env != nullptr && env[0] == '3') ? 3 : 4;
    }
    if (pre_impl == 4 || p.lse != nullptr || p.o32 != nullptr) return pm_attn4_launch(p, static_cast<cudaStream_t>(stream));
    return pm_attn3_launch(p, static_cast<cudaStream_t>(stream));
  }
  if (impl == 2) return pm_attn2_launch(p, static_cast<cudaStream_t>(stream));
  return pm_attn_launch(p, static_cast<cudaStream_t>(stream));
}

int pm_attn_bwd(const pm_attn_bwd_args* a, void* stream) {
  if (a == nullptr || a->o == nullptr || a->delta == nullptr) return PM_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->lse == nullptr || a->B <= 0 || a->H <= 0 || a->lse_ld < a->Nq) return PM_ERR_INVALID;
  float* nds = a->delta;
  float* nlse = a->delta + static_cast<size_t>(a->B) * a->H * a->lse_ld;
  int rc = pm_attn_delta_launch(a->o, a->o_is_f32, a->ldo, a->bso, a->d_o, a->lddo, a->bsdo, a->B, a->H, a->Nq, a->lse, a->scale, nds, nlse,
                                a->lse_ld, st);
  if (rc != 0) return rc;
  AttnBwdParams p;
  p.q = a->q; p.k = a->k; p.v = a->v; p.dO = a->d_o; p.dq = a->dq; p.dk = a->dk; p.dv = a->dv;
  p.nlse = nlse; p.nds = nds; p.lse_ld = a->lse_ld;
  p.debug = reinterpret_cast<long long*>(a->debug);
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.lddo = a->lddo; p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.bsq = a->bsq; p.bsk = a->bsk; p.bsv = a->bsv; p.bsdo = a->bsdo; p.bsdq = a->bsdq; p.bsdk = a->bsdk; p.bsdv = a->bsdv;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.head_dim = a->head_dim;
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  return pm_attn_bwd_launch(p, st);
}

int64_t pm_wgrad_workspace_floats(int32_t M, int32_t N, int32_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return static_cast<int64_t>(pm_wgrad_splits(M, N, K)) * N * K;
}

int pm_wgrad_bf16(const void* dy, int64_t lddy, const void* x, int64_t ldx, int32_t M, int32_t N, int32_t K, float* work,
                  float* out, int64_t ld_out, int32_t accumulate, void* stream) {
  WgradParams p;
  p.dy = dy; p.x = x; p.work = work; p.out = out; p.lddy = lddy; p.ldx = ldx; p.ld_out = ld_out;
  p.M = M; p.N = N; p.K = K; p.splits = 0; p.accumulate = accumulate; p.scale = 1.0f;
  return pm_wgrad_launch(p, static_cast<cudaStream_t>(stream));
}

int64_t pm_colsum_workspace_floats(int32_t M, int32_t N) {
  if (M <= 0 || N <= 0) return 0;
  return static_cast<int64_t>(pm_colsum_rows(M, N)) * N;
}

int pm_colsum_bf16(const void* x, int64_t ld, int32_t M, int32_t N, float* work, float* out, int32_t accumulate, void* stream) {
  return pm_colsum_launch(x, ld, M, N, work, out, accumulate, static_cast<cudaStream_t>(stream));
}

int64_t pm_layernorm_bwd_workspace_floats(int32_t M, int32_t D) {
  if (M <= 0 || D <= 0) return 0;
  return static_cast<int64_t>(pm_ln_bwd_blocks(M)) * 2 * D;
}

int pm_layernorm_bwd(const void* dn, int64_t lddn, const void* x, int64_t ldx, const float* gamma, const void* dres, int64_t ldres,
                     void* dx, int64_t lddx, int32_t M, int32_t D, float eps, float* work, float* dgamma_dbeta, void* stream) {
  return pm_ln_bwd_launch(dn, lddn, x, ldx, gamma, dres, ldres, dx, lddx, M, D, eps, work, dgamma_dbeta, static_cast<cudaStream_t>(stream));
}

int64_t pm_swiglu_bwd_workspace_floats(int32_t M, int32_t hp) {
  if (M <= 0 || hp <= 0) return 0;
  return static_cast<int64_t>(pm_swiglu_bwd_chunks(M, hp)) * 2 * hp;
}

int pm_swiglu_bwd(const void* x12, int64_t ld12, const void* dh, int64_t lddh, void* h, int64_t ldh, void* d12, int64_t ldd12,
                  int32_t M, int32_t hp, float* work, float* b12, void* stream) {
  return pm_swiglu_bwd_launch(x12, ld12, dh, lddh, h, ldh, d12, ldd12, M, hp, work, b12, static_cast<cudaStream_t>(stream));
}

int pm_vq_bwd(const float* z, int64_t ldz, const int64_t* idx, const float* E, int32_t e_dim, const float* d_out, int64_t ldd,
              const float* d_loss, float beta, int32_t M, float* dz, void* dz_split, float* dE, void* stream) {
  if (e_dim != 32) return PM_ERR_INVALID;
  return pm_vq_bwd_launch(z, ldz, reinterpret_cast<const long long*>(idx), E, d_out, ldd, d_loss, beta, M, dz, dz_split, dE,
                          static_cast<cudaStream_t>(stream));
}

int pm_unpatchify8_bwd(const float* d_img, const float* rec, void* out, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  return pm_unpatchify_bwd_launch(d_img, rec, out, B, C, H, W, static_cast<cudaStream_t>(stream));
}

int pm_vq_codebook_prep(const float* E, int32_t n_e, int32_t e_dim, float* en, void* packed, void* stream) {
  if (e_dim != 32) return PM_ERR_INVALID;
  return pm_vq_codebook_prep_launch(E, n_e, en, packed, static_cast<cudaStream_t>(stream));
}

int pm_vq_fwd(const pm_vq_args* a, void* stream) {
  if (a == nullptr) return PM_ERR_INVALID;
  VqParams p;
  p.z = a->z; p.ldz = a->ldz; p.M = a->M;
  p.en = a->en; p.packed = a->packed; p.n_e = a->n_e; p.e_dim = a->e_dim;
  p.splits = a->splits;
  p.cand_val = a->cand_val; p.cand_idx = a->cand_idx;
  p.idx = reinterpret_cast<long long*>(a->idx);
  p.zq = a->zq; p.zq_split = a->zq_split; p.sse = a->sse;
  p.hist = reinterpret_cast<unsigned long long*>(a->hist);
  if (p.cand_val == nullptr || p.cand_idx == nullptr) p.splits = 1;
  return pm_vq_launch(p, static_cast<cudaStream_t>(stream));
}

int pm_vq_gather(const int64_t* idx, int32_t M, int32_t n_rows, int32_t e_dim, const float* table,
                 int32_t normalize, float* out, void* out_split, void* stream) {
  if (e_dim != 32) return PM_ERR_INVALID;
  return pm_vq_gather_launch(reinterpret_cast<const long long*>(idx), M, n_rows, table, normalize, out, out_split,
                             static_cast<cudaStream_t>(stream));
}

int pm_split_rows32(const float* src, int64_t ld, int32_t M, void* out_split, void* stream) {
  return pm_split_rows32_launch(src, ld, M, out_split, static_cast<cudaStream_t>(stream));
}

int pm_patchify8(const float* img, void* out, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  return pm_patchify_launch(img, out, B, C, H, W, 8, static_cast<cudaStream_t>(stream));
}

int pm_patchify8_u8(const uint8_t* img, void* out, int32_t B, int32_t H, int32_t W, void* stream) {
  return pm_patchify_u8_launch(img, out, B, H, W, static_cast<cudaStream_t>(stream));
}

int pm_layernorm(const void* x, int64_t ldx, int32_t M, int32_t D, float eps, const float* gamma,
                 const float* beta, void* y, int64_t ldy, float* stats, void* stream) {
  return pm_layernorm_launch(x, ldx, M, D, eps, gamma, beta, y, ldy, stats, static_cast<cudaStream_t>(stream));
}

int pm_maskgit_sample(const pm_maskgit_sample_args* a, void* stream) {
  if (a == nullptr) return PM_ERR_INVALID;
  MaskgitParams p;
  p.logits = a->logits; p.ld = a->ld; p.M = a->M; p.V = a->V; p.topk = a->topk; p.temperature = a->temperature;
  p.noise = a->noise; p.ld_noise = a->ld_noise; p.seed = a->seed; p.offset = a->offset;
  p.ids = reinterpret_cast<long long*>(a->ids);
  p.pred_ids = reinterpret_cast<long long*>(a->pred_ids);
  p.scores = a->scores; p.mask_id = a->mask_id;
  p.step_tab = reinterpret_cast<const StepScalars*>(a->step_tab);
  p.step_idx = a->step_idx;
  if (p.step_tab != nullptr && p.step_idx == nullptr) return PM_ERR_INVALID;
  return pm_maskgit_sample_launch(p, static_cast<cudaStream_t>(stream));
}

int pm_maskgit_remask(const float* scores, int64_t* ids, int32_t B, int32_t N, int32_t k, int64_t mask_id, void* stream) {
  return pm_maskgit_remask_launch(scores, reinterpret_cast<long long*>(ids), B, N, k, mask_id, nullptr, nullptr, nullptr,
                                  static_cast<cudaStream_t>(stream));
}

int pm_maskgit_remask_step(const float* scores, int64_t* ids, int32_t B, int32_t N, int64_t mask_id, const pm_step_scalars* step_tab,
                           int32_t* step_idx, int32_t* ticket, void* stream) {
  static_assert(sizeof(pm_step_scalars) == sizeof(StepScalars), "pm_step_scalars layout");
  if (step_tab == nullptr) return PM_ERR_INVALID;
  return pm_maskgit_remask_launch(scores, reinterpret_cast<long long*>(ids), B, N, 0, mask_id,
                                  reinterpret_cast<const StepScalars*>(step_tab), step_idx, ticket, static_cast<cudaStream_t>(stream));
}

int pm_maskgit_random_mask(const float* z, int64_t ldz, const float* noise, uint64_t seed, uint64_t offset,
                           const float* mask_token, int32_t B, int32_t N, int32_t len_keep, float* mask, float* x_out,
                           void* stream) {
  return pm_maskgit_random_mask_launch(z, ldz, noise, seed, offset, mask_token, B, N, len_keep, mask, x_out,
                                       static_cast<cudaStream_t>(stream));
}

int pm_ce_label_smooth(const float* logits, int64_t ld, int32_t M, int32_t V, const int64_t* label, const float* mask,
                       float label_smoothing, float* row_loss, float* loss_out, double* sums_out, void* stream) {
  return pm_ce_label_smooth_launch(logits, ld, M, V, reinterpret_cast<const long long*>(label), mask, label_smoothing,
                                   row_loss, loss_out, sums_out, static_cast<cudaStream_t>(stream));
}

int pm_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  return pm_cast_launch(src, dst, n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
