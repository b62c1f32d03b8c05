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
#define PM_OUT_UNPATCH 2 /* fp32 NCHW image, clamp(-1,1): layers.py:150 + vqmodel.py:30        */

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
} pm_gemm_args;

int pm_gemm_bf16(const pm_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PAINTMIND_B200_H_ */
