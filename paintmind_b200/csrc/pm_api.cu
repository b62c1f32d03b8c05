// pm_api.cu — extern "C" entry points declared in include/paintmind_b200.h.
#include "../../include/paintmind_b200.h"
#include "pm_common.cuh"
#include "pm_kernels.h"

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
  if ((p.colsum == nullptr) != (p.stats == nullptr)) return PM_ERR_INVALID;
  int bn = a->bn;
  if (bn == 0) {
    if (a->swiglu) bn = 256;
    else if (a->out_mode == PM_OUT_UNPATCH) bn = (p.N % 192 == 0) ? 192 : 64;
    else if (p.N % 256 == 0) bn = 256;
    else if (p.N % 128 == 0) bn = 128;
    else if (p.N % 64 == 0 || a->out_mode == PM_OUT_BF16) bn = 64;
    else bn = 32;
  }
  if (a->out_mode == PM_OUT_UNPATCH && (p.patch <= 0 || p.channels <= 0 || p.grid <= 0 ||
                                        p.N != p.patch * p.patch * p.channels))
    return PM_ERR_INVALID;
  return pm_gemm_launch(p, bn, a->out_mode, a->swiglu, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
