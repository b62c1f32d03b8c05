// pm_kernels.h — internal (C++) declarations shared by the .cu files of libpaintmind_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

enum : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_UNPATCH = 2 };

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
  int pos_rows;
  int patch, channels, grid;
  int max_ctas;
};

int pm_num_sms();
int pm_gemm_launch(const GemmParams& p, int bn, int out_mode, int swiglu, cudaStream_t stream);

}  // namespace pm
