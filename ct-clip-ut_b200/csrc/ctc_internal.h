// Internal C++ prototypes shared between the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctclip_b200.h"

namespace ctc {

int num_sms();

// gemm.cu
int gemm_row_perm(int* perm32);
int gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* out, long long ldc, int M, int N,
              int K, int epi, const float* bias, const float* resid, long long ldr, void* aux, long long ldaux,
              float* top2_val, int* top2_idx, int impl, cudaStream_t st);
int gemm_argmax_candidates(int N);

// peg_tma.cu: 0 launched, -1 not eligible (fall back to the register-window kernel), > 0 error
int peg_tma_launch(const float* x, int B, int T, int H, int W, int C, const float* w27, const float* bias, int mode,
                   int transpose, float* y, void* y_bf16, cudaStream_t st);

}  // namespace ctc
