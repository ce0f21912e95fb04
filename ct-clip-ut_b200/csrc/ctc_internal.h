// Internal C++ prototypes shared between the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctclip_b200.h"

namespace ctc {

int num_sms();

// gemm.cu
int gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* out, long long ldc, int M, int N,
              int K, int epi, const float* bias, const float* resid, long long ldr, void* aux, long long ldaux,
              float* top2_val, int* top2_idx, int impl, cudaStream_t st);
int gemm_argmax_candidates(int N);

}  // namespace ctc
