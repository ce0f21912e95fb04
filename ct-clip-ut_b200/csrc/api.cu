// C-ABI glue that is not tied to one kernel family.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static long long g_launches = 0;
void count_launch() { ++g_launches; }
}  // namespace ctc

using namespace ctc;

extern "C" int ctc_version(void) { return CTC_VERSION; }
extern "C" long long ctc_launch_count(void) { return g_launches; }
extern "C" const char* ctc_last_error(void) { return g_err; }

extern "C" int ctc_device_check(void) {
    int dev = 0;
    CTC_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CTC_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    CTC_REQUIRE(prop.major == 10, "ctclip_b200 is sm_100a only: device %d (%s) is sm_%d%d; there is no fallback path",
                dev, prop.name, prop.major, prop.minor);
    return 0;
}

extern "C" int ctc_gemm_row_perm(int* perm32) { return gemm_row_perm(perm32); }
extern "C" int ctc_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldc, int M,
                             int N, int K, int epi, const float* bias, const float* resid, int64_t ldr, void* aux,
                             int64_t ldaux, int impl, void* stream) {
    CTC_REQUIRE(epi != CTC_EPI_ARGMAX, "ctc_gemm_bf16: the arg-max epilogue is reached through ctc_vq_argmax");
    return gemm_bf16(A, lda, B, ldb, out, ldc, M, N, K, epi, bias, resid, ldr, aux, ldaux, nullptr, nullptr, impl,
                     (cudaStream_t)stream);
}
