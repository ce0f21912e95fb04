// Common device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// warp reductions, bf16 packing.  Hand-written PTX (no CUTLASS dependency).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ctc {

#define CTC_DEVINL __device__ __forceinline__

// ----------------------------------------------------------------------------------------
// status / error plumbing (host)
// ----------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define CTC_CHECK_CUDA(expr)                                                            \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::ctc::set_last_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,        \
                                  cudaGetErrorName(_e), cudaGetErrorString(_e));        \
            return 1;                                                                   \
        }                                                                               \
    } while (0)
#define CTC_REQUIRE(cond, ...)                                                          \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            ::ctc::set_last_error(__VA_ARGS__);                                         \
            return 2;                                                                   \
        }                                                                               \
    } while (0)
void count_launch();
// The dynamic-shared-memory attribute of a kernel and the SM count are per DEVICE: every "configured once" flag is
// keyed by the current device so that one process can drive engines on several GPUs.
static constexpr int kMaxDevices = 64;
inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < kMaxDevices) ? d : 0;
}
#define CTC_LAUNCH_CHECK()                    \
    do {                                      \
        ::ctc::count_launch();                \
        CTC_CHECK_CUDA(cudaGetLastError());   \
    } while (0)

// ----------------------------------------------------------------------------------------
// small math / packing
// ----------------------------------------------------------------------------------------
CTC_DEVINL uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
CTC_DEVINL float2 unpack_bf16(uint32_t v) {
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(b);
}
CTC_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CTC_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// erf-GELU (F.gelu default, attention.py:41) and its derivative from ONE exponential:
//   erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1/(1 + p z)   (Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7)
// with z = |x|/sqrt(2), so exp(-z^2) = exp(-x^2/2) is also the Gaussian density the derivative needs.
CTC_DEVINL void gelu_parts(float x, float& cdf, float& pdf) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    const float e = __expf(-z * z);
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f),
                                0.254829592f);
    const float erf_abs = fmaf(-poly, e, 1.0f);
    cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
    pdf = 0.3989422804014327f * e;
}
// GEGLU of two adjacent (value, gate) pairs for the GEMM epilogues, where this arithmetic - not the tensor core - paces
// the tile: raw approximate instructions (MUFU.RCP / MUFU.EX2, flush-to-zero, no range-scaling wrappers) and PACKED fp32
// arithmetic (fma / mul .f32x2 -> FFMA2 / FMUL2: one issue slot for both pairs).  Abramowitz-Stegun 7.1.26 with the
// polynomial negated (pn = -poly):  erf|g/sqrt2| = 1 + pn * exp(-g^2/2),  cdf = 0.5 + 0.5 * sign(g) * erf,
//   a = gelu(g) = g * cdf,   h = a * x,   b = x * gelu'(g) = x * (cdf + g * pdf),   pdf = exp(-g^2/2) / sqrt(2 pi).
// 12 issue slots per pair with the adjoint factors (22 with scalar instructions), 8 without.  Every lane operation is a
// correctly rounded fma / mul, so the result does not depend on how pairs are grouped into packed instructions.
struct F2 { unsigned long long v; };
CTC_DEVINL F2 f2_pack(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
CTC_DEVINL void f2_unpack(F2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
CTC_DEVINL F2 f2_fma(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
CTC_DEVINL F2 f2_mul(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
CTC_DEVINL F2 f2_c(float c) { return f2_pack(c, c); }
template <bool FACTORS>
CTC_DEVINL void geglu_pair2(float x0, float x1, float g0, float g1, F2& h, F2& a, F2& b) {
    float t0, t1, e0, e1, w0, w1, r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(__fmaf_rn(fabsf(g0), 0.23164189f, 1.0f)));    // 0.3275911 / sqrt(2)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(__fmaf_rn(fabsf(g1), 0.23164189f, 1.0f)));
    const F2 G = f2_pack(g0, g1), X = f2_pack(x0, x1), T = f2_pack(t0, t1);
    f2_unpack(f2_mul(f2_mul(G, f2_c(-0.72134752f)), G), w0, w1);                                      // -g^2/2 * log2(e)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(w0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(w1));
    const F2 E = f2_pack(e0, e1);
    F2 P = f2_fma(T, f2_c(-1.061405429f), f2_c(1.453152027f));
    P = f2_fma(P, T, f2_c(-1.421413741f));
    P = f2_fma(P, T, f2_c(0.284496736f));
    P = f2_fma(P, T, f2_c(-0.254829592f));
    P = f2_mul(P, T);
    f2_unpack(f2_fma(P, E, f2_c(1.0f)), r0, r1);
    const F2 C = f2_fma(f2_pack(copysignf(r0, g0), copysignf(r1, g1)), f2_c(0.5f), f2_c(0.5f));
    a = f2_mul(G, C);
    h = f2_mul(a, X);
    if constexpr (FACTORS) b = f2_mul(X, f2_fma(f2_mul(G, E), f2_c(0.3989422804f), C));
}
CTC_DEVINL uint32_t f2_pack_bf16(F2 v) { float lo, hi; f2_unpack(v, lo, hi); return pack_bf16(lo, hi); }
// bf16x2 pack on the integer pipe (round half away from zero: +0x8000 on the bit pattern, keep the high halves): the
// F2FP conversion shares the 16-lane XU pipe with MUFU, which the GEGLU epilogues saturate
CTC_DEVINL uint32_t pack_bf16_alu(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
CTC_DEVINL float gelu_erf(float x) {
    float cdf, pdf;
    gelu_parts(x, cdf, pdf);
    return x * cdf;
}
CTC_DEVINL float gelu_erf_grad(float x) {
    float cdf, pdf;
    gelu_parts(x, cdf, pdf);
    return fmaf(x, pdf, cdf);
}

// ----------------------------------------------------------------------------------------
// shared-memory address / mbarrier
// ----------------------------------------------------------------------------------------
CTC_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

CTC_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
CTC_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
CTC_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
CTC_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
                 : "memory");
}
CTC_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
CTC_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel aborts with an error) instead of hanging the GPU.
CTC_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
            printf("ctc: mbarrier wait timeout (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}

// ----------------------------------------------------------------------------------------
// cp.async (LDGSTS): 16-byte global -> shared copies that do not occupy registers, so one thread can
// keep dozens of them in flight (memory-level parallelism for the HBM-bound tile loaders)
// ----------------------------------------------------------------------------------------
CTC_DEVINL void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
CTC_DEVINL void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiles, mbarrier completion
// ----------------------------------------------------------------------------------------
CTC_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
CTC_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

CTC_DEVINL void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2,
                            int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
template <uint32_t kCols>
CTC_DEVINL void tmem_alloc(uint32_t* smem_dst) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
CTC_DEVINL void tmem_dealloc(uint32_t taddr) {  // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
CTC_DEVINL void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
CTC_DEVINL void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16/fp16 in, fp32 accumulate), single CTA
// One lane of a CONVERGED warp (elect.sync).  The tcgen05 issue loops run with the whole warp walking the pipeline and
// only the instructions themselves under `if (elect)`: inside an `if (lane == 0)` region the compiler cannot prove that
// descriptors / addresses are warp-uniform and wraps EVERY tcgen05.mma / commit in a scalarisation loop (ELECT,
// 5 x R2UR.BROADCAST, BRA.U.ANY: ~18 SASS instructions per MMA on one thread), which paced the K <= 512 GEMMs and the
// attention kernels; computed in converged code the operands live in uniform registers.
CTC_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
CTC_DEVINL void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
CTC_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers / thread (thread i of the warp = lane base + i)
CTC_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers / thread
CTC_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
CTC_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 64 bf16 (128 B), 8-row swizzle atoms 1024 B apart.
//   bits [0,14)  start address >> 4         bits [16,30) leading byte offset >> 4 (unused for SW128 K-major: 1)
//   bits [32,46) stride byte offset >> 4     bits [46,48) descriptor version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
CTC_DEVINL uint64_t make_umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, shape M x N.
//   [4,6) c_format (1=F32)  [7,10) a_format (1=BF16)  [10,13) b_format  [15] a_major  [16] b_major
//   [17,23) N>>3   [24,29) M>>4
CTC_DEVINL constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a 2-CTA cluster on the two SMs of one TPC issue ONE tcgen05.mma over
// M = 256 rows (128 per CTA); each CTA stages its own A rows and HALF of the B tile, so the operand bytes an SM
// pulls from L2 per MMA drop by a third (the bound of the K <= 512 launches).
// ----------------------------------------------------------------------------------------
CTC_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
CTC_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
CTC_DEVINL uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// Remote arrive on a barrier of the cluster, CTA-scope release (the form CUTLASS uses for its accumulator-empty barrier).
// The explicit `.release.cluster` this used to carry made ptxas put a MEMBAR.ALL.GPU in front of it: every epilogue warp
// of a CTA pair then waited, once per tile, until all of its outstanding global STORES had drained (ncu: 0.7 - 1.7
// membar stalls per issued instruction in the pair kernels) although the barrier only hands a TMEM buffer back to the
// MMA issuer - and the tcgen05.ld of that buffer have completed (tcgen05.wait::ld) before the arrive is issued.
CTC_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <uint32_t kCols>
CTC_DEVINL void tmem_alloc_cg2(uint32_t* smem_dst) {  // one full warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
CTC_DEVINL void tmem_dealloc_cg2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// TMA tile load whose completion is signalled on an mbarrier of EITHER CTA of the pair (bar_cluster_addr is a
// shared::cluster address, e.g. the leader's full barrier)
CTC_DEVINL void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the pair: issued by ONE thread of the leader CTA (rank 0)
CTC_DEVINL void umma_f16_ss_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the issued MMAs have completed
CTC_DEVINL void umma_commit_cg2(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}

// ----------------------------------------------------------------------------------------
// warp-level mma.sync helpers (used by the attention kernels)
// ----------------------------------------------------------------------------------------
CTC_DEVINL void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
CTC_DEVINL void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
// D(16x8, f32) += A(16x16, bf16 row) * B(16x8, bf16 col)
CTC_DEVINL void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace ctc
