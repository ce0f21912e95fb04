// Helpers shared by the tcgen05 / TMEM attention kernels (attention_tc.cu: forward + dQ; attention_tc_bwd.cu: the
// one-pass backward).
#pragma once
#include "attention_common.cuh"

namespace ctc {

static constexpr int TC_M = 128;                           // query rows of an M-tile = the 128 TMEM lanes

CTC_DEVINL uint64_t make_umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;             // leading byte offset: unused for swizzled K-major
    d |= static_cast<uint64_t>(512 >> 4) << 32;      // stride byte offset: 8 rows x 64 B
    d |= static_cast<uint64_t>(1) << 46;             // descriptor version (sm_100)
    d |= static_cast<uint64_t>(4) << 61;             // SWIZZLE_64B
    return d;
}
// D[tmem] (+)= A[tmem] * B[smem]
CTC_DEVINL void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
CTC_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
CTC_DEVINL void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
CTC_DEVINL void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
CTC_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// bf16x2 pack of two POSITIVE finite floats on the integer pipe (round half up: add 0x8000, keep the high halves),
// keeping the conversion off the XU pipe that the exponentials saturate
CTC_DEVINL uint32_t pack_bf16_rn_alu(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
// 2^x for x in (-125, 0] on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 magic
// constant, degree-4 polynomial for 2^f on [-0.5, 0.5] (relative error 4e-5, an order of magnitude below the bf16
// rounding P receives anyway), exponent patched in with an integer add.  9 FMA/ALU-pipe instructions; used for half of
// the exponentials of a tile so that the 16-per-clock MUFU unit and the FMA pipes share the softmax (the split FA4
// uses on Blackwell, where d_head-sized tiles are exp-bound, not MMA-bound).
CTC_DEVINL float exp2_poly(float x) {
    const float t = x + 12582912.f;                       // integer part in the low mantissa bits
    const float f = x - (t - 12582912.f);                 // [-0.5, 0.5]
    float p = fmaf(f, 9.6181291e-3f, 5.5504109e-2f);
    p = fmaf(f, p, 2.4022651e-1f);
    p = fmaf(f, p, 6.9314718e-1f);
    p = fmaf(f, p, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));    // t's low bits = n (two's complement)
}
CTC_DEVINL void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// normalise rows [row0, row0 + 128) of q into a SWIZZLE_64B tile; `nthreads` threads starting at `tid0` cooperate
CTC_DEVINL void tc_load_q(uint8_t* tile, const AttnParams& p, int s, int head, int row0, const float* sv, int tid,
                          int nthreads) {
    for (int r = tid; r < TC_M; r += nthreads) {
        const int i = row0 + r;
        uint4 c[4];
        if (i < p.n) {
            const uint4* g = reinterpret_cast<const uint4*>(p.q + seq_row(p, s, i) * p.ldq + head * DH);
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = g[j];
            float f[32];
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t w[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 t = unpack_bf16(w[e]);
                    f[j * 8 + e * 2] = t.x; f[j * 8 + e * 2 + 1] = t.y;
                    ss += t.x * t.x + t.y * t.y;
                }
            }
            const float inv = p.scale * LOG2E / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c[j].x = pack_bf16(f[j * 8 + 0] * inv * sv[j * 8 + 0], f[j * 8 + 1] * inv * sv[j * 8 + 1]);
                c[j].y = pack_bf16(f[j * 8 + 2] * inv * sv[j * 8 + 2], f[j * 8 + 3] * inv * sv[j * 8 + 3]);
                c[j].z = pack_bf16(f[j * 8 + 4] * inv * sv[j * 8 + 4], f[j * 8 + 5] * inv * sv[j * 8 + 5]);
                c[j].w = pack_bf16(f[j * 8 + 6] * inv * sv[j * 8 + 6], f[j * 8 + 7] * inv * sv[j * 8 + 7]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(tile + tile_off(r, j)) = c[j];
    }
}


}  // namespace ctc
