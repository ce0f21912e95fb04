// Dense bf16 GEMM family for the CTViT linears:  C[M,N] = A[M,K] · B[N,K]ᵀ  (fp32 accumulate).
//
// sm_100a design: persistent, warp-specialised kernel, one CTA per SM.
//   warp 0      TMA producer   (cp.async.bulk.tensor 2-D, 128B swizzle, mbarrier ring: 4 stages, 6 / 7 in pair mode)
//   warp 1      MMA issuer     (tcgen05.mma kind::f16, M=128, N=BN, K=16; four per 64-deep k-block)
//               both run CONVERGED with one elect.sync lane issuing, so that descriptors / coordinates live in uniform
//               registers (inside an `if (lane == 0)` region every instruction is wrapped in a scalarisation loop)
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-..  epilogue, EW = 8 or 16 warps; a warp may only touch TMEM lanes 32*(warp%4)..+31, so the warps that
//               share a lane quarter split the tile's columns.  Two families:
//               staged  (B as nn.Linear stores it): tcgen05.ld 32x32b -> registers -> fused epilogue -> per-warp
//                        swizzled shared-memory tile -> coalesced 16-byte global accesses
//               direct  (B rows permuted inside 32-row groups at plan time, CTC_GEMM_BPERM): tcgen05.ld 16x256b ->
//                        registers -> fused epilogue -> global, no shared memory (the product path; bit-identical)
// The accumulator lives in TMEM and is double buffered, so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Both operands are K-major ("row-major [rows, K]"), which is exactly the
// nn.Linear weight layout [out_features, in_features] (reference: src/utils/attention.py:47,49,
// 119-120,124; src/utils/ctvit.py:50; src/models/ctclip.py:62-63).
//
// Pair mode (CG = 2, the default for 256-wide tiles): the two CTAs of a 2-CTA cluster sit on the two SMs of a TPC
// and compute ONE 256 x 256 tile with tcgen05.mma.cta_group::2 (M = 256: 128 accumulator rows in each CTA's TMEM).
// Each CTA stages its own 128 A rows and only HALF of the B tile (128 of the 256 weight rows); the tensor core
// reads the other half from the peer's shared memory.  Per SM and 64-deep k-block that is 32 KB from L2 instead
// of 48 KB for the same 4 MMAs — ncu showed the K <= 512 launches (to_q, to_kv, FF1 + GEGLU, dh) bound by exactly
// that L2 -> SM operand stream at 48-54 % tensor-pipe activity (chip-wide L2 read cap / 148 SMs ~ 43 B/clk against
// 94 B/clk demanded) — and the smaller stage buys a 6-deep ring.  Protocol: both producers signal the LEADER's full
// barrier (TMA .cta_group::2 completion on a peer mbarrier), the leader's elected thread issues the MMAs and
// multicasts its commits to the empty / accumulator-full barriers of both CTAs, the peer's epilogue warps arrive
// remotely on the leader's accumulator-empty barrier.
//
// A plain SIMT kernel with the same epilogues (gemm_simt) is kept as a bring-up / cross-check
// comparator for tests; the product path always uses the tcgen05 kernel.
#include <stdlib.h>

#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
static constexpr int kEpiWarps = 8;       // default; the arithmetic-heavy GEGLU epilogue runs with 16 (template parameter EW)

struct GemmArgs {
    int M, N, K;
    void* out;          // bf16 or fp32, row stride ldc (elements)
    long long ldc;
    const float* bias;  // [N] or null
    const float* resid; // fp32 [M, ldr] or null (may alias out)
    long long ldr;
    void* aux;          // EPI_GEGLU: optional adjoint factors [a | b] bf16 [M, N] out; EPI_GEGLU_BWD: the same [M, 2N] in
    long long ldaux;
    float* top2_val;    // EPI_ARGMAX: [M, n_tiles*2]
    int* top2_idx;
    int n_tiles_n;
};

template <int BN, int CG, int EW = kEpiWarps, int DS = 0>
struct GemmSmem {
    static constexpr int kStages = CG == 2 ? (DS ? 7 : 6) : 4;      // direct epilogues need no staging tile: one more stage
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = (BN / CG) * BK * 2;          // pair mode: each CTA stages half of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStageWarp = EW == 16 ? 2048 : 4096;   // epilogue staging per warp: 32 rows x 128 B (64 B with 16 warps)
    static constexpr int kStageOut = DS ? 0 : EW * kStageWarp;
    static constexpr int kOutOffset = kStages * kStageBytes;
    static constexpr int kBarOffset = kOutOffset + kStageOut;
    static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

// ---------------------------------------------------------------------------------------------
// epilogue: one thread owns one accumulator row (TMEM lane) and 32 consecutive columns
// ---------------------------------------------------------------------------------------------
template <int EPI>
CTC_DEVINL void epilogue_store(const GemmArgs& g, int row, int col0, const uint32_t (&acc)[32]) {
    if (row >= g.M) return;
    if constexpr (EPI == CTC_EPI_BF16) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(g.out) + (long long)row * g.ldc + col0;
        if (col0 + 32 <= g.N && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                uint4 q;
                q.x = pack_bf16(__uint_as_float(acc[v * 8 + 0]), __uint_as_float(acc[v * 8 + 1]));
                q.y = pack_bf16(__uint_as_float(acc[v * 8 + 2]), __uint_as_float(acc[v * 8 + 3]));
                q.z = pack_bf16(__uint_as_float(acc[v * 8 + 4]), __uint_as_float(acc[v * 8 + 5]));
                q.w = pack_bf16(__uint_as_float(acc[v * 8 + 6]), __uint_as_float(acc[v * 8 + 7]));
                reinterpret_cast<uint4*>(out)[v] = q;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < g.N) out[j] = __float2bfloat16(__uint_as_float(acc[j]));
        }
    } else {  // fp32 output, optional bias and residual
        float* out = reinterpret_cast<float*>(g.out) + (long long)row * g.ldc + col0;
        const float* res = g.resid ? g.resid + (long long)row * g.ldr + col0 : nullptr;
        const bool vec = (col0 + 32 <= g.N) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                         (!res || (reinterpret_cast<uintptr_t>(res) & 15) == 0);
        if (vec) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                float4 o;
                o.x = __uint_as_float(acc[v * 4 + 0]);
                o.y = __uint_as_float(acc[v * 4 + 1]);
                o.z = __uint_as_float(acc[v * 4 + 2]);
                o.w = __uint_as_float(acc[v * 4 + 3]);
                if (g.bias) {
                    const float4 b = *reinterpret_cast<const float4*>(g.bias + col0 + v * 4);
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                if (res) {
                    const float4 r = reinterpret_cast<const float4*>(res)[v];
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                reinterpret_cast<float4*>(out)[v] = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (col0 + j < g.N) {
                    float o = __uint_as_float(acc[j]);
                    if (g.bias) o += g.bias[col0 + j];
                    if (res) o += res[j];
                    out[j] = o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// staged epilogue: the accumulator arrives one row per thread (TMEM lane), which would make every
// global access a 32-row scatter.  Each epilogue warp therefore bounces its 32 x 128-byte chunk
// through a private, XOR-swizzled (bank-conflict-free) shared-memory tile and then touches global
// memory with one 16-byte access per lane such that a warp instruction covers 4 full 128-byte row
// segments: residual / bias loads and the final stores are fully coalesced.
// ---------------------------------------------------------------------------------------------
CTC_DEVINL uint32_t stage_off(int row, int unit) { return (uint32_t)(row * 128 + ((unit ^ (row & 7)) << 4)); }
// The staging tile is addressed in the shared state space explicitly: through generic pointers the compiler has to
// assume that a global store may alias the next staging load and serialises every load / store pair of the read-out
// (ncu: LD.E.128 -> wait -> STG -> LD ... , `long scoreboard` on each store); with ld.shared the eight loads of a
// read-out are issued back to back.
CTC_DEVINL void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
CTC_DEVINL uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// fp32 chunk: 32 rows x 32 columns
CTC_DEVINL void epilogue_f32_staged(const GemmArgs& g, uint32_t stage, int row0, int col0, int lane,
                                    const uint32_t (&acc)[32], const float4 (&res)[8], bool has_res) {
#pragma unroll
    for (int u = 0; u < 8; ++u) sts128(stage + stage_off(lane, u), acc[4 * u], acc[4 * u + 1], acc[4 * u + 2], acc[4 * u + 3]);
    __syncwarp();
    const int u = lane & 7;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.bias) b = *reinterpret_cast<const float4*>(g.bias + col0 + u * 4);
    uint4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = lds128(stage + stage_off((lane >> 3) + 4 * i, u));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = row0 + (lane >> 3) + 4 * i;
        float4 v = make_float4(__uint_as_float(t[i].x) + b.x, __uint_as_float(t[i].y) + b.y, __uint_as_float(t[i].z) + b.z,
                               __uint_as_float(t[i].w) + b.w);
        if (has_res) { v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w; }
        if (row < g.M)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + (long long)row * g.ldc + col0 + u * 4) = v;
    }
    __syncwarp();
}
CTC_DEVINL void prefetch_resid(const GemmArgs& g, int row0, int col0, int lane, float4 (&res)[8]) {
    const int u = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = row0 + (lane >> 3) + 4 * i;
        res[i] = (row < g.M) ? *reinterpret_cast<const float4*>(g.resid + (long long)row * g.ldr + col0 + u * 4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// bf16 chunk: 32 rows x 64 columns (acc already packed to 32 x bf16x2) -> dst[row, col0 .. col0+64)
CTC_DEVINL void epilogue_bf16_staged(const GemmArgs& g, uint32_t stage, int row0, int col0, int lane,
                                     const uint32_t (&pk)[32], __nv_bfloat16* dst = nullptr, long long ld = 0) {
    if (!dst) { dst = reinterpret_cast<__nv_bfloat16*>(g.out); ld = g.ldc; }
#pragma unroll
    for (int u = 0; u < 8; ++u) sts128(stage + stage_off(lane, u), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
    __syncwarp();
    const int u = lane & 7;
    uint4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = lds128(stage + stage_off((lane >> 3) + 4 * i, u));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = row0 + (lane >> 3) + 4 * i;
        if (row < g.M) *reinterpret_cast<uint4*>(dst + (long long)row * ld + col0 + u * 8) = t[i];
    }
    __syncwarp();
}
// bf16 half chunk: 32 rows x 32 columns (16 x bf16x2 per thread) -> dst[row, col0 .. col0+32); 8 rows x 64 B per access
CTC_DEVINL void epilogue_bf16_staged32(const GemmArgs& g, uint32_t stage, int row0, int col0, int lane,
                                       const uint32_t (&pk)[16], __nv_bfloat16* dst, long long ld) {
#pragma unroll
    for (int u = 0; u < 4; ++u) sts128(stage + stage_off(lane, u), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
    __syncwarp();
    const int u = lane & 3;
    uint4 t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = lds128(stage + stage_off((lane >> 2) + 8 * i, u));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + (lane >> 2) + 8 * i;
        if (row < g.M) *reinterpret_cast<uint4*>(dst + (long long)row * ld + col0 + u * 8) = t[i];
    }
    __syncwarp();
}

// 16-warp epilogues: 2 KB per warp, 32 rows x 64 bytes (four 16-byte units per row).  Thread = row writes its 16 packed
// words, then lane (row = lane / 4 + 8 i, unit = lane % 4) stores unit `unit` of four rows at dst + row * ld + col
// (`col` chosen by the caller per unit: the adjoint factors a | b live 32 columns apart).
CTC_DEVINL uint32_t stage_off64(int row, int unit) { return (uint32_t)(row * 64 + ((unit ^ ((row >> 1) & 3)) << 4)); }
CTC_DEVINL void staged64_store(const GemmArgs& g, uint32_t stage, int row0, int lane, const uint32_t (&pk)[16],
                               __nv_bfloat16* dst, long long ld, int col) {
#pragma unroll
    for (int u = 0; u < 4; ++u) sts128(stage + stage_off64(lane, u), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
    __syncwarp();
    uint4 t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = lds128(stage + stage_off64((lane >> 2) + 8 * i, lane & 3));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + (lane >> 2) + 8 * i;
        if (row < g.M) *reinterpret_cast<uint4*>(dst + (long long)row * ld + col) = t[i];
    }
    __syncwarp();
}

// coalesced global load of a 32-row x 128-byte chunk (src[row, col0 .. col0+64) bf16), split in two halves so that the
// loads of the NEXT chunk are in flight (in registers) while the current chunk is processed:
// issue (global -> registers) ... later ... commit (registers -> swizzled staging tile)
CTC_DEVINL void stage_load_issue(const GemmArgs& g, int row0, int col0, int lane, const __nv_bfloat16* src, long long ld,
                                 uint4 (&reg)[8]) {
    const int u = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = row0 + (lane >> 3) + 4 * i;
        reg[i] = (row < g.M) ? *reinterpret_cast<const uint4*>(src + (long long)row * ld + col0 + u * 8) : make_uint4(0, 0, 0, 0);
    }
}
CTC_DEVINL void stage_load_commit(uint32_t stage, int lane, const uint4 (&reg)[8]) {
    const int u = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) sts128(stage + stage_off((lane >> 3) + 4 * i, u), reg[i].x, reg[i].y, reg[i].z, reg[i].w);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// direct epilogue (DS = 1): no shared-memory staging.
// The staged read-out above costs shared-memory bandwidth the tensor core needs: at K <= 512 a 128 x 256 tile reads
// 384 KB of operands from shared memory and TMA writes the same 384 KB, and the staging added up to 256 KB more per tile
// (timing experiment with the staging accesses removed, tools/gpu/r2j.sh: FF1 + GEGLU 278 -> 222 us, + factors 364 -> 296
// us, dh + adjoint 291 -> 210 us).  tcgen05.ld.16x256b hands a thread two adjacent columns of rows r and r + 8 for every
// 8-column group (the mma.sync C-fragment layout); with the rows of the B operand permuted INSIDE every 32-row group at
// plan time (ctc_gemm_row_perm) those 8 values per row are 8 consecutive output channels, i.e. one 16-byte store per row
// (bf16), and the four threads of a row write 64 contiguous bytes: every store instruction covers eight
// rows x two full 32-byte sectors (fp32: one 256-bit store per row, eight rows x a full 128-byte line).
// Accumulator column a = 8 j + 2 q + e of a 32-column group holds channel 8 q + 2 j + e of the group.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int row_perm(int a) {
    const int j = a >> 3, q = (a >> 1) & 3, e = a & 1;
    return 8 * q + 2 * j + e;
}
// 16 lanes x 32 consecutive fp32 columns: register 4 j + 2 h + e = (lane base + 8 h + thread / 4, column 8 j + 2 (thread % 4) + e)
CTC_DEVINL void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
CTC_DEVINL float asf(uint32_t v) { return __uint_as_float(v); }
// 256-bit global accesses (sm_100: LDG / STG .256): 32 bytes per thread, so that the four threads that share a row in the
// 16x256b register layout cover a full 128-byte line per instruction
CTC_DEVINL void stg256(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
                 "r"(f), "r"(g), "r"(h) : "memory");
}
CTC_DEVINL void ldg256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]),
                 "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}

// taddr: this warp's lane quarter of the accumulator (column 0 of the tile); row0: first of its 32 rows;
// colbase: first output column of the tile; [cbeg, cend): this warp's accumulator columns
template <int EPI>
CTC_DEVINL void epilogue_direct(const GemmArgs& g, uint32_t taddr, int row0, int colbase, int cbeg, int cend, int lane) {
    const int r4 = lane >> 2, q = lane & 3;
    if constexpr (EPI == CTC_EPI_BF16) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(g.out);
#pragma unroll 1
        for (int c = cbeg; c < cend; c += 32) {
            const int col0 = colbase + c;
            if (col0 >= g.N) break;
            uint32_t v[2][16];
            tmem_ld_16x256b_x4(taddr + c, v[0]);
            tmem_ld_16x256b_x4(taddr + (16u << 16) + c, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int L = 0; L < 2; ++L)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = row0 + 16 * L + 8 * h + r4;
                    uint4 o;
                    o.x = pack_bf16(asf(v[L][2 * h]), asf(v[L][2 * h + 1]));
                    o.y = pack_bf16(asf(v[L][4 + 2 * h]), asf(v[L][5 + 2 * h]));
                    o.z = pack_bf16(asf(v[L][8 + 2 * h]), asf(v[L][9 + 2 * h]));
                    o.w = pack_bf16(asf(v[L][12 + 2 * h]), asf(v[L][13 + 2 * h]));
                    if (row < g.M) *reinterpret_cast<uint4*>(out + (long long)row * g.ldc + col0 + 8 * q) = o;
                }
        }
    } else if constexpr (EPI == CTC_EPI_F32) {
        // 8 consecutive fp32 channels per thread and row: one 256-bit load of the residual, one 256-bit store; the four
        // threads of a row cover a full 128-byte line
        float* out = reinterpret_cast<float*>(g.out);
#pragma unroll 1
        for (int c = cbeg; c < cend; c += 32) {
            const int col0 = colbase + c;
            if (col0 >= g.N) break;
            uint32_t res[4][8];
            if (g.resid) {      // in flight while the accumulator chunk is read
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = row0 + 8 * i + r4;
                    if (row < g.M) ldg256(g.resid + (long long)row * g.ldr + col0 + 8 * q, res[i]);
                }
            }
            uint32_t bb[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) bb[k] = 0u;
            if (g.bias) ldg256(g.bias + col0 + 8 * q, bb);
            uint32_t v[2][16];
            tmem_ld_16x256b_x4(taddr + c, v[0]);
            tmem_ld_16x256b_x4(taddr + (16u << 16) + c, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int L = 0; L < 2; ++L)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = 2 * L + h;
                    const int row = row0 + 8 * i + r4;
                    if (row >= g.M) continue;
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        o[2 * j] = asf(v[L][4 * j + 2 * h]) + asf(bb[2 * j]);
                        o[2 * j + 1] = asf(v[L][4 * j + 2 * h + 1]) + asf(bb[2 * j + 1]);
                    }
                    if (g.resid) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) o[k] += asf(res[i][k]);
                    }
                    stg256(out + (long long)row * g.ldc + col0 + 8 * q, __float_as_uint(o[0]), __float_as_uint(o[1]),
                           __float_as_uint(o[2]), __float_as_uint(o[3]), __float_as_uint(o[4]), __float_as_uint(o[5]),
                           __float_as_uint(o[6]), __float_as_uint(o[7]));
                }
        }
    } else if constexpr (EPI == CTC_EPI_GEGLU) {
        __nv_bfloat16* hout = reinterpret_cast<__nv_bfloat16*>(g.out);
        __nv_bfloat16* uout = reinterpret_cast<__nv_bfloat16*>(g.aux);
#pragma unroll 1
        for (int c = cbeg; c < cend; c += 64) {
            const int col0 = colbase + c;
            if (col0 >= g.N) break;
#pragma unroll
            for (int L = 0; L < 2; ++L) {       // 16 rows at a time: bounded live state (16 epilogue warps: 96 registers)
                uint32_t xv[16], gv[16];
                tmem_ld_16x256b_x4(taddr + (uint32_t(16 * L) << 16) + c, xv);
                tmem_ld_16x256b_x4(taddr + (uint32_t(16 * L) << 16) + c + 32, gv);
                tmem_ld_wait();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = row0 + 16 * L + 8 * h + r4;
                    uint32_t hk[4], pa[4], pb[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float x0 = asf(xv[4 * j + 2 * h]), x1 = asf(xv[4 * j + 2 * h + 1]);
                        const float g0 = asf(gv[4 * j + 2 * h]), g1 = asf(gv[4 * j + 2 * h + 1]);
                        F2 hh, aa, bb;
                        if (uout) {
                            geglu_pair2<true>(x0, x1, g0, g1, hh, aa, bb);
                            pa[j] = f2_pack_bf16(aa);
                            pb[j] = f2_pack_bf16(bb);
                        } else {
                            geglu_pair2<false>(x0, x1, g0, g1, hh, aa, bb);
                        }
                        hk[j] = f2_pack_bf16(hh);
                    }
                    if (row < g.M) {
                        *reinterpret_cast<uint4*>(hout + (long long)row * g.ldc + col0 / 2 + 8 * q) = make_uint4(hk[0], hk[1], hk[2], hk[3]);
                        if (uout) {
                            __nv_bfloat16* pu = uout + (long long)row * g.ldaux + col0 + 8 * q;
                            *reinterpret_cast<uint4*>(pu) = make_uint4(pa[0], pa[1], pa[2], pa[3]);
                            *reinterpret_cast<uint4*>(pu + 32) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
                        }
                    }
                }
            }
        }
    } else if constexpr (EPI == CTC_EPI_GEGLU_BWD) {
        const __nv_bfloat16* uin = reinterpret_cast<const __nv_bfloat16*>(g.aux);
        __nv_bfloat16* duout = reinterpret_cast<__nv_bfloat16*>(g.out);
#pragma unroll 1
        for (int c = cbeg; c < cend; c += 32) {
            const int col0 = colbase + c;
            if (col0 >= g.N) break;
            uint4 fa[4], fb[4];     // the saved factors of this thread's 4 rows x 8 channels: in flight during the TMEM read
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = row0 + 8 * i + r4;
                const __nv_bfloat16* pu = uin + (long long)row * g.ldaux + 2 * col0 + 8 * q;
                const bool ok = row < g.M;
                fa[i] = ok ? *reinterpret_cast<const uint4*>(pu) : make_uint4(0, 0, 0, 0);
                fb[i] = ok ? *reinterpret_cast<const uint4*>(pu + 32) : make_uint4(0, 0, 0, 0);
            }
            uint32_t v[2][16];
            tmem_ld_16x256b_x4(taddr + c, v[0]);
            tmem_ld_16x256b_x4(taddr + (16u << 16) + c, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int L = 0; L < 2; ++L)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = 2 * L + h;
                    const int row = row0 + 8 * i + r4;
                    const uint32_t as[4] = {fa[i].x, fa[i].y, fa[i].z, fa[i].w}, bs[4] = {fb[i].x, fb[i].y, fb[i].z, fb[i].w};
                    uint32_t dv[4], dg[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 af = unpack_bf16(as[j]), bf = unpack_bf16(bs[j]);
                        const float d0 = asf(v[L][4 * j + 2 * h]), d1 = asf(v[L][4 * j + 2 * h + 1]);
                        dv[j] = pack_bf16(af.x * d0, af.y * d1);
                        dg[j] = pack_bf16(bf.x * d0, bf.y * d1);
                    }
                    if (row < g.M) {
                        __nv_bfloat16* pd = duout + (long long)row * g.ldc + 2 * col0 + 8 * q;
                        *reinterpret_cast<uint4*>(pd) = make_uint4(dv[0], dv[1], dv[2], dv[3]);
                        *reinterpret_cast<uint4*>(pd + 32) = make_uint4(dg[0], dg[1], dg[2], dg[3]);
                    }
                }
        }
    }
}

// running top-2 (value, column) over a row, used by the VQ nearest-code search
struct Top2 {
    float v0, v1;
    int i0, i1;
    CTC_DEVINL void init() { v0 = v1 = -3.0e38f; i0 = i1 = 0; }
    CTC_DEVINL void push(float v, int i) {   // branch-free: 2 compares + 6 selects
        const bool g0 = v > v0, g1 = v > v1;
        v1 = g0 ? v0 : (g1 ? v : v1);
        i1 = g0 ? i0 : (g1 ? i : i1);
        v0 = g0 ? v : v0;
        i0 = g0 ? i : i0;
    }
};

template <int BN, int EPI, int CG, int EW, int DS>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmArgs g) {
    using S = GemmSmem<BN, CG, EW, DS>;
    constexpr int kStages = S::kStages;
    constexpr int TM = BM * CG;                                   // rows of one tile (pair mode: 256)
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;       // 0 = leader (issues the MMAs)
    const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // tile-scheduling unit: CTA or CTA pair
    const int n_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_m = (g.M + TM - 1) / TM;
    const int tiles_n = (g.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int k_blocks = (g.K + BK - 1) / BK;
    constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512 (power of two >= 32)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EW * CG); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (CG == 2) tmem_alloc_cg2<kTmemCols>(tmem_ptr);
        else tmem_alloc<kTmemCols>(tmem_ptr);
    }
    tcgen05_fence_before();
    if constexpr (CG == 2) cluster_sync_all();      // the peer's barriers must be initialised before anything signals them
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        // whole warp converged, one elected lane issues (see elect_one): addresses / coordinates stay in uniform registers
        {
            const bool issuer = elect_one();
            int stage = 0; uint32_t phase = 0;
            for (int tile = unit; tile < num_tiles; tile += n_units) {
                const int tm = tile / tiles_n, tn = tile % tiles_n;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::kStageBytes;
                    uint8_t* sb = sa + S::kABytes;
                    if constexpr (CG == 2) {
                        // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of both
                        const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if (issuer) {
                            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                            tma_load_2d_cg2(sa, &tmap_a, fb, kb * BK, tm * TM + (int)rank * BM);
                            tma_load_2d_cg2(sb, &tmap_b, fb, kb * BK, tn * BN + (int)rank * (BN / 2));
                        }
                    } else {
                        if (issuer) {
                            mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
                            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, tm * BM);
                            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, tn * BN);
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the pipeline and one elected lane issues (see elect_one): every descriptor is computed in
        // converged code, so the tcgen05 instructions take uniform registers without a scalarisation loop.
        if (rank == 0) {
            const bool issuer = elect_one();
            constexpr uint32_t idesc = make_idesc_bf16(TM, BN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = unit; tile < num_tiles; tile += n_units) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
                    const uint32_t sb = sa + S::kABytes;
                    const uint64_t da = make_umma_desc_sw128(sa);
                    const uint64_t db = make_umma_desc_sw128(sb);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr>>4) field
                            if constexpr (CG == 2)
                                umma_f16_ss_cg2(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                                (kb > 0 || k > 0) ? 1u : 0u);
                            else
                                umma_f16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                            (kb > 0 || k > 0) ? 1u : 0u);
                        }
                        // frees the smem slot (of both CTAs in pair mode) once these MMAs retire
                        if constexpr (CG == 2) umma_commit_cg2(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue (of both CTAs)
                if (issuer) {
                    if constexpr (CG == 2) umma_commit_cg2(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int ew = (warp - 4) & 3;   // TMEM lane quarter: warp (id % 4) may only touch lanes 32*(id%4)..+31
        const int eh = (warp - 4) >> 2;  // which half of the tile's columns this warp owns
        constexpr int kColsPerWarp = BN / (EW / 4);
        const int cbeg = eh * kColsPerWarp, cend = cbeg + kColsPerWarp;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = unit; tile < num_tiles; tile += n_units) {
            const int tm = tile / tiles_n, tn = tile % tiles_n;
            const int rbase = tm * TM + (int)rank * BM;          // first row of this CTA's 128 accumulator rows
            if constexpr (EPI == CTC_EPI_F32) {
                // While this tile's MMAs run, pull the residual rows this warp will add (32 rows x its columns, fp32)
                // into L2: the epilogue's loads then hit L2 instead of waiting on HBM with only 4 KB in flight per warp.
                if (g.resid) {
                    const int prow = rbase + ew * 32 + lane;
                    const int pcol = tn * BN + cbeg;
                    if (prow < g.M) {
                        const float* pr = g.resid + (long long)prow * g.ldr + pcol;
#pragma unroll
                        for (int c = 0; c < kColsPerWarp; c += 32)
                            if (pcol + c < g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + c));
                    }
                }
            }
            if constexpr (EPI == CTC_EPI_GEGLU_BWD) {
                // same idea for the saved pre-activation u the adjoint reads (bf16, 2 x this warp's columns per row)
                const int prow = rbase + ew * 32 + lane;
                const int pcol = 2 * (tn * BN + cbeg);
                if (prow < g.M) {
                    const __nv_bfloat16* pu = reinterpret_cast<const __nv_bfloat16*>(g.aux) + (long long)prow * g.ldaux + pcol;
#pragma unroll
                    for (int c = 0; c < 2 * kColsPerWarp; c += 64)
                        if (pcol + c < 2 * g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pu + c));
                }
            }
            uint4 ureg[(EPI == CTC_EPI_GEGLU_BWD && DS == 0) ? 8 : 1];
            if constexpr (EPI == CTC_EPI_GEGLU_BWD && DS == 0) {
                // first chunk of the saved adjoint factors: in registers before the accumulator is even complete
                if (tn * BN + cbeg < g.N)
                    stage_load_issue(g, rbase + ew * 32, 2 * (tn * BN + cbeg), lane,
                                     reinterpret_cast<const __nv_bfloat16*>(g.aux), g.ldaux, ureg);
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const int row = rbase + ew * 32 + lane;
            const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;
            const uint32_t stage = smem_u32(smem + S::kOutOffset + (warp - 4) * S::kStageWarp);
            const int row0 = rbase + ew * 32;
            if constexpr (DS == 1) {
                epilogue_direct<EPI>(g, taddr, row0, tn * BN, cbeg, cend, lane);
            } else if constexpr (EPI == CTC_EPI_ARGMAX) {
                // four independent trackers break the 256-long dependent compare chain
                Top2 t2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) t2[i].init();
                const bool full = (tn + 1) * BN <= g.N;
#pragma unroll 1
                for (int c = cbeg; c < cend; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c, v);
                    tmem_ld_wait();
                    const int col0 = tn * BN + c;
                    // groups of 4 consecutive codes: only the group maximum enters the top-2 tracker (11 ALU ops per
                    // 4 scores instead of 32); the recorded index is the group's first code and vq_refine re-scores all
                    // four codes of a candidate group in fp32
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float m;
                        if (full || col0 + j + 3 < g.N) {
                            m = fmaxf(fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                      fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
                        } else {
                            m = -3.0e38f;
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (col0 + j + e < g.N) m = fmaxf(m, __uint_as_float(v[j + e]));
                        }
                        t2[(j >> 2) & 3].push(m, col0 + j);
                    }
                }
#pragma unroll
                for (int i = 1; i < 4; ++i) { t2[0].push(t2[i].v0, t2[i].i0); t2[0].push(t2[i].v1, t2[i].i1); }
                if (row < g.M) {
                    const long long o = (((long long)row * g.n_tiles_n + tn) * (EW / 4) + eh) * 2;
                    g.top2_val[o] = t2[0].v0; g.top2_val[o + 1] = t2[0].v1;
                    g.top2_idx[o] = t2[0].i0; g.top2_idx[o + 1] = t2[0].i1;
                }
            } else if constexpr (EPI == CTC_EPI_GEGLU) {
                // columns come in 64-wide groups [32 value | 32 gate] (weights interleaved at plan time):
                // h = gelu(gate) * value (attention.py:38-41) straight from the fp32 accumulators.  For the backward
                // pass the epilogue saves, in the same [32 | 32] layout, the two ADJOINT FACTORS instead of the
                // pre-activation:  a = gelu(gate) = d h / d value,  b = value * gelu'(gate) = d h / d gate  (one exp
                // gives both the cdf and the pdf), so that the adjoint is two multiplies per element and fits the
                // dh GEMM's epilogue (EPI_GEGLU_BWD) instead of a stand-alone HBM pass over u, dh and du.
                __nv_bfloat16* hout = reinterpret_cast<__nv_bfloat16*>(g.out);
                __nv_bfloat16* uout = reinterpret_cast<__nv_bfloat16*>(g.aux);
                if constexpr (EW == 16) {
                    // 16 epilogue warps (ncu on the 8-warp version: 0.44 IPC per scheduler, XU pipe 23 %, tensor pipe 45 %:
                    // two resident warps per scheduler cannot hide the MUFU / TMEM / staging latencies of this epilogue).
                    // Each warp owns ONE 64-column group of the tile and walks it in two half passes of 16 (value, gate)
                    // pairs, which keeps the live state under the 96 registers a 640-thread CTA leaves per thread.
                    static_assert(EW != 16 || BN == 256, "16 epilogue warps: one [32 value | 32 gate] group per warp");
                    const int col0 = tn * BN + cbeg;
                    if (col0 < g.N) {
                        const int un = lane & 3;
                        uint32_t hk[16];
#pragma unroll
                        for (int hp = 0; hp < 2; ++hp) {
                            uint32_t xv[16], gv[16], pk[16];
                            tmem_ld_32x32b_x16(taddr + cbeg + hp * 16, xv);
                            tmem_ld_32x32b_x16(taddr + cbeg + 32 + hp * 16, gv);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                F2 hh, aa, bb;
                                geglu_pair2<true>(__uint_as_float(xv[2 * j]), __uint_as_float(xv[2 * j + 1]), __uint_as_float(gv[2 * j]),
                                                  __uint_as_float(gv[2 * j + 1]), hh, aa, bb);
                                hk[hp * 8 + j] = f2_pack_bf16(hh);
                                pk[j] = f2_pack_bf16(aa);
                                pk[8 + j] = f2_pack_bf16(bb);
                            }
                            // units 0, 1 of a staged row = a (16 columns), units 2, 3 = b: 32 columns further in [a | b]
                            if (uout) staged64_store(g, stage, row0, lane, pk, uout, g.ldaux, col0 + hp * 16 + (un & 1) * 8 + (un >> 1) * 32);
                        }
                        staged64_store(g, stage, row0, lane, hk, hout, g.ldc, col0 / 2 + un * 8);
                    }
                } else
#pragma unroll 1
                for (int c = cbeg; c < cend; c += 64) {
                    const int col0 = tn * BN + c;
                    if (col0 >= g.N) break;
                    uint32_t xv[32], gv[32];
                    tmem_ld_32x32b_x32(taddr + c, xv);
                    tmem_ld_32x32b_x32(taddr + c + 32, gv);
                    tmem_ld_wait();
                    uint32_t hk[16];
                    if (uout) {
                        uint32_t pk[32];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            F2 hh, aa, bb;
                            geglu_pair2<true>(__uint_as_float(xv[2 * j]), __uint_as_float(xv[2 * j + 1]), __uint_as_float(gv[2 * j]),
                                              __uint_as_float(gv[2 * j + 1]), hh, aa, bb);
                            hk[j] = f2_pack_bf16(hh);
                            pk[j] = f2_pack_bf16(aa);
                            pk[16 + j] = f2_pack_bf16(bb);
                        }
                        epilogue_bf16_staged(g, stage, row0, col0, lane, pk, uout, g.ldaux);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            // same arithmetic as the branch above, so that h does not depend on whether u is saved
                            F2 hh, aa, bb;
                            geglu_pair2<false>(__uint_as_float(xv[2 * j]), __uint_as_float(xv[2 * j + 1]), __uint_as_float(gv[2 * j]),
                                               __uint_as_float(gv[2 * j + 1]), hh, aa, bb);
                            hk[j] = f2_pack_bf16(hh);
                        }
                    }
                    epilogue_bf16_staged32(g, stage, row0, col0 / 2, lane, hk, hout, g.ldc);
                }
            } else if constexpr (EPI == CTC_EPI_GEGLU_BWD) {
                // acc = dh chunk (32 columns); aux = the saved adjoint factors [a | b] of the same columns, one 128-byte
                // row segment: du_value = a * dh, du_gate = b * dh, written back in the same grouped layout
                const __nv_bfloat16* uin = reinterpret_cast<const __nv_bfloat16*>(g.aux);
                __nv_bfloat16* duout = reinterpret_cast<__nv_bfloat16*>(g.out);
#pragma unroll 1
                for (int c = cbeg; c < cend; c += 32) {
                    const int col0 = tn * BN + c;
                    if (col0 >= g.N) break;
                    stage_load_commit(stage, lane, ureg);
                    uint32_t dh[32];
                    tmem_ld_32x32b_x32(taddr + c, dh);
                    uint4 aa4[4], ba4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        aa4[u] = lds128(stage + stage_off(lane, u));
                        ba4[u] = lds128(stage + stage_off(lane, u + 4));
                    }
                    // next chunk's factors on their way while this one is multiplied and stored
                    if (c + 32 < cend && col0 + 32 < g.N) stage_load_issue(g, row0, 2 * (col0 + 32), lane, uin, g.ldaux, ureg);
                    tmem_ld_wait();
                    uint32_t pk[32];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint4 aa = aa4[u];
                        const uint4 ba = ba4[u];
                        const uint32_t as[4] = {aa.x, aa.y, aa.z, aa.w}, bs[4] = {ba.x, ba.y, ba.z, ba.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 af = unpack_bf16(as[e]), bf = unpack_bf16(bs[e]);
                            const float d0 = __uint_as_float(dh[u * 8 + 2 * e]), d1 = __uint_as_float(dh[u * 8 + 2 * e + 1]);
                            pk[u * 4 + e] = pack_bf16(af.x * d0, af.y * d1);
                            pk[16 + u * 4 + e] = pack_bf16(bf.x * d0, bf.y * d1);
                        }
                    }
                    __syncwarp();
                    epilogue_bf16_staged(g, stage, row0, 2 * col0, lane, pk, duout, g.ldc);
                }
            } else if constexpr (EPI == CTC_EPI_F32) {
                const bool aligned = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.out) & 15) == 0) &&
                                     (!g.resid || ((g.ldr % 4 == 0) && (reinterpret_cast<uintptr_t>(g.resid) & 15) == 0)) &&
                                     (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
#pragma unroll 1
                for (int c = cbeg; c < cend; c += 32) {
                    const int col0 = tn * BN + c;
                    if (col0 >= g.N) break;
                    const bool fast = aligned && (col0 + 32 <= g.N);
                    float4 res[8];
                    if (fast && g.resid) prefetch_resid(g, row0, col0, lane, res);
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c, v);
                    tmem_ld_wait();
                    if (fast) epilogue_f32_staged(g, stage, row0, col0, lane, v, res, g.resid != nullptr);
                    else epilogue_store<EPI>(g, row, col0, v);
                }
            } else {
                const bool aligned = (g.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(g.out) & 15) == 0);
#pragma unroll 1
                for (int c = cbeg; c < cend; c += 64) {
                    const int col0 = tn * BN + c;
                    if (col0 >= g.N) break;
                    uint32_t v[32], pk[32];
                    tmem_ld_32x32b_x32(taddr + c, v);
                    tmem_ld_wait();
                    if (aligned && col0 + 64 <= g.N) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            pk[j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                        tmem_ld_32x32b_x32(taddr + c + 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            pk[16 + j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                        epilogue_bf16_staged(g, stage, row0, col0, lane, pk);
                    } else {
                        epilogue_store<EPI>(g, row, col0, v);
                        if (col0 + 32 < g.N) {
                            tmem_ld_32x32b_x32(taddr + c + 32, v);
                            tmem_ld_wait();
                            epilogue_store<EPI>(g, row, col0 + 32, v);
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                // pair mode: the leader's MMA issuer waits for the epilogue warps of BOTH CTAs
                if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tcgen05_fence_before();
    if constexpr (CG == 2) cluster_sync_all();      // neither CTA may leave while the pair's MMAs / remote arrives are in flight
    else __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        if constexpr (CG == 2) tmem_dealloc_cg2<kTmemCols>(tmem_base);
        else tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// SIMT comparator (tests / bring-up only)
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, long long lda,
                                 const __nv_bfloat16* __restrict__ B, long long ldb, const GemmArgs g, int perm_kind) {
    __shared__ float sa[16][17], sb[16][17];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.y * 16 + ty;
    int col = blockIdx.x * 16 + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < g.K; k0 += 16) {
        const int ka = k0 + tx;
        sa[ty][tx] = (row < g.M && ka < g.K) ? __bfloat162float(A[(long long)row * lda + ka]) : 0.f;
        const int brow = blockIdx.x * 16 + ty;
        sb[ty][tx] = (brow < g.N && ka < g.K) ? __bfloat162float(B[(long long)brow * ldb + ka]) : 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sb[tx][k];
        __syncthreads();
    }
    if (row >= g.M || col >= g.N) return;
    // B rows permuted for the direct epilogues: row (col) of B holds the weights of output channel row_perm(col)
    const int col_b = col;
    (void)col_b;
    if (perm_kind) col = (col & ~31) + row_perm(col & 31);
    if constexpr (EPI == CTC_EPI_BF16) {
        reinterpret_cast<__nv_bfloat16*>(g.out)[(long long)row * g.ldc + col] = __float2bfloat16(acc);
    } else {
        if (g.bias) acc += g.bias[col];
        if (g.resid) acc += g.resid[(long long)row * g.ldr + col];
        reinterpret_cast<float*>(g.out)[(long long)row * g.ldc + col] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row stride ld (elements); box = [box_rows, 64]
static int make_tmap_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld,
                          int box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    CTC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    CTC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand not 16-byte aligned");
    CTC_REQUIRE((ld * 2) % 16 == 0, "GEMM operand row stride (%lld elements) is not a multiple of 16 bytes", ld);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

static int g_num_sms[kMaxDevices] = {};
int num_sms() {
    const int dev = current_device();
    if (!g_num_sms[dev]) {
        cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms[dev] <= 0) g_num_sms[dev] = 148;
    }
    return g_num_sms[dev];
}

template <int BN, int EPI, int CG, int EW = kEpiWarps, int DS = 0>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, cudaStream_t st) {
    using S = GemmSmem<BN, CG, EW, DS>;
    constexpr int kGemmThreads = 128 + 32 * EW;
    static bool configured_dev[kMaxDevices] = {};
    bool& configured = configured_dev[current_device()];
    if (!configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, EPI, CG, EW, DS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            S::kTotal));
        configured = true;
    }
    const int tiles = ((g.M + BM * CG - 1) / (BM * CG)) * ((g.N + BN - 1) / BN);
    const int units = num_sms() / CG;                       // CTAs, or CTA pairs (one pair per TPC)
    const int grid = (tiles < units ? tiles : units) * CG;
    if constexpr (CG == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = S::kTotal; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CTC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, EPI, CG, EW, DS>, ta, tb, g));
    } else {
        gemm_tcgen05_kernel<BN, EPI, CG, EW, DS><<<grid, kGemmThreads, S::kTotal, st>>>(ta, tb, g);
    }
    CTC_LAUNCH_CHECK();
    return 0;
}

int gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* out, long long ldc, int M, int N,
              int K, int epi, const float* bias, const float* resid, long long ldr, void* aux, long long ldaux,
              float* top2_val, int* top2_idx, int impl, cudaStream_t st) {
    CTC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    // B rows permuted inside every 32-row group (ctc_gemm_row_perm) -> the staging-free direct epilogue
    const int perm_kind = (impl & CTC_GEMM_BPERM) ? 1 : 0;
    impl &= 0xff;
    if (perm_kind) {
        CTC_REQUIRE(epi != CTC_EPI_ARGMAX && N % 32 == 0, "gemm: permuted B rows need N %% 32 == 0 (N=%d) and a storing epilogue", N);
        const bool f32 = epi == CTC_EPI_F32;   // fp32 rows move as 256-bit accesses
        CTC_REQUIRE(ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & (f32 ? 31 : 15)) == 0 &&
                    (!resid || (ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(resid) & 31) == 0)) &&
                    (!bias || (reinterpret_cast<uintptr_t>(bias) & 31) == 0),
                    "gemm: the direct epilogue needs %d-byte aligned rows (ldc=%lld)", f32 ? 32 : 16, ldc);
    }
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K; g.out = out; g.ldc = ldc; g.bias = bias; g.resid = resid; g.ldr = ldr;
    g.top2_val = top2_val; g.top2_idx = top2_idx; g.aux = aux; g.ldaux = ldaux;
    if (epi == CTC_EPI_GEGLU || epi == CTC_EPI_GEGLU_BWD) {
        CTC_REQUIRE(impl != CTC_GEMM_SIMT, "gemm: the fused GEGLU epilogues exist only in the tcgen05 kernels");
        CTC_REQUIRE(N % 64 == 0 && ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                    (!aux || (ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0)),
                    "gemm: GEGLU epilogues need N %% 64 == 0 and 16-byte aligned bf16 rows (N=%d)", N);
        CTC_REQUIRE(epi != CTC_EPI_GEGLU_BWD || aux, "gemm: GEGLU backward epilogue needs the saved pre-activation u");
    }
    if (impl == CTC_GEMM_SIMT) {
        CTC_REQUIRE(epi == CTC_EPI_BF16 || epi == CTC_EPI_F32, "gemm: SIMT comparator has only the plain epilogues");
        dim3 grid((N + 15) / 16, (M + 15) / 16), block(16, 16);
        if (epi == CTC_EPI_BF16)
            gemm_simt_kernel<CTC_EPI_BF16><<<grid, block, 0, st>>>((const __nv_bfloat16*)A, lda,
                                                                   (const __nv_bfloat16*)B, ldb, g, perm_kind);
        else
            gemm_simt_kernel<CTC_EPI_F32><<<grid, block, 0, st>>>((const __nv_bfloat16*)A, lda,
                                                                  (const __nv_bfloat16*)B, ldb, g, perm_kind);
        CTC_LAUNCH_CHECK();
        return 0;
    }
    // tile width: 128x256 tiles; 128x128 only for narrow outputs (N < 512) that divide by 128 but not by 256.
    // (N = 1408, the padded FF inner dim: 256-wide tiles with a half-empty last tile measured 158 us against
    // 191 us for 128-wide tiles - twice the MMA work per byte staged through shared memory.)
    const bool bn256 = (epi == CTC_EPI_ARGMAX) || (N % 256 == 0) || (N % 128 != 0) || (N >= 512);
    const int BNsel = bn256 ? 256 : 128;
    // CTA pairs (cta_group::2, 256 x 256 tiles) for every 256-wide case unless the caller asks for the single-CTA kernel
    // (CTC_GEMM_PAIR=0 in the environment switches the default back to single CTAs: an A/B measurement aid)
    static const int pair_env = [] { const char* e = getenv("CTC_GEMM_PAIR"); return e ? atoi(e) : 1; }();
    // Measured on B200 at M = 110 592 (tools/kernel_bench.py, profiles/r02_kernel_bench_v3.log).  Until the remote
    // accumulator-empty arrive lost its GPU-scope fence (mbar_arrive_cluster) pairs only won where K >= 512 and N >= 512
    // and lost under the GEGLU epilogue; since then they win on every shape of the step: to_q 34.3 -> 32.3 us,
    // to_kv 60.0 -> 50.5, FF1 + GEGLU 265 -> 255 (+ factors 348 -> 307), dh 153 -> 137, patch embedding 390 -> 326 and its
    // adjoint 462 -> 347 us (cuBLAS, bare GEMM: 35.9 / 55.2 / 265 / 145 / 315 / 461 us).
    // impl = CTC_GEMM_TCGEN05_PAIR / _1CTA force either kernel (tests, A/B measurement).
    static const int argmax_pair_env = [] { const char* e = getenv("CTC_GEMM_ARGMAX_PAIR"); return e ? atoi(e) : 0; }();
    const bool want_pair = impl == CTC_GEMM_TCGEN05_PAIR || (pair_env && (epi != CTC_EPI_ARGMAX || argmax_pair_env));
    const bool pair = bn256 && impl != CTC_GEMM_TCGEN05_1CTA && want_pair && (num_sms() % 2 == 0);
    // 16 epilogue warps for the GEGLU epilogue (CTC_GEMM_EW16=0 switches back to 8: an A/B measurement aid)
    static const int ew16_env = [] { const char* e = getenv("CTC_GEMM_EW16"); return e ? atoi(e) : 1; }();
    g.n_tiles_n = (N + BNsel - 1) / BNsel;
    CUtensorMap ta, tb;
    if (int e = make_tmap_bf16(&ta, A, M, K, lda, BM)) return e;
    if (int e = make_tmap_bf16(&tb, B, N, K, ldb, pair ? BNsel / 2 : BNsel)) return e;
#define CTC_GEMM_DISPATCH(EPI)                                                                                   \
    if (perm_kind)                                                                                               \
        return pair ? launch_tc<256, EPI, 2, kEpiWarps, 1>(ta, tb, g, st)                                        \
                    : (bn256 ? launch_tc<256, EPI, 1, kEpiWarps, 1>(ta, tb, g, st)                               \
                             : launch_tc<128, EPI, 1, kEpiWarps, 1>(ta, tb, g, st));                             \
    return pair ? launch_tc<256, EPI, 2>(ta, tb, g, st)                                                          \
                : (bn256 ? launch_tc<256, EPI, 1>(ta, tb, g, st) : launch_tc<128, EPI, 1>(ta, tb, g, st))
    switch (epi) {
        case CTC_EPI_BF16: CTC_GEMM_DISPATCH(CTC_EPI_BF16);
        case CTC_EPI_F32: CTC_GEMM_DISPATCH(CTC_EPI_F32);
        case CTC_EPI_GEGLU:
            if (bn256 && ew16_env) {       // 16 epilogue warps (the staged pair kernel keeps 8: 6 stages + 16 x 4 KB do not fit)
                if (perm_kind) return pair ? launch_tc<256, CTC_EPI_GEGLU, 2, 16, 1>(ta, tb, g, st)
                                           : launch_tc<256, CTC_EPI_GEGLU, 1, 16, 1>(ta, tb, g, st);
                if (!pair) return launch_tc<256, CTC_EPI_GEGLU, 1, 16>(ta, tb, g, st);
            }
            CTC_GEMM_DISPATCH(CTC_EPI_GEGLU);
        case CTC_EPI_GEGLU_BWD: CTC_GEMM_DISPATCH(CTC_EPI_GEGLU_BWD);
        case CTC_EPI_ARGMAX:
            CTC_REQUIRE(top2_val && top2_idx, "gemm: arg-max epilogue needs top2 buffers");
            return pair ? launch_tc<256, CTC_EPI_ARGMAX, 2>(ta, tb, g, st) : launch_tc<256, CTC_EPI_ARGMAX, 1>(ta, tb, g, st);
        default:
            CTC_REQUIRE(false, "gemm: unknown epilogue %d", epi);
    }
#undef CTC_GEMM_DISPATCH
    return 0;
}

// perm[a] = output channel (within its 32-channel group) held by accumulator column a: B row (G*32 + a) := W[G*32 + perm[a]]
int gemm_row_perm(int* perm32) {
    for (int a = 0; a < 32; ++a) perm32[a] = row_perm(a);
    return 0;
}

// candidates per row left by the ARGMAX epilogue: top-2 of every (256 / halves)-code slice
int gemm_argmax_candidates(int N) { return ((N + 255) / 256) * (kEpiWarps / 4) * 2; }

}  // namespace ctc
