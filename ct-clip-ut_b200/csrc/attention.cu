// Fused cosine-similarity attention for the factorised CTViT transformer
// (reference: Attention.forward, src/utils/attention.py:144-180):
//     q^ = l2norm(q) * q_scale,  k^ = l2norm(k) * k_scale
//     P  = softmax(scale * q^ k^T + bias),   O = P v
// One kernel family serves both sequence modes (ctvit.py:94-101) without ever materialising the
// '(b t) (h w) d' <-> '(b h w) t d' rearranges: SPATIAL sequences are the H*W tokens of a (b,t)
// slice, TEMPORAL sequences are the T tokens of a (b,h,w) column, addressed by stride.
//
// Shapes on this path: d_head = 32, n = 576 (spatial) or 24 (temporal), 8 heads.  The whole K/V of
// one (sequence, head) fits shared memory (36 KB each at n = 576), so there is no K/V streaming:
// a CTA loads + normalises K/V once, each warp owns 16 query rows and walks the keys in blocks
// of 64 with an online softmax.  The l2norm, q/k scales, softmax scale and the relative-position
// bias (a (2H-1)(2W-1) table per head held in shared memory instead of the reference's
// [heads, n, n] tensor) are all fused.  Matrix products use warp-level mma.sync m16n8k16 bf16
// with fp32 accumulation (the d=32 core is exp/LDS-bound, not MMA-bound: SURVEY §7 hard part 3).
//
// Backward (input gradients only) recomputes P from the saved row log-sum-exp:
//   kernel dQ  : per 64-query block  — dP = dO V^T, dS = P∘(dP − D), dQ^ = dS K^, l2norm/scale adjoint
//   kernel dKV : per 64-key block    — the transposed problem for dV = P^T dO, dK^ = dS^T Q^
#include <stdlib.h>

#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {

static constexpr int DH = 32;          // head dim (fixed on this path: inference_ctclip.py:29)
// Tile configurations <QB rows per (CTA, head), KBLK keys per softmax step, HPC heads per CTA>:
//   spatial  (n = 576): one head per CTA, large row tiles so that 12 (fwd) / 6 (bwd) warps share one
//                       resident K/V copy (occupancy is bounded by the 72 KB K/V tile, not by threads);
//   temporal (n = 24) : forward / backward run on the warp-autonomous kernels further down; the CTA-per-sequence
//                       configuration <32, 32, 2> only materialises probabilities (attention_probs).
static constexpr float LOG2E = 1.4426950408889634f;
static constexpr float LN2 = 0.6931471805599453f;

// resident CTAs per SM the register allocation must allow: small CTAs (temporal sequences) rely on several
// independent CTAs per SM to overlap their load / compute phases
constexpr int attn_min_blocks(int threads, int two_block_limit) {
    return threads <= 64 ? 8 : (threads <= 128 ? 4 : (threads <= two_block_limit ? 2 : 1));
}

struct AttnParams {
    const __nv_bfloat16* q; long long ldq;
    const __nv_bfloat16* k; const __nv_bfloat16* v; long long ldkv;
    const __nv_bfloat16* o; const __nv_bfloat16* d_o;   // [R, heads*32]
    const float* q_scale; const float* k_scale; float scale;
    const float* bias_table;  // [heads, (2H-1)*(2W-1)] or null
    int n, n_pad, n_seq, heads, mode, T, HW, H, W;
    __nv_bfloat16* out; float* lse;                      // fwd outputs
    float* probs;                                        // probs kernel output
    __nv_bfloat16* dq; long long lddq; __nv_bfloat16* dk; __nv_bfloat16* dv; long long lddkv;
    float* delta;                                        // [R, heads]
};

// MUFU.EX2 directly (fast_exp2() without fast-math adds denormal range handling around it)
CTC_DEVINL float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

CTC_DEVINL long long seq_row(const AttnParams& p, int s, int i) {
    if (p.mode == CTC_MODE_SPATIAL) return (long long)s * p.HW + i;
    const int b = s / p.HW, hw = s % p.HW;
    return ((long long)b * p.T + i) * p.HW + hw;
}
// 64-byte rows (32 bf16), 16-byte chunks XOR-swizzled so that ldmatrix is bank-conflict free
CTC_DEVINL uint32_t tile_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// Load `rows` rows of 32 bf16 for each of `hpc` heads (row i of the sequence at src + seq_row*ld + head*32) into
// per-head swizzled tiles (tile + hl*tile_stride).  NORM: l2-normalise and multiply by vec[d] * mul (fp32).
// Rows >= n are zero-filled.
template <bool NORM>
CTC_DEVINL void load_tile(uint8_t* tile, int tile_stride, const __nv_bfloat16* src, long long ld, const AttnParams& p,
                          int s, int head0, int hpc, int row0, int rows, const float* vec, float mul) {
    for (int idx = threadIdx.x; idx < rows * hpc; idx += blockDim.x) {
        const int r = idx / hpc, hl = idx - r * hpc;       // consecutive threads -> consecutive heads of one row
        const int i = row0 + r;
        uint4 c[4];
        if (i < p.n) {
            const uint4* g = reinterpret_cast<const uint4*>(src + seq_row(p, s, i) * ld + (head0 + hl) * DH);
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = g[j];
            if (NORM) {
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
                const float inv = mul / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    c[j].x = pack_bf16(f[j * 8 + 0] * inv * vec[j * 8 + 0], f[j * 8 + 1] * inv * vec[j * 8 + 1]);
                    c[j].y = pack_bf16(f[j * 8 + 2] * inv * vec[j * 8 + 2], f[j * 8 + 3] * inv * vec[j * 8 + 3]);
                    c[j].z = pack_bf16(f[j * 8 + 4] * inv * vec[j * 8 + 4], f[j * 8 + 5] * inv * vec[j * 8 + 5]);
                    c[j].w = pack_bf16(f[j * 8 + 6] * inv * vec[j * 8 + 6], f[j * 8 + 7] * inv * vec[j * 8 + 7]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = make_uint4(0, 0, 0, 0);
        }
        uint8_t* t = tile + (long long)hl * tile_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(t + tile_off(r, j)) = c[j];
    }
}

// A fragments (16 rows x 32 dims = 2 k-steps) of the warp's row block from a swizzled tile
CTC_DEVINL void load_a_frags(uint32_t (&a)[2][4], uint32_t tile_addr, int row0, int lane) {
    const int r = row0 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ldmatrix_x4(a[ks], tile_addr + tile_off(r, ks * 2 + (lane >> 4)));
}
// acc(16 x 8) += A(16 x 32) * B^T where B rows (8 of them, starting at row n0) are "n rows x 32 k"
CTC_DEVINL void mma_rowsB(float (&acc)[4], const uint32_t (&a)[2][4], uint32_t tile_addr, int n0, int lane) {
    uint32_t b[4];
    ldmatrix_x4(b, tile_addr + tile_off(n0 + (lane & 7), lane >> 3));
    mma_bf16_16816(acc, a[0], b[0], b[1]);
    mma_bf16_16816(acc, a[1], b[2], b[3]);
}
// acc[4](16 x 32) += A(16 x 16) * B where B rows k0..k0+15 are "k rows x 32 n" (transposed load)
CTC_DEVINL void mma_colsB(float (&acc)[4][4], const uint32_t (&a)[4], uint32_t tile_addr, int k0, int lane) {
    const int r = k0 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, tile_addr + tile_off(r, dp * 2 + (lane >> 4)));
        mma_bf16_16816(acc[dp * 2], a, b[0], b[1]);
        mma_bf16_16816(acc[dp * 2 + 1], a, b[2], b[3]);
    }
}


// common prologue: bias table (pre-multiplied by log2e) and key index table
CTC_DEVINL void load_bias(const AttnParams& p, int head, float* bias, int* tab, int count) {
    if (p.bias_table) {
        const int nb = (2 * p.H - 1) * (2 * p.W - 1);
        for (int i = threadIdx.x; i < nb; i += blockDim.x) bias[i] = p.bias_table[(long long)head * nb + i] * LOG2E;
        const int nW = 2 * p.W - 1;
        for (int j = threadIdx.x; j < count; j += blockDim.x) {
            const int jj = min(j, p.n - 1);   // padded keys are masked later; keep the index in range
            tab[j] = (jj / p.W) * nW + (jj % p.W);
        }
    }
}
CTC_DEVINL int bias_base(const AttnParams& p, int i) {
    const int ii = min(i, p.n - 1);
    return (ii / p.W + p.H - 1) * (2 * p.W - 1) + (ii % p.W + p.W - 1);
}
// Fast bias path (W % 8 == 0, so the 8 keys / queries of an MMA n-tile never straddle a grid row and the two
// columns a thread owns are table neighbours): the table is held as fp32 PAIRS pair[k] = (bias[k], bias[k-1])
// (pre-multiplied by log2e), so ONE 64-bit shared load yields both columns of a row, and the per-column index
// tables shrink to one entry per 8-column block, fetched with two broadcast 128-bit loads per 64-column step.
// (A bf16x2 pair table would halve the shared-memory wavefronts again, but it rounds the bias to 2^-9 and
// moved the noise-dominated random-init logit by 1.8e-3 in the full-size test; exact fp32 is kept.)
CTC_DEVINL void load_bias_pairs(const AttnParams& p, int head, float2* pair, int* blk, int count, bool rows_are_keys) {
    const int nW = 2 * p.W - 1;
    const int nb = (2 * p.H - 1) * nW;
    const float* tb = p.bias_table + (long long)head * nb;
    for (int k = threadIdx.x; k < nb; k += blockDim.x)
        pair[k] = make_float2(tb[k] * LOG2E, k > 0 ? tb[k - 1] * LOG2E : 0.f);
    for (int jb = threadIdx.x; jb < count / 8; jb += blockDim.x) {
        const int j = min(jb * 8, p.n - 8);               // padded blocks are masked later; keep the index in range
        // columns are keys (fwd, dQ): tab_j; columns are queries (dK/dV): base_i
        blk[jb] = rows_are_keys ? (j / p.W + p.H - 1) * nW + (j % p.W + p.W - 1) : (j / p.W) * nW + (j % p.W);
    }
}

// S tile (16 rows x KBLK keys) = A(16 x 32) * rows-of-B^T (+ bias) in the log2 domain
template <int KBLK, bool FB>
CTC_DEVINL void score_tile(float (&sc)[KBLK / 8][4], const uint32_t (&a)[2][4], uint32_t b_addr, int k0, int lane,
                           const float* bias, const int* tabj, int base0, int base1, bool has_bias) {
    const int t = lane & 3;
    if constexpr (FB) {
        // the bias is the accumulator's initial value; key-block indices come as two broadcast int4 loads
        const float2* pair = reinterpret_cast<const float2*>(bias);
        int tj8[KBLK / 8];
#pragma unroll
        for (int v = 0; v < KBLK / 32; ++v) {
            const int4 q4 = *reinterpret_cast<const int4*>(tabj + k0 / 8 + 4 * v);
            tj8[4 * v] = q4.x; tj8[4 * v + 1] = q4.y; tj8[4 * v + 2] = q4.z; tj8[4 * v + 3] = q4.w;
        }
#pragma unroll
        for (int nt = 0; nt < KBLK / 8; ++nt) {
            const int tj = tj8[nt] + 2 * t;
            const float2 f0 = pair[base0 - tj], f1 = pair[base1 - tj];
            sc[nt][0] = f0.x; sc[nt][1] = f0.y; sc[nt][2] = f1.x; sc[nt][3] = f1.y;
            mma_rowsB(sc[nt], a, b_addr, k0 + nt * 8, lane);
        }
        return;
    }
#pragma unroll
    for (int nt = 0; nt < KBLK / 8; ++nt) {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
        mma_rowsB(sc[nt], a, b_addr, k0 + nt * 8, lane);
        if (has_bias) {
            const int j = k0 + nt * 8 + 2 * t;
            const int tj0 = tabj[j], tj1 = tabj[j + 1];
            sc[nt][0] += bias[base0 - tj0]; sc[nt][1] += bias[base0 - tj1];
            sc[nt][2] += bias[base1 - tj0]; sc[nt][3] += bias[base1 - tj1];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward (PROBS=false) / probability materialisation (PROBS=true)
// ---------------------------------------------------------------------------------------------
template <int QB, int KBLK, int HPC, bool PROBS, bool FB>
__global__ void __launch_bounds__(HPC * (QB / 16) * 32, attn_min_blocks(HPC * (QB / 16) * 32, 384))
attn_fwd_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int WPH = QB / 16;
    const int head0 = blockIdx.y * HPC, s = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hl = warp / WPH, wq = warp % WPH, head = head0 + hl;
    const int kv_bytes = p.n_pad * 64;
    uint8_t* ks = sm;                                   // [HPC][n_pad][64 B]
    uint8_t* vs = ks + HPC * kv_bytes;
    uint8_t* qs = vs + HPC * kv_bytes;                  // [HPC][QB][64 B]
    float* sv = reinterpret_cast<float*>(qs + HPC * QB * 64);   // q_scale[32], k_scale[32]
    int* tabj = reinterpret_cast<int*>(sv + 64);
    float* bias = reinterpret_cast<float*>(tabj + p.n_pad);
    if (threadIdx.x < 32) sv[threadIdx.x] = p.q_scale[threadIdx.x];
    else if (threadIdx.x < 64) sv[threadIdx.x] = p.k_scale[threadIdx.x - 32];
    if (FB) load_bias_pairs(p, head0, reinterpret_cast<float2*>(bias), tabj, p.n_pad, false);
    else load_bias(p, head0, bias, tabj, p.n_pad);
    __syncthreads();
    load_tile<true>(ks, kv_bytes, p.k, p.ldkv, p, s, head0, HPC, 0, p.n_pad, sv + 32, 1.0f);
    if (!PROBS) load_tile<false>(vs, kv_bytes, p.v, p.ldkv, p, s, head0, HPC, 0, p.n_pad, nullptr, 1.0f);
    const uint32_t ks_a = smem_u32(ks + hl * kv_bytes), vs_a = smem_u32(vs + hl * kv_bytes);
    const uint32_t qs_a = smem_u32(qs + hl * QB * 64);
    const int g = lane >> 2, t = lane & 3;
    // K/V of the (sequence, head) stay resident; the CTA walks the query blocks gridDim.z apart (gridDim.z = 1:
    // one K/V load + normalisation per (sequence, head) instead of one per query block)
  for (int qb = blockIdx.z; qb * QB < p.n; qb += gridDim.z) {
    __syncthreads();                                    // previous block's Q fragments are in registers
    load_tile<true>(qs, QB * 64, p.q, p.ldq, p, s, head0, HPC, qb * QB, QB, sv, p.scale * LOG2E);
    __syncthreads();

    const int row_base = qb * QB + wq * 16;
    if (row_base >= p.n) continue;
    uint32_t aq[2][4];
    load_a_frags(aq, qs_a, wq * 16, lane);
    const int i0 = row_base + g, i1 = i0 + 8;
    const bool has_bias = p.bias_table != nullptr;
    const int base0 = has_bias ? bias_base(p, i0) : 0, base1 = has_bias ? bias_base(p, i1) : 0;
    const bool need_mask = (p.n % KBLK) != 0;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float oacc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) oacc[a][b] = 0.f;
    float lse2_0 = 0.f, lse2_1 = 0.f;
    if (PROBS) {
        lse2_0 = (i0 < p.n) ? p.lse[seq_row(p, s, i0) * p.heads + head] * LOG2E : 0.f;
        lse2_1 = (i1 < p.n) ? p.lse[seq_row(p, s, i1) * p.heads + head] * LOG2E : 0.f;
    }

    for (int kb = 0; kb < p.n_pad / KBLK; ++kb) {
        float sc[KBLK / 8][4];
        score_tile<KBLK, FB>(sc, aq, ks_a, kb * KBLK, lane, bias, tabj, base0, base1, has_bias);
        if (need_mask) {
#pragma unroll
            for (int nt = 0; nt < KBLK / 8; ++nt) {
                const int j = kb * KBLK + nt * 8 + 2 * t;
                if (j >= p.n) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
                if (j + 1 >= p.n) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
            }
        }
        if (PROBS) {
#pragma unroll
            for (int nt = 0; nt < KBLK / 8; ++nt) {
                const int j = kb * KBLK + nt * 8 + 2 * t;
                if (i0 < p.n) {
                    float* pr = p.probs + (((long long)s * p.heads + head) * p.n + i0) * p.n + j;
                    if (j < p.n) pr[0] = fast_exp2(sc[nt][0] - lse2_0);
                    if (j + 1 < p.n) pr[1] = fast_exp2(sc[nt][1] - lse2_0);
                }
                if (i1 < p.n) {
                    float* pr = p.probs + (((long long)s * p.heads + head) * p.n + i1) * p.n + j;
                    if (j < p.n) pr[0] = fast_exp2(sc[nt][2] - lse2_1);
                    if (j + 1 < p.n) pr[1] = fast_exp2(sc[nt][3] - lse2_1);
                }
            }
            continue;
        }
        // online softmax (log2 domain)
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < KBLK / 8; ++nt) {
            bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
            bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
        const float c0 = fast_exp2(m0 - nm0), c1 = fast_exp2(m1 - nm1);
        m0 = nm0; m1 = nm1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < KBLK / 8; ++nt) {
            sc[nt][0] = fast_exp2(sc[nt][0] - m0); sc[nt][1] = fast_exp2(sc[nt][1] - m0);
            sc[nt][2] = fast_exp2(sc[nt][2] - m1); sc[nt][3] = fast_exp2(sc[nt][3] - m1);
            rs0 += sc[nt][0] + sc[nt][1]; rs1 += sc[nt][2] + sc[nt][3];
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
        for (int a = 0; a < 4; ++a) { oacc[a][0] *= c0; oacc[a][1] *= c0; oacc[a][2] *= c1; oacc[a][3] *= c1; }
#pragma unroll
        for (int kk = 0; kk < KBLK / 16; ++kk) {
            uint32_t a[4];
            a[0] = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]);
            a[1] = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
            a[2] = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
            a[3] = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
            mma_colsB(oacc, a, vs_a, kb * KBLK + kk * 16, lane);
        }
    }
    if (PROBS) continue;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    if (i0 < p.n) {
        const long long r = seq_row(p, s, i0);
        __nv_bfloat16* orow = p.out + r * (p.heads * DH) + head * DH;
#pragma unroll
        for (int a = 0; a < 4; ++a)
            *reinterpret_cast<uint32_t*>(orow + a * 8 + 2 * t) = pack_bf16(oacc[a][0] * inv0, oacc[a][1] * inv0);
        if (t == 0) p.lse[r * p.heads + head] = (m0 + log2f(l0)) * LN2;
    }
    if (i1 < p.n) {
        const long long r = seq_row(p, s, i1);
        __nv_bfloat16* orow = p.out + r * (p.heads * DH) + head * DH;
#pragma unroll
        for (int a = 0; a < 4; ++a)
            *reinterpret_cast<uint32_t*>(orow + a * 8 + 2 * t) = pack_bf16(oacc[a][2] * inv1, oacc[a][3] * inv1);
        if (t == 0) p.lse[r * p.heads + head] = (m1 + log2f(l1)) * LN2;
    }
  }
}

// adjoint of x^ = l2norm(x) * vec for one row held in mma C layout (quad of lanes owns the row):
// g = gradient w.r.t. x^ (before the vec factor is applied here).  Returns dx for the 8 elements this
// thread owns (cols a*8 + 2t, +1 for a = 0..3).
// Contains full-mask shuffles: EVERY lane of the warp must call it (pass xrow = nullptr for rows outside the
// sequence; their result is garbage and must not be stored).
CTC_DEVINL void l2norm_adjoint_row(const __nv_bfloat16* xrow, const float* vec, int t, float (&g)[8], float (&dx)[8]) {
    float x[8];
    float ss = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float2 v = xrow ? unpack_bf16(*reinterpret_cast<const uint32_t*>(xrow + a * 8 + 2 * t)) : make_float2(0.f, 0.f);
        x[a * 2] = v.x; x[a * 2 + 1] = v.y;
        ss += v.x * v.x + v.y * v.y;
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    float dot = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        g[a * 2] *= vec[a * 8 + 2 * t]; g[a * 2 + 1] *= vec[a * 8 + 2 * t + 1];
        x[a * 2] *= inv; x[a * 2 + 1] *= inv;
        dot += x[a * 2] * g[a * 2] + x[a * 2 + 1] * g[a * 2 + 1];
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1); dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
    for (int e = 0; e < 8; ++e) dx[e] = (g[e] - x[e] * dot) * inv;
}

// ---------------------------------------------------------------------------------------------
// backward, dQ: one CTA per (QB-query block, head group, sequence)
// ---------------------------------------------------------------------------------------------
template <int QB, int KBLK, int HPC, bool FB>
__global__ void __launch_bounds__(HPC * (QB / 16) * 32, attn_min_blocks(HPC * (QB / 16) * 32, 256))
attn_bwd_dq_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int WPH = QB / 16;
    const int head0 = blockIdx.y * HPC, s = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hl = warp / WPH, wq = warp % WPH, head = head0 + hl;
    const int kv_bytes = p.n_pad * 64;
    uint8_t* ks = sm;
    uint8_t* vs = ks + HPC * kv_bytes;
    uint8_t* qs = vs + HPC * kv_bytes;
    uint8_t* dos = qs + HPC * QB * 64;
    float* sv = reinterpret_cast<float*>(dos + HPC * QB * 64);
    float* dl = sv + 64;                                   // D for the HPC*QB rows
    int* tabj = reinterpret_cast<int*>(dl + HPC * QB);
    float* bias = reinterpret_cast<float*>(tabj + p.n_pad);
    if (threadIdx.x < 32) sv[threadIdx.x] = p.q_scale[threadIdx.x];
    else if (threadIdx.x < 64) sv[threadIdx.x] = p.k_scale[threadIdx.x - 32];
    if (FB) load_bias_pairs(p, head0, reinterpret_cast<float2*>(bias), tabj, p.n_pad, false);
    else load_bias(p, head0, bias, tabj, p.n_pad);
    __syncthreads();
    load_tile<true>(ks, kv_bytes, p.k, p.ldkv, p, s, head0, HPC, 0, p.n_pad, sv + 32, 1.0f);
    load_tile<false>(vs, kv_bytes, p.v, p.ldkv, p, s, head0, HPC, 0, p.n_pad, nullptr, 1.0f);
    const uint32_t ks_a = smem_u32(ks + hl * kv_bytes), vs_a = smem_u32(vs + hl * kv_bytes);
    const uint32_t qs_a = smem_u32(qs + hl * QB * 64), dos_a = smem_u32(dos + hl * QB * 64);
    const int g = lane >> 2, t = lane & 3;
    const bool has_bias = p.bias_table != nullptr;
    const bool need_mask = (p.n % KBLK) != 0;
  // K/V stay resident; the CTA walks the query blocks gridDim.z apart
  for (int qb = blockIdx.z; qb * QB < p.n; qb += gridDim.z) {
    __syncthreads();
    load_tile<true>(qs, QB * 64, p.q, p.ldq, p, s, head0, HPC, qb * QB, QB, sv, p.scale * LOG2E);
    load_tile<false>(dos, QB * 64, p.d_o, (long long)p.heads * DH, p, s, head0, HPC, qb * QB, QB, nullptr, 1.0f);
    // D_i = sum_d dO_i,d * O_i,d   (thread per (row, head))
    for (int idx = threadIdx.x; idx < QB * HPC; idx += blockDim.x) {
        const int r_ = idx / HPC, h_ = idx - r_ * HPC;
        const int i = qb * QB + r_;
        float d = 0.f;
        if (i < p.n) {
            const long long r = seq_row(p, s, i);
            const uint4* go = reinterpret_cast<const uint4*>(p.o + r * (p.heads * DH) + (head0 + h_) * DH);
            const uint4* gd = reinterpret_cast<const uint4*>(p.d_o + r * (p.heads * DH) + (head0 + h_) * DH);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 a = go[j], b = gd[j];
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = unpack_bf16(aw[e]), y = unpack_bf16(bw[e]);
                    d += x.x * y.x + x.y * y.y;
                }
            }
            p.delta[r * p.heads + head0 + h_] = d;
        }
        dl[h_ * QB + r_] = d;
    }
    __syncthreads();

    const int row_base = qb * QB + wq * 16;
    if (row_base >= p.n) continue;
    uint32_t aq[2][4], ado[2][4];
    load_a_frags(aq, qs_a, wq * 16, lane);
    load_a_frags(ado, dos_a, wq * 16, lane);
    const int i0 = row_base + g, i1 = i0 + 8;
    const int base0 = has_bias ? bias_base(p, i0) : 0, base1 = has_bias ? bias_base(p, i1) : 0;
    const float lse2_0 = (i0 < p.n) ? p.lse[seq_row(p, s, i0) * p.heads + head] * LOG2E : INFINITY;
    const float lse2_1 = (i1 < p.n) ? p.lse[seq_row(p, s, i1) * p.heads + head] * LOG2E : INFINITY;
    const float d0 = dl[hl * QB + wq * 16 + g], d1 = dl[hl * QB + wq * 16 + g + 8];
    float dq[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dq[a][b] = 0.f;

    for (int kb = 0; kb < p.n_pad / KBLK; ++kb) {
        float sc[KBLK / 8][4];
        score_tile<KBLK, FB>(sc, aq, ks_a, kb * KBLK, lane, bias, tabj, base0, base1, has_bias);
#pragma unroll
        for (int nt = 0; nt < KBLK / 8; ++nt) {
            float dp[4] = {0.f, 0.f, 0.f, 0.f};
            mma_rowsB(dp, ado, vs_a, kb * KBLK + nt * 8, lane);
            float p0 = fast_exp2(sc[nt][0] - lse2_0), p1 = fast_exp2(sc[nt][1] - lse2_0);
            float p2 = fast_exp2(sc[nt][2] - lse2_1), p3 = fast_exp2(sc[nt][3] - lse2_1);
            if (need_mask) {
                const int j = kb * KBLK + nt * 8 + 2 * t;
                if (j >= p.n) { p0 = 0.f; p2 = 0.f; }
                if (j + 1 >= p.n) { p1 = 0.f; p3 = 0.f; }
            }
            sc[nt][0] = p0 * (dp[0] - d0); sc[nt][1] = p1 * (dp[1] - d0);
            sc[nt][2] = p2 * (dp[2] - d1); sc[nt][3] = p3 * (dp[3] - d1);
        }
#pragma unroll
        for (int kk = 0; kk < KBLK / 16; ++kk) {
            uint32_t a[4];
            a[0] = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]);
            a[1] = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
            a[2] = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
            a[3] = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
            mma_colsB(dq, a, ks_a, kb * KBLK + kk * 16, lane);
        }
    }
    // dq^ -> dq through l2norm and q_scale; ds/dq^_d = scale * q_scale_d * k^_d  (k^ includes k_scale)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int i = half ? i1 : i0;
        const bool ok = i < p.n;                       // no early-out: the adjoint shuffles need the whole warp
        const long long r = seq_row(p, s, ok ? i : 0);
        float gq[8], dx[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) { gq[a * 2] = dq[a][half * 2] * p.scale; gq[a * 2 + 1] = dq[a][half * 2 + 1] * p.scale; }
        l2norm_adjoint_row(ok ? p.q + r * p.ldq + head * DH : nullptr, sv, t, gq, dx);
        __nv_bfloat16* drow = p.dq + r * p.lddq + head * DH;
        if (ok) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
                *reinterpret_cast<uint32_t*>(drow + a * 8 + 2 * t) = pack_bf16(dx[a * 2], dx[a * 2 + 1]);
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, dK/dV: one CTA per (QB-key block, head group, sequence); queries are the reduction axis,
// walked in blocks of KBLK.
// ---------------------------------------------------------------------------------------------
template <int QB, int KBLK, int HPC, bool FB>
__global__ void __launch_bounds__(HPC * (QB / 16) * 32, attn_min_blocks(HPC * (QB / 16) * 32, 192))
attn_bwd_dkv_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int WPH = QB / 16;
    const int head0 = blockIdx.y * HPC, s = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hl = warp / WPH, wq = warp % WPH, head = head0 + hl;
    const int kv_bytes = p.n_pad * 64;
    uint8_t* qs = sm;                                       // [HPC] all queries (scaled q^)
    uint8_t* dos = qs + HPC * kv_bytes;                     // [HPC] all dO rows
    uint8_t* ks = dos + HPC * kv_bytes;                     // [HPC] my QB keys (k^)
    uint8_t* vs = ks + HPC * QB * 64;                       // [HPC] my QB values
    float* sv = reinterpret_cast<float*>(vs + HPC * QB * 64);
    float* lse2 = sv + 64;                                  // [HPC][n_pad]
    float* dl = lse2 + HPC * p.n_pad;                       // [HPC][n_pad]
    int* basei = reinterpret_cast<int*>(dl + HPC * p.n_pad);   // [n_pad]
    float* bias = reinterpret_cast<float*>(basei + p.n_pad);
    if (threadIdx.x < 32) sv[threadIdx.x] = p.q_scale[threadIdx.x];
    else if (threadIdx.x < 64) sv[threadIdx.x] = p.k_scale[threadIdx.x - 32];
    const int nW = 2 * p.W - 1;
    const bool has_bias = p.bias_table != nullptr;
    if (FB) {
        load_bias_pairs(p, head0, reinterpret_cast<float2*>(bias), basei, p.n_pad, true);
    } else if (has_bias) {
        const int nb = (2 * p.H - 1) * nW;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) bias[i] = p.bias_table[(long long)head0 * nb + i] * LOG2E;
        for (int i = threadIdx.x; i < p.n_pad; i += blockDim.x) basei[i] = bias_base(p, i);
    }
    for (int idx = threadIdx.x; idx < p.n_pad * HPC; idx += blockDim.x) {
        const int i = idx / HPC, h_ = idx - i * HPC;
        if (i < p.n) {
            const long long r = seq_row(p, s, i);
            lse2[h_ * p.n_pad + i] = p.lse[r * p.heads + head0 + h_] * LOG2E;
            dl[h_ * p.n_pad + i] = p.delta[r * p.heads + head0 + h_];
        } else { lse2[h_ * p.n_pad + i] = INFINITY; dl[h_ * p.n_pad + i] = 0.f; }
    }
    __syncthreads();
    load_tile<true>(qs, kv_bytes, p.q, p.ldq, p, s, head0, HPC, 0, p.n_pad, sv, p.scale * LOG2E);
    load_tile<false>(dos, kv_bytes, p.d_o, (long long)p.heads * DH, p, s, head0, HPC, 0, p.n_pad, nullptr, 1.0f);
    const uint32_t ks_a = smem_u32(ks + hl * QB * 64), vs_a = smem_u32(vs + hl * QB * 64);
    const uint32_t qs_a = smem_u32(qs + hl * kv_bytes), dos_a = smem_u32(dos + hl * kv_bytes);
    const float* lse2h = lse2 + hl * p.n_pad;
    const float* dlh = dl + hl * p.n_pad;
    const int g = lane >> 2, t = lane & 3;
  // Q / dO / lse / delta of the (sequence, head) stay resident; the CTA walks the key blocks gridDim.z apart
  for (int kblk = blockIdx.z; kblk * QB < p.n; kblk += gridDim.z) {
    __syncthreads();
    load_tile<true>(ks, QB * 64, p.k, p.ldkv, p, s, head0, HPC, kblk * QB, QB, sv + 32, 1.0f);
    load_tile<false>(vs, QB * 64, p.v, p.ldkv, p, s, head0, HPC, kblk * QB, QB, nullptr, 1.0f);
    __syncthreads();

    const int row_base = kblk * QB + wq * 16;
    if (row_base >= p.n) continue;
    uint32_t ak[2][4], av[2][4];
    load_a_frags(ak, ks_a, wq * 16, lane);
    load_a_frags(av, vs_a, wq * 16, lane);
    const int j0 = row_base + g, j1 = j0 + 8;
    int tab0 = 0, tab1 = 0;
    if (has_bias) {
        tab0 = (min(j0, p.n - 1) / p.W) * nW + min(j0, p.n - 1) % p.W;
        tab1 = (min(j1, p.n - 1) / p.W) * nW + min(j1, p.n - 1) % p.W;
    }
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { dk[a][b] = 0.f; dv[a][b] = 0.f; }

    for (int qblk = 0; qblk < p.n_pad / KBLK; ++qblk) {
        float pt[KBLK / 8][4], ds[KBLK / 8][4];
        int bi8[KBLK / 8];
        if constexpr (FB) {
#pragma unroll
            for (int v = 0; v < KBLK / 32; ++v) {
                const int4 q4 = *reinterpret_cast<const int4*>(basei + qblk * (KBLK / 8) + 4 * v);
                bi8[4 * v] = q4.x; bi8[4 * v + 1] = q4.y; bi8[4 * v + 2] = q4.z; bi8[4 * v + 3] = q4.w;
            }
        }
#pragma unroll
        for (int nt = 0; nt < KBLK / 8; ++nt) {
            // S^T tile: rows = my keys, cols = queries
            float st[4] = {0.f, 0.f, 0.f, 0.f};
            const int i = qblk * KBLK + nt * 8 + 2 * t;
            if constexpr (FB) {
                // queries i, i+1 are table neighbours: pair[b + 1] = (lo bias[b + 1] -> query i+1, hi bias[b] -> query i)
                const float2* pair = reinterpret_cast<const float2*>(bias);
                const int b = bi8[nt] + 2 * t + 1;
                const float2 f0 = pair[b - tab0], f1 = pair[b - tab1];
                st[0] = f0.y; st[1] = f0.x; st[2] = f1.y; st[3] = f1.x;
            }
            mma_rowsB(st, ak, qs_a, qblk * KBLK + nt * 8, lane);
            if (!FB && has_bias) {
                const int b0 = basei[i], b1 = basei[i + 1];
                st[0] += bias[b0 - tab0]; st[1] += bias[b1 - tab0];
                st[2] += bias[b0 - tab1]; st[3] += bias[b1 - tab1];
            }
            const float l0 = lse2h[i], l1 = lse2h[i + 1];      // +inf for padded queries -> P = 0
            pt[nt][0] = fast_exp2(st[0] - l0); pt[nt][1] = fast_exp2(st[1] - l1);
            pt[nt][2] = fast_exp2(st[2] - l0); pt[nt][3] = fast_exp2(st[3] - l1);
            float dp[4] = {0.f, 0.f, 0.f, 0.f};
            mma_rowsB(dp, av, dos_a, qblk * KBLK + nt * 8, lane);
            const float dd0 = dlh[i], dd1 = dlh[i + 1];
            ds[nt][0] = pt[nt][0] * (dp[0] - dd0); ds[nt][1] = pt[nt][1] * (dp[1] - dd1);
            ds[nt][2] = pt[nt][2] * (dp[2] - dd0); ds[nt][3] = pt[nt][3] * (dp[3] - dd1);
        }
#pragma unroll
        for (int kk = 0; kk < KBLK / 16; ++kk) {
            uint32_t a[4];
            a[0] = pack_bf16(pt[2 * kk][0], pt[2 * kk][1]);
            a[1] = pack_bf16(pt[2 * kk][2], pt[2 * kk][3]);
            a[2] = pack_bf16(pt[2 * kk + 1][0], pt[2 * kk + 1][1]);
            a[3] = pack_bf16(pt[2 * kk + 1][2], pt[2 * kk + 1][3]);
            mma_colsB(dv, a, dos_a, qblk * KBLK + kk * 16, lane);
            a[0] = pack_bf16(ds[2 * kk][0], ds[2 * kk][1]);
            a[1] = pack_bf16(ds[2 * kk][2], ds[2 * kk][3]);
            a[2] = pack_bf16(ds[2 * kk + 1][0], ds[2 * kk + 1][1]);
            a[3] = pack_bf16(ds[2 * kk + 1][2], ds[2 * kk + 1][3]);
            mma_colsB(dk, a, qs_a, qblk * KBLK + kk * 16, lane);
        }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = half ? j1 : j0;
        const bool ok = j < p.n;                       // no early-out: the adjoint shuffles need the whole warp
        const long long r = seq_row(p, s, ok ? j : 0);
        __nv_bfloat16* dvrow = p.dv + r * p.lddkv + head * DH;
        if (ok) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
                *reinterpret_cast<uint32_t*>(dvrow + a * 8 + 2 * t) = pack_bf16(dv[a][half * 2], dv[a][half * 2 + 1]);
        }
        // dk^ accumulated against q^ * scale * log2e  ->  undo log2e; k_scale applied inside the adjoint
        float gk[8], dx[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) { gk[a * 2] = dk[a][half * 2] * LN2; gk[a * 2 + 1] = dk[a][half * 2 + 1] * LN2; }
        l2norm_adjoint_row(ok ? p.k + r * p.ldkv + head * DH : nullptr, sv + 32, t, gk, dx);
        __nv_bfloat16* dkrow = p.dk + r * p.lddkv + head * DH;
        if (ok) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
                *reinterpret_cast<uint32_t*>(dkrow + a * 8 + 2 * t) = pack_bf16(dx[a * 2], dx[a * 2 + 1]);
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Small sequences (temporal attention: n = T <= 32 tokens, no bias): warp-autonomous kernels.
//
// A (sequence, head) problem is only 24 x 24 scores; with one CTA per sequence the work per CTA is a few
// hundred nanoseconds of tensor-core time behind a global-load latency and a barrier (ncu: 5.8-7.2 warps
// stalled on the CTA barrier per issued instruction, 12 % of DRAM bandwidth).  Here every WARP owns whole
// (sequence, head) problems and walks a strided task list with a private two-stage shared-memory ring:
// cp.async (16-byte LDGSTS) prefetches the next task's q/k/v(/dO) rows while the current task is normalised in
// place and multiplied, and the only synchronisation is __syncwarp().  No CTA barrier after the prologue.
// ---------------------------------------------------------------------------------------------
static constexpr int SMALL_N = 32;                 // padded rows per tile
static constexpr int SMALL_TILE = SMALL_N * 64;    // bytes per [32 rows x 32 bf16] tile
static constexpr int SMALL_FWD_WARPS = 8, SMALL_BWD_WARPS = 12;

CTC_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
CTC_DEVINL void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows [0, n) of one head's [n x 32] bf16 slice -> swizzled tile, 16 bytes per cp.async, 4 consecutive lanes per row
CTC_DEVINL void small_issue_tile(uint8_t* tile, const __nv_bfloat16* src, long long ld, const AttnParams& p, int s,
                                 int head, int lane) {
    for (int idx = lane; idx < p.n * 4; idx += 32) {
        const int r = idx >> 2, c = idx & 3;
        cp_async_16(tile + tile_off(r, c), src + seq_row(p, s, r) * ld + head * DH + c * 8);
    }
}
// in-place l2norm * vec * mul of the rows [0, n) of a tile (lane = row)
CTC_DEVINL void small_normalise(uint8_t* tile, int n, const float* vec, float mul, int lane) {
    if (lane >= n) return;
    uint4 c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = *reinterpret_cast<const uint4*>(tile + tile_off(lane, j));
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
    const float inv = mul / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        c[j].x = pack_bf16(f[j * 8 + 0] * inv * vec[j * 8 + 0], f[j * 8 + 1] * inv * vec[j * 8 + 1]);
        c[j].y = pack_bf16(f[j * 8 + 2] * inv * vec[j * 8 + 2], f[j * 8 + 3] * inv * vec[j * 8 + 3]);
        c[j].z = pack_bf16(f[j * 8 + 4] * inv * vec[j * 8 + 4], f[j * 8 + 5] * inv * vec[j * 8 + 5]);
        c[j].w = pack_bf16(f[j * 8 + 6] * inv * vec[j * 8 + 6], f[j * 8 + 7] * inv * vec[j * 8 + 7]);
        *reinterpret_cast<uint4*>(tile + tile_off(lane, j)) = c[j];
    }
}

__global__ void __launch_bounds__(SMALL_FWD_WARPS * 32, 2)
attn_small_fwd_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sv = reinterpret_cast<float*>(sm);                       // q_scale[32], k_scale[32]
    uint8_t* mine = sm + 256 + warp * (2 * 3 * SMALL_TILE);          // [stage][q, k, v]
    if (threadIdx.x < 32) sv[threadIdx.x] = p.q_scale[threadIdx.x];
    else if (threadIdx.x < 64) sv[threadIdx.x] = p.k_scale[threadIdx.x - 32];
    for (int i = lane; i < 2 * 3 * SMALL_TILE / 16; i += 32) reinterpret_cast<uint4*>(mine)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();                                                 // the only CTA barrier
    const int n_tasks = p.n_seq * p.heads;
    const int stride = gridDim.x * SMALL_FWD_WARPS;
    int task = blockIdx.x * SMALL_FWD_WARPS + warp;
    auto issue = [&](int tk, int st) {
        if (tk < n_tasks) {
            const int s = tk / p.heads, head = tk - s * p.heads;
            uint8_t* b = mine + st * 3 * SMALL_TILE;
            small_issue_tile(b, p.q, p.ldq, p, s, head, lane);
            small_issue_tile(b + SMALL_TILE, p.k, p.ldkv, p, s, head, lane);
            small_issue_tile(b + 2 * SMALL_TILE, p.v, p.ldkv, p, s, head, lane);
        }
        cp_async_commit();
    };
    issue(task, 0);
    const int g = lane >> 2, t = lane & 3;
    for (int it = 0; task < n_tasks; task += stride, ++it) {
        const int st = it & 1;
        issue(task + stride, st ^ 1);
        cp_async_wait_group<1>();
        __syncwarp();
        uint8_t* qs = mine + st * 3 * SMALL_TILE;
        uint8_t* ks = qs + SMALL_TILE;
        uint8_t* vs = ks + SMALL_TILE;
        small_normalise(qs, p.n, sv, p.scale * LOG2E, lane);
        small_normalise(ks, p.n, sv + 32, 1.0f, lane);
        __syncwarp();
        const int s = task / p.heads, head = task - s * p.heads;
        const uint32_t qs_a = smem_u32(qs), ks_a = smem_u32(ks), vs_a = smem_u32(vs);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            if (mt * 16 >= p.n) break;
            uint32_t aq[2][4];
            load_a_frags(aq, qs_a, mt * 16, lane);
            float sc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
                mma_rowsB(sc[nt], aq, ks_a, nt * 8, lane);
                const int j = nt * 8 + 2 * t;
                if (j >= p.n) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
                if (j + 1 >= p.n) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
            }
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
                m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                sc[nt][0] = fast_exp2(sc[nt][0] - m0); sc[nt][1] = fast_exp2(sc[nt][1] - m0);
                sc[nt][2] = fast_exp2(sc[nt][2] - m1); sc[nt][3] = fast_exp2(sc[nt][3] - m1);
                l0 += sc[nt][0] + sc[nt][1]; l1 += sc[nt][2] + sc[nt][3];
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            float oacc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a) oacc[a][0] = oacc[a][1] = oacc[a][2] = oacc[a][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                uint32_t a[4];
                a[0] = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]);
                a[1] = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
                a[2] = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
                a[3] = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
                mma_colsB(oacc, a, vs_a, kk * 16, lane);
            }
            const float inv0 = 1.f / l0, inv1 = 1.f / l1;
            const int i0 = mt * 16 + g, i1 = i0 + 8;
            if (i0 < p.n) {
                const long long r = seq_row(p, s, i0);
                __nv_bfloat16* orow = p.out + r * (p.heads * DH) + head * DH;
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    *reinterpret_cast<uint32_t*>(orow + a * 8 + 2 * t) = pack_bf16(oacc[a][0] * inv0, oacc[a][1] * inv0);
                if (t == 0) p.lse[r * p.heads + head] = (m0 + log2f(l0)) * LN2;
            }
            if (i1 < p.n) {
                const long long r = seq_row(p, s, i1);
                __nv_bfloat16* orow = p.out + r * (p.heads * DH) + head * DH;
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    *reinterpret_cast<uint32_t*>(orow + a * 8 + 2 * t) = pack_bf16(oacc[a][2] * inv1, oacc[a][3] * inv1);
                if (t == 0) p.lse[r * p.heads + head] = (m1 + log2f(l1)) * LN2;
            }
        }
        __syncwarp();                                                // stage st is re-filled by the next issue()
    }
    cp_async_wait_group<0>();
}

// Backward for small sequences: one warp computes dQ, dK and dV of a (sequence, head) from resident q^, k^, v, dO
// tiles (dQ from S / dP tiles with query rows, dK / dV from the transposed tiles with key rows - recomputing the
// 24 x 24 scores twice is cheaper than transposing fragments).
__global__ void __launch_bounds__(SMALL_BWD_WARPS * 32, 1)
attn_small_bwd_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sv = reinterpret_cast<float*>(sm);
    uint8_t* mine = sm + 256 + warp * (2 * 4 * SMALL_TILE);          // [stage][q, k, v, dO]
    if (threadIdx.x < 32) sv[threadIdx.x] = p.q_scale[threadIdx.x];
    else if (threadIdx.x < 64) sv[threadIdx.x] = p.k_scale[threadIdx.x - 32];
    for (int i = lane; i < 2 * 4 * SMALL_TILE / 16; i += 32) reinterpret_cast<uint4*>(mine)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int n_tasks = p.n_seq * p.heads;
    const int stride = gridDim.x * SMALL_BWD_WARPS;
    int task = blockIdx.x * SMALL_BWD_WARPS + warp;
    const long long ldo = (long long)p.heads * DH;
    auto issue = [&](int tk, int st) {
        if (tk < n_tasks) {
            const int s = tk / p.heads, head = tk - s * p.heads;
            uint8_t* b = mine + st * 4 * SMALL_TILE;
            small_issue_tile(b, p.q, p.ldq, p, s, head, lane);
            small_issue_tile(b + SMALL_TILE, p.k, p.ldkv, p, s, head, lane);
            small_issue_tile(b + 2 * SMALL_TILE, p.v, p.ldkv, p, s, head, lane);
            small_issue_tile(b + 3 * SMALL_TILE, p.d_o, ldo, p, s, head, lane);
        }
        cp_async_commit();
    };
    issue(task, 0);
    const int g = lane >> 2, t = lane & 3;
    for (int it = 0; task < n_tasks; task += stride, ++it) {
        const int st = it & 1;
        issue(task + stride, st ^ 1);
        const int s = task / p.heads, head = task - s * p.heads;
        // lane = row: the output row o (for D = rowsum(dO o O)) and the row log-sum-exp come straight from global
        uint4 orow[4];
        float lse2 = INFINITY;                                       // padded rows: P = exp2(S - inf) = 0
        if (lane < p.n) {
            const long long r = seq_row(p, s, lane);
            const uint4* go = reinterpret_cast<const uint4*>(p.o + r * ldo + head * DH);
#pragma unroll
            for (int j = 0; j < 4; ++j) orow[j] = go[j];
            lse2 = p.lse[r * p.heads + head] * LOG2E;
        }
        cp_async_wait_group<1>();
        __syncwarp();
        uint8_t* qs = mine + st * 4 * SMALL_TILE;
        uint8_t* ks = qs + SMALL_TILE;
        uint8_t* vs = ks + SMALL_TILE;
        uint8_t* dos = vs + SMALL_TILE;
        float dlt = 0.f;
        if (lane < p.n) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 a = orow[j], b = *reinterpret_cast<const uint4*>(dos + tile_off(lane, j));
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = unpack_bf16(aw[e]), y = unpack_bf16(bw[e]);
                    dlt += x.x * y.x + x.y * y.y;
                }
            }
            if (p.delta) p.delta[seq_row(p, s, lane) * p.heads + head] = dlt;
        }
        small_normalise(qs, p.n, sv, p.scale * LOG2E, lane);
        small_normalise(ks, p.n, sv + 32, 1.0f, lane);
        __syncwarp();
        const uint32_t qs_a = smem_u32(qs), ks_a = smem_u32(ks), vs_a = smem_u32(vs), dos_a = smem_u32(dos);
        // ---- dQ: rows = queries
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            if (mt * 16 >= p.n) break;
            uint32_t aq[2][4], ado[2][4];
            load_a_frags(aq, qs_a, mt * 16, lane);
            load_a_frags(ado, dos_a, mt * 16, lane);
            const int i0 = mt * 16 + g, i1 = i0 + 8;
            const float l0 = __shfl_sync(0xffffffffu, lse2, i0), l1 = __shfl_sync(0xffffffffu, lse2, i1);
            const float d0 = __shfl_sync(0xffffffffu, dlt, i0), d1 = __shfl_sync(0xffffffffu, dlt, i1);
            float ds[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                mma_rowsB(sc, aq, ks_a, nt * 8, lane);
                mma_rowsB(dp, ado, vs_a, nt * 8, lane);
                float p0 = fast_exp2(sc[0] - l0), p1 = fast_exp2(sc[1] - l0);
                float p2 = fast_exp2(sc[2] - l1), p3 = fast_exp2(sc[3] - l1);
                const int j = nt * 8 + 2 * t;
                if (j >= p.n) { p0 = 0.f; p2 = 0.f; }
                if (j + 1 >= p.n) { p1 = 0.f; p3 = 0.f; }
                ds[nt][0] = p0 * (dp[0] - d0); ds[nt][1] = p1 * (dp[1] - d0);
                ds[nt][2] = p2 * (dp[2] - d1); ds[nt][3] = p3 * (dp[3] - d1);
            }
            float dq[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a) dq[a][0] = dq[a][1] = dq[a][2] = dq[a][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                uint32_t a[4];
                a[0] = pack_bf16(ds[2 * kk][0], ds[2 * kk][1]);
                a[1] = pack_bf16(ds[2 * kk][2], ds[2 * kk][3]);
                a[2] = pack_bf16(ds[2 * kk + 1][0], ds[2 * kk + 1][1]);
                a[3] = pack_bf16(ds[2 * kk + 1][2], ds[2 * kk + 1][3]);
                mma_colsB(dq, a, ks_a, kk * 16, lane);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = half ? i1 : i0;
                const bool ok = i < p.n;               // no early-out: the adjoint shuffles need the whole warp
                const long long r = seq_row(p, s, ok ? i : 0);
                float gq[8], dx[8];
#pragma unroll
                for (int a = 0; a < 4; ++a) { gq[a * 2] = dq[a][half * 2] * p.scale; gq[a * 2 + 1] = dq[a][half * 2 + 1] * p.scale; }
                l2norm_adjoint_row(ok ? p.q + r * p.ldq + head * DH : nullptr, sv, t, gq, dx);
                __nv_bfloat16* drow = p.dq + r * p.lddq + head * DH;
                if (ok) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        *reinterpret_cast<uint32_t*>(drow + a * 8 + 2 * t) = pack_bf16(dx[a * 2], dx[a * 2 + 1]);
                }
            }
        }
        // ---- dK, dV: rows = keys, columns = queries
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            if (mt * 16 >= p.n) break;
            uint32_t ak[2][4], av[2][4];
            load_a_frags(ak, ks_a, mt * 16, lane);
            load_a_frags(av, vs_a, mt * 16, lane);
            float pt[4][4], dst[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float stt[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                mma_rowsB(stt, ak, qs_a, nt * 8, lane);
                mma_rowsB(dp, av, dos_a, nt * 8, lane);
                const int i = nt * 8 + 2 * t;
                const float l0 = __shfl_sync(0xffffffffu, lse2, i), l1 = __shfl_sync(0xffffffffu, lse2, i + 1);
                const float dd0 = __shfl_sync(0xffffffffu, dlt, i), dd1 = __shfl_sync(0xffffffffu, dlt, i + 1);
                pt[nt][0] = fast_exp2(stt[0] - l0); pt[nt][1] = fast_exp2(stt[1] - l1);
                pt[nt][2] = fast_exp2(stt[2] - l0); pt[nt][3] = fast_exp2(stt[3] - l1);
                dst[nt][0] = pt[nt][0] * (dp[0] - dd0); dst[nt][1] = pt[nt][1] * (dp[1] - dd1);
                dst[nt][2] = pt[nt][2] * (dp[2] - dd0); dst[nt][3] = pt[nt][3] * (dp[3] - dd1);
            }
            float dk[4][4], dv[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { dk[a][0] = dk[a][1] = dk[a][2] = dk[a][3] = 0.f; dv[a][0] = dv[a][1] = dv[a][2] = dv[a][3] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                uint32_t a[4];
                a[0] = pack_bf16(pt[2 * kk][0], pt[2 * kk][1]);
                a[1] = pack_bf16(pt[2 * kk][2], pt[2 * kk][3]);
                a[2] = pack_bf16(pt[2 * kk + 1][0], pt[2 * kk + 1][1]);
                a[3] = pack_bf16(pt[2 * kk + 1][2], pt[2 * kk + 1][3]);
                mma_colsB(dv, a, dos_a, kk * 16, lane);
                a[0] = pack_bf16(dst[2 * kk][0], dst[2 * kk][1]);
                a[1] = pack_bf16(dst[2 * kk][2], dst[2 * kk][3]);
                a[2] = pack_bf16(dst[2 * kk + 1][0], dst[2 * kk + 1][1]);
                a[3] = pack_bf16(dst[2 * kk + 1][2], dst[2 * kk + 1][3]);
                mma_colsB(dk, a, qs_a, kk * 16, lane);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int j = mt * 16 + g + 8 * half;
                const bool ok = j < p.n;               // no early-out: the adjoint shuffles need the whole warp
                const long long r = seq_row(p, s, ok ? j : 0);
                __nv_bfloat16* dvrow = p.dv + r * p.lddkv + head * DH;
                if (ok) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        *reinterpret_cast<uint32_t*>(dvrow + a * 8 + 2 * t) = pack_bf16(dv[a][half * 2], dv[a][half * 2 + 1]);
                }
                float gk[8], dx[8];
#pragma unroll
                for (int a = 0; a < 4; ++a) { gk[a * 2] = dk[a][half * 2] * LN2; gk[a * 2 + 1] = dk[a][half * 2 + 1] * LN2; }
                l2norm_adjoint_row(ok ? p.k + r * p.ldkv + head * DH : nullptr, sv + 32, t, gk, dx);
                __nv_bfloat16* dkrow = p.dk + r * p.lddkv + head * DH;
                if (ok) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        *reinterpret_cast<uint32_t*>(dkrow + a * 8 + 2 * t) = pack_bf16(dx[a * 2], dx[a * 2 + 1]);
                }
            }
        }
        __syncwarp();
    }
    cp_async_wait_group<0>();
}

// ---------------------------------------------------------------------------------------------
// Spatial attention forward on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// One CTA per (frame, head), two CTAs per SM (256 of the 512 TMEM columns each).  K^ (normalised, 64-byte rows,
// SWIZZLE_64B K-major = the tile_off() layout) and V^T (32 rows x keys, SWIZZLE_128B K-major) stay resident in shared
// memory; the CTA walks its query rows in M-tiles of 128 (= the 128 TMEM lanes) and the keys in tiles of 128:
//   MMA thread   : S  = Q^ K^T            tcgen05.mma SS, M128 x N128 x K32 (two K16 steps) -> TMEM cols [0,128)
//   softmax warps: thread = query row (no shuffles): tcgen05.ld S, + bias pair table, p = exp2(s - shift),
//                  row sum in a register, P as bf16 pairs -> tcgen05.st into TMEM cols [128,192)
//   MMA thread   : O += P V               tcgen05.mma TS (A = P from TMEM), M128 x N32, K16 per 16 keys -> cols [192,224)
// The softmax needs NO running maximum: q^ and k^ are l2-normalised, so every score is bounded by
// shift = scale * max|q_scale| * max|k_scale| + max|bias| (a property of the weights, supplied by the host), and
// softmax(s) = exp(s - shift) / sum exp(s - shift) exactly; with shift < 43 nothing can overflow or vanish in fp32.
// Hence there is no rescaling of O and no cross-lane reduction anywhere.
// ---------------------------------------------------------------------------------------------
static constexpr int TC_M = 128, TC_NT = 64;
static constexpr int TC_SOFTMAX_WARPS = 8;                 // two per TMEM lane quarter: each owns 32 of a tile's 64 keys
static constexpr int TC_WARP_MMA = 8, TC_WARP_LOAD = 9;
static constexpr int TC_THREADS = 320;
static constexpr uint32_t TC_TMEM_COLS = 256;              // S 2 x 64 | P 2 x 32 | O 2 x 32
static constexpr uint32_t TC_COL_S = 0, TC_COL_P = 128, TC_COL_O = 192;

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
CTC_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
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

// Pipeline (t = global key-tile counter of the CTA, b = t & 1 selects the S / P buffer):
//   MMA thread : S(t) -> s_full[b];  after p_full[b]: PV(t) -> pv_done[b], then S(t+2) into the S buffer just read
//   softmax    : wait s_full[b]; tcgen05.ld; exp2; wait pv_done[b] of tile t-2; tcgen05.st P(t) -> p_full[b]
// so the tensor core computes S(t+1) while the softmax warps work on S(t), and no warp waits on a barrier round trip.
// The stream of key tiles runs straight through the M-tile boundaries: O is double-buffered (o_free), the next Q tile
// is normalised by a loader warp into the other Q buffer (q_full / q_free), and there is no CTA barrier in the loop.
__global__ void __launch_bounds__(TC_THREADS, 2)
attn_tc_fwd_kernel(const AttnParams p, const float shift2) {
    extern __shared__ uint8_t sm_raw[];
    uint8_t* smb = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
    const int s = blockIdx.x, head = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.n, n_pad = p.n_pad;                               // n_pad: multiple of 64
    const int nW = 2 * p.W - 1, nb = (2 * p.H - 1) * nW;
    uint8_t* qs = smb;                                                // [2][128][64 B]       SWIZZLE_64B
    uint8_t* ks = qs + 2 * TC_M * 64;                                 // [n_pad][64 B]        SWIZZLE_64B
    uint8_t* vt = ks + n_pad * 64;                                    // [n_pad/64][32][128B] SWIZZLE_128B
    float2* pair = reinterpret_cast<float2*>(vt + n_pad * 64);        // [nb]
    int* tab8 = reinterpret_cast<int*>(pair + ((nb + 1) & ~1));       // [n_pad / 8], 16-byte aligned
    float* sv = reinterpret_cast<float*>(tab8 + ((n_pad / 8 + 3) & ~3));
    float* lsum = sv + 64;                                            // [128] row-sum exchange between the column halves
    uint64_t* bars = reinterpret_cast<uint64_t*>(lsum + TC_M);        // 6 x [2] barriers
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
    uint64_t* s_full = bars, *p_full = bars + 2, *pv_done = bars + 4, *q_full = bars + 6, *q_free = bars + 8,
              *o_free = bars + 10;

    if (threadIdx.x < 32) sv[threadIdx.x] = p.q_scale[threadIdx.x];
    else if (threadIdx.x < 64) sv[threadIdx.x] = p.k_scale[threadIdx.x - 32];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1); mbar_init(&p_full[b], TC_SOFTMAX_WARPS); mbar_init(&pv_done[b], 1);
            mbar_init(&q_full[b], 1); mbar_init(&q_free[b], 1); mbar_init(&o_free[b], TC_SOFTMAX_WARPS);
        }
        fence_barrier_init();
    }
    {   // bias pair table with the softmax shift folded in, and the per-8-key block index table
        const float* tb = p.bias_table + (long long)head * nb;
        for (int k = threadIdx.x; k < nb; k += blockDim.x)
            pair[k] = make_float2(tb[k] * LOG2E - shift2, (k > 0 ? tb[k - 1] * LOG2E : 0.f) - shift2);
        for (int jb = threadIdx.x; jb < n_pad / 8; jb += blockDim.x) {
            const int j = min(jb * 8, n - 8);
            tab8[jb] = (j / p.W) * nW + (j % p.W);
        }
    }
    __syncthreads();
    load_tile<true>(ks, 0, p.k, p.ldkv, p, s, head, 1, 0, n_pad, sv + 32, 1.0f);
    // V^T: element (d, key j) at block j/64, row d, 16-byte chunk ((j%64)/8) ^ (d%8), slot j%8
    for (int j = threadIdx.x; j < n_pad; j += blockDim.x) {
        uint4 c[4];
        if (j < n) {
            const uint4* g = reinterpret_cast<const uint4*>(p.v + seq_row(p, s, j) * p.ldkv + head * DH);
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = g[q];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = make_uint4(0, 0, 0, 0);
        }
        uint8_t* blk = vt + (j >> 6) * 4096 + (j & 7) * 2;
        const int ch = (j & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t w[4] = {c[q].x, c[q].y, c[q].z, c[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d0 = q * 8 + e * 2;
                *reinterpret_cast<uint16_t*>(blk + d0 * 128 + ((ch ^ (d0 & 7)) << 4)) = (uint16_t)(w[e] & 0xFFFFu);
                *reinterpret_cast<uint16_t*>(blk + (d0 + 1) * 128 + ((ch ^ ((d0 + 1) & 7)) << 4)) = (uint16_t)(w[e] >> 16);
            }
        }
    }
    tc_load_q(qs, p, s, head, 0, sv, threadIdx.x, blockDim.x);        // first Q tile by everybody
    if (warp == TC_WARP_MMA) tmem_alloc<TC_TMEM_COLS>(tmem_ptr);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int n_mt = (n + TC_M - 1) / TC_M, n_kt = n_pad / TC_NT;
    const int n_tiles = n_mt * n_kt;
    if (warp == TC_WARP_MMA) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(TC_M, TC_NT), idesc_o = make_idesc_bf16(TC_M, DH);
            auto issue_s = [&](int t) {                               // S(t) = Q(mt) K(kt)^T into S buffer t & 1
                const int mt = t / n_kt, kt = t - mt * n_kt;
                if (kt == 0 && mt > 0) {                              // first use of this M-tile's Q buffer
                    mbar_wait(&q_full[mt & 1], (((mt + 1) >> 1) - 1) & 1);
                    tcgen05_fence_after();
                }
                const uint64_t dq = make_umma_desc_sw64(smem_u32(qs + (mt & 1) * TC_M * 64));
                const uint64_t dk = make_umma_desc_sw64(smem_u32(ks + kt * TC_NT * 64));
                const uint32_t ts = tmem_base + TC_COL_S + (t & 1) * TC_NT;
                umma_f16_ss(ts, dq, dk, idesc_s, 0u);
                umma_f16_ss(ts, dq + 2, dk + 2, idesc_s, 1u);         // second K16 step: +32 B inside the 64 B row
                umma_commit(&s_full[t & 1]);
                if (kt == n_kt - 1) umma_commit(&q_free[mt & 1]);     // every S of this M-tile has been issued
            };
            issue_s(0);
            if (n_tiles > 1) issue_s(1);
            for (int t = 0; t < n_tiles; ++t) {
                const int mt = t / n_kt, kt = t - mt * n_kt;
                const uint32_t b = t & 1;
                mbar_wait(&p_full[b], (t >> 1) & 1);                  // P(t) is in TMEM, S(t) has been read
                if (kt == 0 && mt >= 2) mbar_wait(&o_free[mt & 1], ((mt >> 1) - 1) & 1);   // O of M-tile mt-2 was read
                tcgen05_fence_after();
                const uint64_t dv = make_umma_desc_sw128(smem_u32(vt + kt * 4096));
                const uint32_t tp = tmem_base + TC_COL_P + b * (TC_NT / 2);
                const uint32_t to = tmem_base + TC_COL_O + (mt & 1) * DH;
#pragma unroll
                for (int kk = 0; kk < TC_NT / 16; ++kk)
                    umma_f16_ts(to, tp + kk * 8, dv + (uint64_t)(kk * 2), idesc_o, (kt > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&pv_done[b]);
                if (t + 2 < n_tiles) issue_s(t + 2);
            }
        }
    } else if (warp == TC_WARP_LOAD) {
        // Q rows of M-tile m into buffer m & 1 as soon as the S MMAs of M-tile m - 2 no longer read it
        for (int m = 1; m < n_mt; ++m) {
            if (m >= 2) mbar_wait(&q_free[m & 1], ((m >> 1) - 1) & 1);
            tc_load_q(qs + (m & 1) * TC_M * 64, p, s, head, m * TC_M, sv, lane, 32);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_full[m & 1]);
        }
    } else {
        const int quarter = warp & 3, chalf = warp >> 2;              // TMEM lane quarter, column half of the key tile
        const int r = quarter * 32 + lane;                            // TMEM lane = query row of the tile
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        for (int mt = 0; mt < n_mt; ++mt) {
            const int i = mt * TC_M + r;
            const int base_i = bias_base(p, i);
            float l = 0.f;
            for (int kt = 0; kt < n_kt; ++kt) {
                const uint32_t t = mt * n_kt + kt, b = t & 1;
                mbar_wait(&s_full[b], (t >> 1) & 1);
                tcgen05_fence_after();
                // four 8-column loads, each in flight while the previous 8 scores are exponentiated: TMEM reads
                // (64 B/clk/SM) and the MUFU pipe (16 exp/clk/SM) have the same floor here and must overlap
                const uint32_t ts = tmem_base + TC_COL_S + b * TC_NT + lane_sel + chalf * 32;
                const int key0 = kt * TC_NT + chalf * 32;
                const int4 tb4 = *reinterpret_cast<const int4*>(tab8 + key0 / 8);
                const int tb[4] = {tb4.x, tb4.y, tb4.z, tb4.w};
                uint32_t pk[16];
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                uint32_t v[2][8];
                tmem_ld_32x32b_x8(ts, v[0]);
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    tmem_ld_wait();
                    if (bb < 3) tmem_ld_32x32b_x8(ts + (bb + 1) * 8, v[(bb + 1) & 1]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float2 f = pair[base_i - tb[bb] - 2 * u];
                        const float p0 = fast_exp2(__uint_as_float(v[bb & 1][2 * u]) + f.x);
                        const float p1 = fast_exp2(__uint_as_float(v[bb & 1][2 * u + 1]) + f.y);
                        if (u & 1) { l1 += p0; l3 += p1; } else { l0 += p0; l2 += p1; }
                        pk[bb * 4 + u] = pack_bf16(p0, p1);
                    }
                }
                l += (l0 + l1) + (l2 + l3);
                if (t >= 2) {                                         // P(t-2) (same buffer) consumed by its PV MMAs
                    mbar_wait(&pv_done[b], ((t >> 1) - 1) & 1);
                    tcgen05_fence_after();
                }
                tmem_st_32x32b_x16(tmem_base + TC_COL_P + b * (TC_NT / 2) + lane_sel + chalf * 16, pk);
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);
            }
            // combine the two column halves' row sums, then each half normalises and writes 16 of the 32 output dims
            if (chalf == 1) lsum[r] = l;
            named_bar_sync(1, TC_SOFTMAX_WARPS * 32);
            if (chalf == 0) lsum[r] = l = l + lsum[r];
            named_bar_sync(1, TC_SOFTMAX_WARPS * 32);
            l = lsum[r];
            const uint32_t tl = mt * n_kt + n_kt - 1;
            mbar_wait(&pv_done[tl & 1], (tl >> 1) & 1);               // the last PV commit covers every earlier MMA
            tcgen05_fence_after();
            uint32_t o[16];
            tmem_ld_32x32b_x16(tmem_base + TC_COL_O + (mt & 1) * DH + lane_sel + chalf * 16, o);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[mt & 1]);
            if (i < n) {
                const float inv = 1.f / l;
                const long long row = seq_row(p, s, i);
                uint4* orow = reinterpret_cast<uint4*>(p.out + row * (p.heads * DH) + head * DH + chalf * 16);
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    orow[q] = make_uint4(pack_bf16(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv),
                                         pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv),
                                         pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv),
                                         pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv));
                if (chalf == 0) p.lse[row * p.heads + head] = (log2f(l) + shift2) * LN2;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == TC_WARP_MMA) {
        tcgen05_fence_after();
        tmem_dealloc<TC_TMEM_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// Spatial attention backward, dQ, on tcgen05 / TMEM (same skeleton as attn_tc_fwd_kernel):
//   MMA thread   : S = Q^ K^T and dP = dO V^T      two SS MMAs per 64-key tile  -> TMEM S[b], dP[b]
//   softmax warps: thread = query row (lse_i, D_i in registers): p = exp2(s + bias - lse_i), dS = p (dP - D_i)
//                  as bf16 pairs -> tcgen05.st dS[b]
//   MMA thread   : dQ^ += dS K^                    TS MMA, B = K^T (32 x keys, SWIZZLE_128B)
// and the l2norm / q_scale adjoint of the finished row runs in the thread that owns it (no shuffles).
// One CTA per (frame, head) and SM (150 KB of resident K^, V, K^T + tables): 16 softmax warps, four per lane quarter.
// D_i = rowsum(dO o O) is produced by the loader warp together with the Q / dO tiles (and stored for the dK/dV kernel).
// ---------------------------------------------------------------------------------------------
static constexpr int TQ_SOFTMAX_WARPS = 16, TQ_WARP_MMA = 16, TQ_WARP_LOAD = 17, TQ_THREADS = 576;
static constexpr uint32_t TQ_TMEM_COLS = 512, TQ_COL_S = 0, TQ_COL_DP = 128, TQ_COL_DS = 256, TQ_COL_DQ = 320;

__global__ void __launch_bounds__(TQ_THREADS, 1)
attn_tc_bwd_dq_kernel(const AttnParams p) {
    extern __shared__ uint8_t sm_raw[];
    uint8_t* smb = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
    const int s = blockIdx.x, head = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.n, n_pad = p.n_pad;
    const int nW = 2 * p.W - 1, nb = (2 * p.H - 1) * nW;
    uint8_t* qs = smb;                                                // [2][128][64 B]  q^ * scale * log2e, SW64
    uint8_t* dos = qs + 2 * TC_M * 64;                                // [2][128][64 B]  dO, SW64
    uint8_t* ks = dos + 2 * TC_M * 64;                                // [n_pad][64 B]   k^, SW64
    uint8_t* vs = ks + n_pad * 64;                                    // [n_pad][64 B]   v, SW64
    uint8_t* ktr = vs + n_pad * 64;                                    // [n_pad/64][32][128 B]  k^ transposed, SW128
    float2* pair = reinterpret_cast<float2*>(ktr + n_pad * 64);
    int* tab8 = reinterpret_cast<int*>(pair + ((nb + 1) & ~1));
    float* sv = reinterpret_cast<float*>(tab8 + ((n_pad / 8 + 3) & ~3));
    float* dl = sv + 64;                                              // [2][128] D_i of the Q tile in each buffer
    uint64_t* bars = reinterpret_cast<uint64_t*>(dl + 2 * TC_M);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
    uint64_t* s_full = bars, *p_full = bars + 2, *pv_done = bars + 4, *q_full = bars + 6, *q_free = bars + 8,
              *o_free = bars + 10;
    const long long ldo = (long long)p.heads * DH;

    if (threadIdx.x < 32) sv[threadIdx.x] = p.q_scale[threadIdx.x];
    else if (threadIdx.x < 64) sv[threadIdx.x] = p.k_scale[threadIdx.x - 32];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1); mbar_init(&p_full[b], TQ_SOFTMAX_WARPS); mbar_init(&pv_done[b], 1);
            mbar_init(&q_full[b], 1); mbar_init(&q_free[b], 1); mbar_init(&o_free[b], 4);
        }
        fence_barrier_init();
    }
    {
        const float* tb = p.bias_table + (long long)head * nb;
        for (int k = threadIdx.x; k < nb; k += blockDim.x)
            pair[k] = make_float2(tb[k] * LOG2E, k > 0 ? tb[k - 1] * LOG2E : 0.f);
        for (int jb = threadIdx.x; jb < n_pad / 8; jb += blockDim.x) {
            const int j = min(jb * 8, n - 8);
            tab8[jb] = (j / p.W) * nW + (j % p.W);
        }
    }
    __syncthreads();
    load_tile<true>(ks, 0, p.k, p.ldkv, p, s, head, 1, 0, n_pad, sv + 32, 1.0f);
    load_tile<false>(vs, 0, p.v, p.ldkv, p, s, head, 1, 0, n_pad, nullptr, 1.0f);
    __syncthreads();
    // K^T from the normalised tile: element (d, key j) at block j/64, row d, chunk ((j%64)/8) ^ (d%8), slot j%8
    for (int j = threadIdx.x; j < n_pad; j += blockDim.x) {
        uint8_t* blk = ktr + (j >> 6) * 4096 + (j & 7) * 2;
        const int ch = (j & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 c = *reinterpret_cast<const uint4*>(ks + tile_off(j, q));
            const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d0 = q * 8 + e * 2;
                *reinterpret_cast<uint16_t*>(blk + d0 * 128 + ((ch ^ (d0 & 7)) << 4)) = (uint16_t)(w[e] & 0xFFFFu);
                *reinterpret_cast<uint16_t*>(blk + (d0 + 1) * 128 + ((ch ^ ((d0 + 1) & 7)) << 4)) = (uint16_t)(w[e] >> 16);
            }
        }
    }
    // Q / dO tile + D_i for M-tile m into buffer m & 1 (rows handled by `nthreads` threads starting at `tid`)
    auto load_rows = [&](int m, int tid, int nthreads) {
        const int buf = m & 1;
        tc_load_q(qs + buf * TC_M * 64, p, s, head, m * TC_M, sv, tid, nthreads);
        for (int r = tid; r < TC_M; r += nthreads) {
            const int i = m * TC_M + r;
            uint4 c[4];
            float d = 0.f;
            if (i < n) {
                const long long row = seq_row(p, s, i);
                const uint4* gd = reinterpret_cast<const uint4*>(p.d_o + row * ldo + head * DH);
                const uint4* go = reinterpret_cast<const uint4*>(p.o + row * ldo + head * DH);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    c[j] = gd[j];
                    const uint4 a = go[j];
                    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 x = unpack_bf16(aw[e]), y = unpack_bf16(bw[e]);
                        d += x.x * y.x + x.y * y.y;
                    }
                }
                p.delta[row * p.heads + head] = d;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) c[j] = make_uint4(0, 0, 0, 0);
            }
            dl[buf * TC_M + r] = d;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dos + buf * TC_M * 64 + tile_off(r, j)) = c[j];
        }
    };
    load_rows(0, threadIdx.x, blockDim.x);
    if (warp == TQ_WARP_MMA) tmem_alloc<TQ_TMEM_COLS>(tmem_ptr);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int n_mt = (n + TC_M - 1) / TC_M, n_kt = n_pad / TC_NT;
    const int n_tiles = n_mt * n_kt;
    if (warp == TQ_WARP_MMA) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(TC_M, TC_NT), idesc_o = make_idesc_bf16(TC_M, DH);
            auto issue_s = [&](int t) {
                const int mt = t / n_kt, kt = t - mt * n_kt;
                if (kt == 0 && mt > 0) {
                    mbar_wait(&q_full[mt & 1], (((mt + 1) >> 1) - 1) & 1);
                    tcgen05_fence_after();
                }
                const uint64_t dq_ = make_umma_desc_sw64(smem_u32(qs + (mt & 1) * TC_M * 64));
                const uint64_t dd = make_umma_desc_sw64(smem_u32(dos + (mt & 1) * TC_M * 64));
                const uint64_t dk = make_umma_desc_sw64(smem_u32(ks + kt * TC_NT * 64));
                const uint64_t dv = make_umma_desc_sw64(smem_u32(vs + kt * TC_NT * 64));
                const uint32_t ts = tmem_base + TQ_COL_S + (t & 1) * TC_NT, tdp = tmem_base + TQ_COL_DP + (t & 1) * TC_NT;
                umma_f16_ss(ts, dq_, dk, idesc_s, 0u);
                umma_f16_ss(ts, dq_ + 2, dk + 2, idesc_s, 1u);
                umma_f16_ss(tdp, dd, dv, idesc_s, 0u);
                umma_f16_ss(tdp, dd + 2, dv + 2, idesc_s, 1u);
                umma_commit(&s_full[t & 1]);
                if (kt == n_kt - 1) umma_commit(&q_free[mt & 1]);
            };
            issue_s(0);
            if (n_tiles > 1) issue_s(1);
            for (int t = 0; t < n_tiles; ++t) {
                const int mt = t / n_kt, kt = t - mt * n_kt;
                const uint32_t b = t & 1;
                mbar_wait(&p_full[b], (t >> 1) & 1);                  // dS(t) is in TMEM, S(t) / dP(t) have been read
                if (kt == 0 && mt >= 2) mbar_wait(&o_free[mt & 1], ((mt >> 1) - 1) & 1);
                tcgen05_fence_after();
                const uint64_t db = make_umma_desc_sw128(smem_u32(ktr + kt * 4096));
                const uint32_t tds = tmem_base + TQ_COL_DS + b * (TC_NT / 2);
                const uint32_t tdq = tmem_base + TQ_COL_DQ + (mt & 1) * DH;
#pragma unroll
                for (int kk = 0; kk < TC_NT / 16; ++kk)
                    umma_f16_ts(tdq, tds + kk * 8, db + (uint64_t)(kk * 2), idesc_o, (kt > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&pv_done[b]);
                if (t + 2 < n_tiles) issue_s(t + 2);
            }
        }
    } else if (warp == TQ_WARP_LOAD) {
        for (int m = 1; m < n_mt; ++m) {
            if (m >= 2) mbar_wait(&q_free[m & 1], ((m >> 1) - 1) & 1);
            load_rows(m, lane, 32);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_full[m & 1]);
        }
    } else {
        const int quarter = warp & 3, cpart = warp >> 2;              // TMEM lane quarter; 16-key slice of the 64-key tile
        const int r = quarter * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        for (int mt = 0; mt < n_mt; ++mt) {
            const int i = mt * TC_M + r;
            const int base_i = bias_base(p, i);
            if (mt > 0) {                                             // D_i of this tile is written by the loader warp
                mbar_wait(&q_full[mt & 1], (((mt + 1) >> 1) - 1) & 1);
            }
            const float d_i = dl[(mt & 1) * TC_M + r];
            const float lse2 = (i < n) ? p.lse[seq_row(p, s, i) * p.heads + head] * LOG2E : INFINITY;
            for (int kt = 0; kt < n_kt; ++kt) {
                const uint32_t t = mt * n_kt + kt, b = t & 1;
                mbar_wait(&s_full[b], (t >> 1) & 1);
                tcgen05_fence_after();
                uint32_t vs_[16], vd_[16];
                tmem_ld_32x32b_x16(tmem_base + TQ_COL_S + b * TC_NT + lane_sel + cpart * 16, vs_);
                tmem_ld_32x32b_x16(tmem_base + TQ_COL_DP + b * TC_NT + lane_sel + cpart * 16, vd_);
                const int key0 = kt * TC_NT + cpart * 16;
                const int tb0 = tab8[key0 / 8], tb1 = tab8[key0 / 8 + 1];
                tmem_ld_wait();
                uint32_t pk[8];
#pragma unroll
                for (int bb = 0; bb < 2; ++bb)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float2 f = pair[base_i - (bb ? tb1 : tb0) - 2 * u];
                        const int e = bb * 8 + 2 * u;
                        const float p0 = fast_exp2(__uint_as_float(vs_[e]) + f.x - lse2);
                        const float p1 = fast_exp2(__uint_as_float(vs_[e + 1]) + f.y - lse2);
                        pk[bb * 4 + u] = pack_bf16(p0 * (__uint_as_float(vd_[e]) - d_i), p1 * (__uint_as_float(vd_[e + 1]) - d_i));
                    }
                if (t >= 2) {
                    mbar_wait(&pv_done[b], ((t >> 1) - 1) & 1);
                    tcgen05_fence_after();
                }
                tmem_st_32x32b_x8(tmem_base + TQ_COL_DS + b * (TC_NT / 2) + lane_sel + cpart * 8, pk);
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);
            }
            if (cpart == 0) {                                         // one warp per lane quarter finishes the rows
                const uint32_t tl = mt * n_kt + n_kt - 1;
                mbar_wait(&pv_done[tl & 1], (tl >> 1) & 1);
                tcgen05_fence_after();
                uint32_t o[32];
                tmem_ld_32x32b_x32(tmem_base + TQ_COL_DQ + (mt & 1) * DH + lane_sel, o);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_free[mt & 1]);
                if (i < n) {
                    // adjoint of q^ = l2norm(q) * q_scale (times scale): dq = (g - x^ (x^ . g)) / |q|, g = dq^ * scale * q_scale
                    const long long row = seq_row(p, s, i);
                    const uint4* gq = reinterpret_cast<const uint4*>(p.q + row * p.ldq + head * DH);
                    float x[32];
                    float ss = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 c = gq[j];
                        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 tt = unpack_bf16(w[e]);
                            x[j * 8 + e * 2] = tt.x; x[j * 8 + e * 2 + 1] = tt.y;
                            ss += tt.x * tt.x + tt.y * tt.y;
                        }
                    }
                    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
                    float g[32];
                    float dot = 0.f;
#pragma unroll
                    for (int dd = 0; dd < 32; ++dd) {
                        g[dd] = __uint_as_float(o[dd]) * p.scale * sv[dd];
                        x[dd] *= inv;
                        dot += x[dd] * g[dd];
                    }
                    uint4* drow = reinterpret_cast<uint4*>(p.dq + row * p.lddq + head * DH);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int dd = j * 8 + e * 2;
                            w[e] = pack_bf16((g[dd] - x[dd] * dot) * inv, (g[dd + 1] - x[dd + 1] * dot) * inv);
                        }
                        drow[j] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == TQ_WARP_MMA) {
        tcgen05_fence_after();
        tmem_dealloc<TQ_TMEM_COLS>(tmem_base);
    }
}

// max_h ( scale * max|q_scale| * max|k_scale| + max|bias_h| ): the bound on every attention score (natural units)
__global__ void attn_score_bound_kernel(const float* __restrict__ q_scale, const float* __restrict__ k_scale, float scale,
                                        const float* __restrict__ bias_table, int n_bias, float* __restrict__ out) {
    __shared__ float red[32];
    float mq = 0.f, mk = 0.f, mb = 0.f;
    if (threadIdx.x < DH) { mq = fabsf(q_scale[threadIdx.x]); mk = fabsf(k_scale[threadIdx.x]); }
    for (int i = threadIdx.x; bias_table && i < n_bias; i += blockDim.x) mb = fmaxf(mb, fabsf(bias_table[i]));
    mq = warp_max(mq); mk = warp_max(mk); mb = warp_max(mb);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = mb;
    __syncthreads();
    if (warp == 0) {
        float m = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.f;
        m = warp_max(m);
        if (lane == 0) out[0] = scale * mq * mk + m;
    }
}

// ---------------------------------------------------------------------------------------------
// host side: configuration selection and launches
// ---------------------------------------------------------------------------------------------
static int fill_params(AttnParams& p, int B, int T, int H, int W, int heads, int mode, int kblk) {
    CTC_REQUIRE(mode == CTC_MODE_SPATIAL || mode == CTC_MODE_TEMPORAL, "attention: bad mode %d", mode);
    p.heads = heads; p.mode = mode; p.T = T; p.HW = H * W; p.H = H; p.W = W;
    p.n = (mode == CTC_MODE_SPATIAL) ? H * W : T;
    p.n_seq = (mode == CTC_MODE_SPATIAL) ? B * T : B * H * W;
    p.n_pad = (p.n + kblk - 1) / kblk * kblk;
    CTC_REQUIRE(p.n_pad <= 1024, "attention: sequence length %d exceeds the shared-memory resident design (1024)", p.n);
    return 0;
}

template <void (*kern)(const AttnParams)>
static int launch_attn(const AttnParams& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    static size_t configured = 0;   // one static per kernel (the kernel is a non-type template argument)
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // without this the driver sizes the shared-memory carve-out for ONE block (ncu: occupancy_limit_shared_mem = 1)
        CTC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    kern<<<grid, threads, smem, st>>>(p);
    CTC_LAUNCH_CHECK();
    return 0;
}

// small-sequence configuration (temporal, n <= 32) of the CTA-per-sequence kernels: only attention_probs uses them
// (forward / backward run on the warp-autonomous kernels above); two heads per CTA when the head count is even
static int small_hpc(int heads) { return heads % 2 == 0 ? 2 : 1; }
static bool use_small(const AttnParams& p) { return p.n <= 32 && p.bias_table == nullptr; }

// Row blocks of one (sequence, head) are walked by ONE CTA (gridDim.z = 1, resident K/V or Q/dO loaded and
// normalised once) whenever the other grid dimensions already fill the machine; small launches keep one CTA
// per block for parallelism.
static int attn_grid_z(const AttnParams& p, int qb) {
    const int blocks = (p.n + qb - 1) / qb;
    return ((long long)p.n_seq * p.heads >= 2ll * num_sms()) ? 1 : blocks;
}

// fast bias path: see load_bias_pairs
static bool fast_bias(const AttnParams& p) { return p.bias_table != nullptr && p.W % 8 == 0 && p.n >= 8; }

template <int QB, int KBLK, int HPC, bool PROBS>
static int run_fwd(AttnParams& p, cudaStream_t st) {
    const size_t nb = p.bias_table ? (size_t)(2 * p.H - 1) * (2 * p.W - 1) : 0;
    const size_t smem = (size_t)HPC * (p.n_pad * 128 + QB * 64) + 256 + p.n_pad * 4 + nb * 8;
    dim3 grid(p.n_seq, p.heads / HPC, attn_grid_z(p, QB));
    if constexpr (HPC == 1 && KBLK == 64) {
        if (fast_bias(p))
            return launch_attn<attn_fwd_kernel<QB, KBLK, HPC, PROBS, true>>(p, grid, HPC * (QB / 16) * 32, smem, st);
    }
    return launch_attn<attn_fwd_kernel<QB, KBLK, HPC, PROBS, false>>(p, grid, HPC * (QB / 16) * 32, smem, st);
}
// The tcgen05 dQ kernel measures 621 us against 597 us for the mma.sync kernel at batch 8 (both bound by per-score
// TMEM / shared-memory traffic at d_head = 32, profiles/r01_launches_step_b8_v2.md), so it is opt-in
// (ctc_attention_set_tc_bwd) and the mma.sync kernel stays the default.
static int g_tc_bwd = 0;
static bool tc_bwd_eligible(const AttnParams& p) {
    return g_tc_bwd && p.bias_table != nullptr && p.mode == CTC_MODE_SPATIAL && p.n % 64 == 0 && p.W % 8 == 0 && p.n >= 64 &&
           (size_t)p.n * 192 + 60 * 1024 <= 220 * 1024;
}
static int run_tc_bwd_dq(const AttnParams& p, cudaStream_t st);

template <int QB, int KBLK, int HPC>
static int run_bwd(AttnParams& p, cudaStream_t st) {
    const size_t nb = p.bias_table ? (size_t)(2 * p.H - 1) * (2 * p.W - 1) : 0;
    dim3 grid(p.n_seq, p.heads / HPC, attn_grid_z(p, QB));
    const size_t smem_dq = (size_t)HPC * (p.n_pad * 128 + 2 * QB * 64 + QB * 4) + 256 + p.n_pad * 4 + nb * 8;
    const size_t smem_dkv = (size_t)HPC * (p.n_pad * 128 + 2 * QB * 64 + 2 * p.n_pad * 4) + 256 + p.n_pad * 4 + nb * 8;
    const int threads = HPC * (QB / 16) * 32;
    if constexpr (HPC == 1 && KBLK == 64) {
        if (fast_bias(p)) {
            if (tc_bwd_eligible(p)) {           // dQ (and D = rowsum(dO o O)) on tcgen05 / TMEM
                if (int e = run_tc_bwd_dq(p, st)) return e;
            } else {
                if (int e = launch_attn<attn_bwd_dq_kernel<QB, KBLK, HPC, true>>(p, grid, threads, smem_dq, st)) return e;
            }
            return launch_attn<attn_bwd_dkv_kernel<QB, KBLK, HPC, true>>(p, grid, threads, smem_dkv, st);
        }
    }
    if (int e = launch_attn<attn_bwd_dq_kernel<QB, KBLK, HPC, false>>(p, grid, threads, smem_dq, st)) return e;
    return launch_attn<attn_bwd_dkv_kernel<QB, KBLK, HPC, false>>(p, grid, threads, smem_dkv, st);
}

static bool small_warp_path(const AttnParams& p) { return p.n <= SMALL_N && p.bias_table == nullptr; }
static int run_small_fwd(const AttnParams& p, cudaStream_t st) {
    const size_t smem = 256 + (size_t)SMALL_FWD_WARPS * 2 * 3 * SMALL_TILE;
    const int n_tasks = p.n_seq * p.heads;
    int grid = (n_tasks + SMALL_FWD_WARPS - 1) / SMALL_FWD_WARPS;
    if (grid > 2 * num_sms()) grid = 2 * num_sms();
    return launch_attn<attn_small_fwd_kernel>(p, dim3(grid), SMALL_FWD_WARPS * 32, smem, st);
}
static int run_small_bwd(const AttnParams& p, cudaStream_t st) {
    const size_t smem = 256 + (size_t)SMALL_BWD_WARPS * 2 * 4 * SMALL_TILE;
    const int n_tasks = p.n_seq * p.heads;
    int grid = (n_tasks + SMALL_BWD_WARPS - 1) / SMALL_BWD_WARPS;
    if (grid > num_sms()) grid = num_sms();
    return launch_attn<attn_small_bwd_kernel>(p, dim3(grid), SMALL_BWD_WARPS * 32, smem, st);
}

static int run_tc_bwd_dq(const AttnParams& p, cudaStream_t st) {
    const size_t nb = (size_t)(2 * p.H - 1) * (2 * p.W - 1);
    const size_t smem = 1024 + 4 * (size_t)TC_M * 64 + 3 * (size_t)p.n_pad * 64 + ((nb + 1) & ~(size_t)1) * 8 +
                        (((size_t)p.n_pad / 8 + 3) & ~(size_t)3) * 4 + 256 + 2 * TC_M * 4 + 128;
    static size_t configured = 0;
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_dq_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    attn_tc_bwd_dq_kernel<<<dim3(p.n_seq, p.heads), TQ_THREADS, smem, st>>>(p);
    CTC_LAUNCH_CHECK();
    return 0;
}

static int run_tc_fwd(const AttnParams& p, float score_bound, cudaStream_t st) {
    const size_t nb = (size_t)(2 * p.H - 1) * (2 * p.W - 1);
    const size_t smem = 1024 + 2 * (size_t)TC_M * 64 + (size_t)p.n_pad * 128 + ((nb + 1) & ~(size_t)1) * 8 +
                        (((size_t)p.n_pad / 8 + 3) & ~(size_t)3) * 4 + 256 + TC_M * 4 + 64;
    static size_t configured = 0;
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    attn_tc_fwd_kernel<<<dim3(p.n_seq, p.heads), TC_THREADS, smem, st>>>(p, score_bound * LOG2E);
    CTC_LAUNCH_CHECK();
    return 0;
}

}  // namespace ctc

using namespace ctc;

extern "C" int ctc_attention_set_tc_bwd(int on) {
    const int prev = g_tc_bwd;
    g_tc_bwd = on ? 1 : 0;
    return prev;
}

extern "C" int ctc_attention_score_bound(const float* q_scale, const float* k_scale, float scale,
                                         const float* bias_table, int heads, int H, int W, float* bound_dev,
                                         void* stream) {
    attn_score_bound_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(q_scale, k_scale, scale, bias_table,
                                                                 heads * (2 * H - 1) * (2 * W - 1), bound_dev);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_attention_fwd_tc(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int B, int T,
                                    int H, int W, int heads, const float* q_scale, const float* k_scale, float scale,
                                    const float* bias_table, float score_bound, void* o, float* lse, void* stream) {
    AttnParams p{};
    p.bias_table = bias_table;
    if (int e = fill_params(p, B, T, H, W, heads, CTC_MODE_SPATIAL, 64)) return e;
    CTC_REQUIRE(bias_table != nullptr && p.n % 64 == 0 && W % 8 == 0 && p.n_pad * 128 + 30000 <= 110 * 1024,
                "attention_fwd_tc: needs the spatial geometry (bias table, H*W %% 64 == 0, W %% 8 == 0, H*W <= 640); got "
                "H=%d W=%d", H, W);
    CTC_REQUIRE(score_bound > 0.f && score_bound < 43.f,
                "attention_fwd_tc: score bound %.2f outside (0, 43): the fixed-shift softmax is not safe, use ctc_attention_fwd",
                score_bound);
    p.q = (const __nv_bfloat16*)q; p.ldq = ldq; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
    p.ldkv = ldkv; p.q_scale = q_scale; p.k_scale = k_scale; p.scale = scale;
    p.out = (__nv_bfloat16*)o; p.lse = lse;
    return run_tc_fwd(p, score_bound, (cudaStream_t)stream);
}

extern "C" int ctc_attention_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int B, int T,
                                 int H, int W, int heads, const float* q_scale, const float* k_scale, float scale,
                                 const float* bias_table, int mode, void* o, float* lse, void* stream) {
    AttnParams p{};
    p.bias_table = bias_table;
    const int n = (mode == CTC_MODE_SPATIAL) ? H * W : T;
    const bool small = n <= 32 && bias_table == nullptr;
    if (int e = fill_params(p, B, T, H, W, heads, mode, small ? 32 : 64)) return e;
    p.q = (const __nv_bfloat16*)q; p.ldq = ldq; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
    p.ldkv = ldkv; p.q_scale = q_scale; p.k_scale = k_scale; p.scale = scale;
    p.out = (__nv_bfloat16*)o; p.lse = lse;
    cudaStream_t st = (cudaStream_t)stream;
    if (small && small_warp_path(p)) return run_small_fwd(p, st);
    if (small) {
        const int h = small_hpc(heads);
        return h >= 2 ? run_fwd<32, 32, 2, false>(p, st) : run_fwd<32, 32, 1, false>(p, st);
    }
    return (p.n > 128) ? run_fwd<192, 64, 1, false>(p, st) : run_fwd<64, 64, 1, false>(p, st);
}

extern "C" int ctc_attention_probs(const void* q, int64_t ldq, const void* k, int64_t ldkv, const float* lse, int B,
                                   int T, int H, int W, int heads, const float* q_scale, const float* k_scale,
                                   float scale, const float* bias_table, int mode, float* probs, void* stream) {
    AttnParams p{};
    p.bias_table = bias_table;
    const int n = (mode == CTC_MODE_SPATIAL) ? H * W : T;
    const bool small = n <= 32 && bias_table == nullptr;
    if (int e = fill_params(p, B, T, H, W, heads, mode, small ? 32 : 64)) return e;
    p.q = (const __nv_bfloat16*)q; p.ldq = ldq; p.k = (const __nv_bfloat16*)k; p.v = nullptr; p.ldkv = ldkv;
    p.q_scale = q_scale; p.k_scale = k_scale; p.scale = scale;
    p.lse = const_cast<float*>(lse); p.probs = probs;
    cudaStream_t st = (cudaStream_t)stream;
    if (small) {
        const int h = small_hpc(heads);
        return h >= 2 ? run_fwd<32, 32, 2, true>(p, st) : run_fwd<32, 32, 1, true>(p, st);
    }
    return (p.n > 128) ? run_fwd<192, 64, 1, true>(p, st) : run_fwd<64, 64, 1, true>(p, st);
}

extern "C" int ctc_attention_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* o,
                                 const void* d_o, const float* lse, int B, int T, int H, int W, int heads,
                                 const float* q_scale, const float* k_scale, float scale, const float* bias_table,
                                 int mode, void* dq, int64_t lddq, void* dk, void* dv, int64_t lddkv, float* delta_ws,
                                 void* stream) {
    AttnParams p{};
    p.bias_table = bias_table;
    const int n = (mode == CTC_MODE_SPATIAL) ? H * W : T;
    const bool small = n <= 32 && bias_table == nullptr;
    if (int e = fill_params(p, B, T, H, W, heads, mode, small ? 32 : 64)) return e;
    p.q = (const __nv_bfloat16*)q; p.ldq = ldq; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
    p.ldkv = ldkv; p.o = (const __nv_bfloat16*)o; p.d_o = (const __nv_bfloat16*)d_o;
    p.q_scale = q_scale; p.k_scale = k_scale; p.scale = scale;
    p.lse = const_cast<float*>(lse); p.delta = delta_ws;
    p.dq = (__nv_bfloat16*)dq; p.lddq = lddq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv; p.lddkv = lddkv;
    cudaStream_t st = (cudaStream_t)stream;
    if (small && small_warp_path(p)) return run_small_bwd(p, st);
    if (small) {
        const int h = small_hpc(heads);
        return h >= 2 ? run_bwd<32, 32, 2>(p, st) : run_bwd<32, 32, 1>(p, st);
    }
    return (p.n > 128) ? run_bwd<96, 64, 1>(p, st) : run_bwd<64, 64, 1>(p, st);
}
