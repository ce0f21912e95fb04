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
#include "attention_common.cuh"

namespace ctc {

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
// host side: configuration selection and launches
// ---------------------------------------------------------------------------------------------
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
// Spatial attention backward: mode 2 (default) = ONE pass on tcgen05 / TMEM (attention_tc_bwd.cu: 1 010 us per layer at
// batch 8 against 1 245 us for the two mma.sync kernels in the same process, profiles/r02_kernel_bench_v5.log - once its
// MMA-issue warp ran converged with uniform-register descriptors and its softmax warps in two groups on alternating
// tiles); mode 0 = the two mma.sync kernels (dQ, dK/dV: the path for every other geometry and the cross-check of the
// tests); mode 1 = tcgen05 dQ + mma.sync dK/dV (621 us against 597 us for dQ: never the default).
// ctc_attention_set_tc_bwd / CTC_ATTN_BWD in the environment of the Python binding select a mode.
static int g_tc_bwd = 2;
static bool tc_bwd_eligible(const AttnParams& p) {
    return g_tc_bwd == 1 && p.bias_table != nullptr && p.mode == CTC_MODE_SPATIAL && p.n % 64 == 0 && p.W % 8 == 0 && p.n >= 64 &&
           (size_t)p.n * 192 + 60 * 1024 <= 220 * 1024;
}

template <int QB, int KBLK, int HPC>
static int run_bwd(AttnParams& p, cudaStream_t st) {
    const size_t nb = p.bias_table ? (size_t)(2 * p.H - 1) * (2 * p.W - 1) : 0;
    dim3 grid(p.n_seq, p.heads / HPC, attn_grid_z(p, QB));
    const size_t smem_dq = (size_t)HPC * (p.n_pad * 128 + 2 * QB * 64 + QB * 4) + 256 + p.n_pad * 4 + nb * 8;
    const size_t smem_dkv = (size_t)HPC * (p.n_pad * 128 + 2 * QB * 64 + 2 * p.n_pad * 4) + 256 + p.n_pad * 4 + nb * 8;
    const int threads = HPC * (QB / 16) * 32;
    if constexpr (HPC == 1 && KBLK == 64) {
        // g_tc_bwd == 2: dQ, dK and dV in one pass on tcgen05 / TMEM (attention_tc_bwd.cu)
        if (g_tc_bwd == 2 && tc_bwd_onepass_eligible(p)) return run_tc_bwd_onepass(p, st);
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


}  // namespace ctc

using namespace ctc;

extern "C" int ctc_attention_set_tc_bwd(int mode) {
    const int prev = g_tc_bwd;
    g_tc_bwd = (mode == 1 || mode == 2) ? mode : 0;
    return prev;
}

namespace ctc { extern int g_exp2_poly; }
extern "C" int ctc_attention_set_exp2_poly(int on) {
    const int prev = g_exp2_poly;
    g_exp2_poly = on ? 1 : 0;
    return prev;
}

extern "C" int ctc_attention_score_bound(const float* q_scale, const float* k_scale, float scale,
                                         const float* bias_table, int heads, int H, int W, float* bound_dev,
                                         void* stream) {
    return attn_score_bound(q_scale, k_scale, scale, bias_table, heads * (2 * H - 1) * (2 * W - 1), bound_dev,
                            (cudaStream_t)stream);
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
