// Spatial attention on the 5th-generation tensor cores (tcgen05 + TMEM): forward and the opt-in dQ backward.
// Reference numerics: Attention.forward, src/utils/attention.py:144-180.
#include "attention_tc.cuh"

namespace ctc {

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
static constexpr int TC_NT = 64;
static constexpr int TC_SOFTMAX_WARPS = 8;                 // two per TMEM lane quarter: each owns 32 of a tile's 64 keys
static constexpr int TC_WARP_MMA = 8, TC_WARP_LOAD = 9;
static constexpr int TC_THREADS = 320;
static constexpr uint32_t TC_TMEM_COLS = 256;              // S 2 x 64 | P 2 x 32 | O 2 x 32
static constexpr uint32_t TC_COL_S = 0, TC_COL_P = 128, TC_COL_O = 192;

// Pipeline (t = global key-tile counter of the CTA, b = t & 1 selects the S / P buffer):
//   MMA thread : S(t) -> s_full[b];  after p_full[b]: PV(t) -> pv_done[b], then S(t+2) into the S buffer just read
//   softmax    : wait s_full[b]; tcgen05.ld; exp2; wait pv_done[b] of tile t-2; tcgen05.st P(t) -> p_full[b]
// so the tensor core computes S(t+1) while the softmax warps work on S(t), and no warp waits on a barrier round trip.
// The stream of key tiles runs straight through the M-tile boundaries: O is double-buffered (o_free), the next Q tile
// is normalised by a loader warp into the other Q buffer (q_full / q_free), and there is no CTA barrier in the loop.
template <bool POLY>
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
        // the whole warp walks the tile stream, one elected lane issues (elect_one: descriptors stay in uniform registers)
        const bool issuer = elect_one();
        {
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
                if (issuer) umma_f16_ss(ts, dq, dk, idesc_s, 0u);
                if (issuer) umma_f16_ss(ts, dq + 2, dk + 2, idesc_s, 1u);         // second K16 step: +32 B inside the 64 B row
                if (issuer) umma_commit(&s_full[t & 1]);
                if (kt == n_kt - 1) if (issuer) umma_commit(&q_free[mt & 1]);     // every S of this M-tile has been issued
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
                    if (issuer) umma_f16_ts(to, tp + kk * 8, dv + (uint64_t)(kk * 2), idesc_o, (kt > 0 || kk > 0) ? 1u : 0u);
                if (issuer) umma_commit(&pv_done[b]);
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
                        const float x1 = __uint_as_float(v[bb & 1][2 * u + 1]) + f.y;
                        const float p1 = POLY ? exp2_poly(x1) : fast_exp2(x1);       // every other score off the MUFU pipe
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
        // the whole warp walks the tile stream, one elected lane issues (elect_one: descriptors stay in uniform registers)
        const bool issuer = elect_one();
        {
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
                if (issuer) umma_f16_ss(ts, dq_, dk, idesc_s, 0u);
                if (issuer) umma_f16_ss(ts, dq_ + 2, dk + 2, idesc_s, 1u);
                if (issuer) umma_f16_ss(tdp, dd, dv, idesc_s, 0u);
                if (issuer) umma_f16_ss(tdp, dd + 2, dv + 2, idesc_s, 1u);
                if (issuer) umma_commit(&s_full[t & 1]);
                if (kt == n_kt - 1) if (issuer) umma_commit(&q_free[mt & 1]);
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
                    if (issuer) umma_f16_ts(tdq, tds + kk * 8, db + (uint64_t)(kk * 2), idesc_o, (kt > 0 || kk > 0) ? 1u : 0u);
                if (issuer) umma_commit(&pv_done[b]);
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

// half of the forward softmax's exponentials on FMA-pipe polynomials (exp2_poly); ctc_attention_set_exp2_poly toggles it.
// OFF by default: measured on B200 at batch 8 (tools/kernel_bench.py attn_s, profiles/r02_kernel_bench.md) the split
// runs 388 us against 320 us with every exponential on MUFU - the kernel is issue-bound (9 FMA/ALU-pipe instructions
// replace one MUFU op in warps that already saturate their schedulers), not MUFU-bound.
int g_exp2_poly = 0;

int run_tc_bwd_dq(const AttnParams& p, cudaStream_t st) {
    const size_t nb = (size_t)(2 * p.H - 1) * (2 * p.W - 1);
    const size_t smem = 1024 + 4 * (size_t)TC_M * 64 + 3 * (size_t)p.n_pad * 64 + ((nb + 1) & ~(size_t)1) * 8 +
                        (((size_t)p.n_pad / 8 + 3) & ~(size_t)3) * 4 + 256 + 2 * TC_M * 4 + 128;
    static size_t configured_dev[kMaxDevices] = {};
    
    size_t& configured = configured_dev[current_device()];
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

int run_tc_fwd(const AttnParams& p, float score_bound, cudaStream_t st) {
    const size_t nb = (size_t)(2 * p.H - 1) * (2 * p.W - 1);
    const size_t smem = 1024 + 2 * (size_t)TC_M * 64 + (size_t)p.n_pad * 128 + ((nb + 1) & ~(size_t)1) * 8 +
                        (((size_t)p.n_pad / 8 + 3) & ~(size_t)3) * 4 + 256 + TC_M * 4 + 64;
    static size_t configured_dev[kMaxDevices] = {};
    
    size_t& configured = configured_dev[current_device()];
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    if (g_exp2_poly)
        attn_tc_fwd_kernel<true><<<dim3(p.n_seq, p.heads), TC_THREADS, smem, st>>>(p, score_bound * LOG2E);
    else
        attn_tc_fwd_kernel<false><<<dim3(p.n_seq, p.heads), TC_THREADS, smem, st>>>(p, score_bound * LOG2E);
    CTC_LAUNCH_CHECK();
    return 0;
}

int attn_score_bound(const float* q_scale, const float* k_scale, float scale, const float* bias_table, int n_bias,
                     float* out, cudaStream_t st) {
    attn_score_bound_kernel<<<1, 256, 0, st>>>(q_scale, k_scale, scale, bias_table, n_bias, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

}  // namespace ctc
