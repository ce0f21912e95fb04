// Spatial attention backward in ONE pass on tcgen05 / TMEM: dQ, dK and dV of a (frame, head) from a single
// recomputation of the probabilities (the mma.sync path recomputes S and P twice, in a dQ and a dK/dV kernel).
// Reference numerics: the adjoint of Attention.forward, src/utils/attention.py:144-180.
//
// One CTA per (frame, head) and SM.  Tiles: 128 query rows (= the 128 TMEM lanes) x 32 keys; the KEY tile is the
// outer loop, the query M-tile the inner one, so that
//   dQ   [M-tile][128 x 32]  accumulates over the outer loop in n_mt <= 5 resident TMEM accumulators (160 columns),
//   dV^T, dK^T [32 x 32 keys] accumulate over the inner loop (double-buffered, drained once per key tile).
// Per tile t (b = t & 1); the MMA warp runs converged and one elected lane issues (common.cuh: elect_one):
//   MMA thread   : S = Q^ K^T, dP = dO V^T                (SS, M128 N32 K32)            -> TMEM S[b], dP[b]
//   softmax warps: two groups of 8 warps on ALTERNATING tiles (group = b), thread = query row i, 16 keys per warp
//                  (lse_i, D_i from shared memory): p = exp2(s + bias - lse_i), dS = p (dP - D_i)
//                  dS as bf16 -> TMEM dS[b] (A operand of the dQ MMA); P and dS as bf16 -> shared-memory staging tiles [b]
//                  in [key][query] order (K-major B operands of the dV / dK MMAs)
//   MMA thread   : dQ[mt]  += dS K^                        (TS, A = dS from TMEM, B = K^ transposed tile)
//                  [dV^T ; dK^T][kt] += [dO^T ; Q^^T] [P ; dS]^T   ONE SS stream, M128 N64 K128: the A operand stacks the
//                  resident transposed copies (rows 0-31 = dO^T, rows 32-63 = Q^^T of a 64-query block; rows 64-127 read
//                  the next block and only produce accumulator lanes nobody reads), the B operand stacks the staging
//                  tiles (rows 0-31 = P, 32-63 = dS), so accumulator lanes 0-31 x columns 0-31 hold dV^T and lanes 32-63
//                  x columns 32-63 hold dK^T (the two off-diagonal blocks are by-products nobody reads)
//   drain warps  : per key tile, dV^T / dK^T -> transpose through shared memory -> dv rows, l2norm / k_scale adjoint -> dk
// S / dP of tile t+2 are issued as soon as the softmax warps have READ tile t (s_free), i.e. ahead of the dQ / dV / dK
// products of tile t: the first version issued them behind those 18 MMAs and ncu showed 25 % of all warp samples on the
// s_full wait (profiles/r02_ncu_attn_onepass.md).
//   loader warp  : streams the 32-key K^ / V / K^T tiles (normalised on the fly) two tiles ahead
// Everything that is summed is summed by the tensor core in issue order: no atomics, bit-reproducible.
// D_i = rowsum(dO o O) and lse are staged once per CTA together with the resident Q^ / dO tiles and their transposes.
#include "attention_tc.cuh"

namespace ctc {

static constexpr int OB_NK = 32;                                   // keys per tile
static constexpr int OB_SOFTMAX_WARPS = 16;                        // four per TMEM lane quarter, 8 keys of the tile each
static constexpr int OB_WARP_DRAIN_V = 16, OB_WARP_DRAIN_K = 17;    // warp % 4 = 0 / 1: TMEM lanes 0-31 (dV^T) / 32-63 (dK^T)
static constexpr int OB_WARP_MMA = 18, OB_WARP_LOAD = 19;
static constexpr int OB_THREADS = 20 * 32;
static constexpr int OB_STG = 2 * 8192;                              // bytes of one P / dS staging buffer (128 queries)
static constexpr int OB_MAX_MT = 5;                                // n <= 640
static constexpr uint32_t OB_TMEM_COLS = 512;
static constexpr uint32_t OB_COL_S = 0, OB_COL_DP = 64, OB_COL_DS = 128, OB_COL_DQ = 160, OB_COL_DVK = 320;   // dVK: 2 x 64

CTC_DEVINL void tmem_st_32x32b_x4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
}

struct ObSmem {                                                    // byte offsets from the 1024-byte aligned base
    int qs, dos, tT, stg, kt, pair, tab8, lse2, dl, sv, tr, bars, total;
};
CTC_DEVINL ObSmem __host__ ob_layout(int n_rows, int nb) {
    ObSmem L;
    int o = 0;
    L.qs = o;   o += n_rows * 64;                                  // [n_rows][64 B]  q^ * scale * log2e   SWIZZLE_64B
    L.dos = o;  o += n_rows * 64;                                  // [n_rows][64 B]  dO                   SWIZZLE_64B
    L.tT = o;   o += (n_rows / 64) * 8192;                         // [n_rows/64][64 rows: dO^T | (q^)^T][128 B]  SWIZZLE_128B
    L.stg = o;  o += 2 * OB_STG;                                   // [2 buffers][2 blocks of 64 queries][64 rows: P keys | dS keys][128 B]
    L.kt = o;   o += 2 * 6144;                                     // [2] x { K^ 32 x 64 B | V | (K^)^T }  SWIZZLE_64B
    L.pair = o; o += ((nb + 1 + 3) & ~3) * 4;                      // fp32 bias table * log2e, one leading zero (index -1)
    L.tab8 = o; o += ((n_rows / 8 + 3) & ~3) * 4;
    L.lse2 = o; o += n_rows * 4;
    L.dl = o;   o += n_rows * 4;
    L.sv = o;   o += 256;
    L.tr = o;   o += 2 * 32 * 33 * 4;
    o = (o + 15) & ~15;
    L.bars = o; o += 256;
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(OB_THREADS, 1)
attn_tc_bwd_onepass_kernel(const AttnParams p) {
    extern __shared__ uint8_t sm_raw[];
    uint8_t* smb = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
    const int s = blockIdx.x, head = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.n;
    const int n_mt = (n + TC_M - 1) / TC_M, n_rows = n_mt * TC_M, n_kt = n / OB_NK;
    const int nW = 2 * p.W - 1, nb = (2 * p.H - 1) * nW;
    const ObSmem L = ob_layout(n_rows, nb);
    {
        uint32_t dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if ((uint32_t)(smb - sm_raw) + (uint32_t)L.total > dyn) __trap();
    }
    uint8_t* qs = smb + L.qs;   uint8_t* dos = smb + L.dos;
    uint8_t* tT = smb + L.tT;
    uint8_t* stg = smb + L.stg;
    uint8_t* ktl = smb + L.kt;
    float* tabf = reinterpret_cast<float*>(smb + L.pair) + 1;      // tabf[-1] = 0
    int* tab8 = reinterpret_cast<int*>(smb + L.tab8);
    float* lse2 = reinterpret_cast<float*>(smb + L.lse2);
    float* dl = reinterpret_cast<float*>(smb + L.dl);
    float* sv = reinterpret_cast<float*>(smb + L.sv);
    float* tr = reinterpret_cast<float*>(smb + L.tr);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smb + L.bars);
    uint64_t *s_full = bars, *s_free = bars + 2, *p_full = bars + 4, *ds_free = bars + 6, *k_full = bars + 8,
             *k_free = bars + 10, *kv_full = bars + 12, *kv_free = bars + 14, *stg_free = bars + 16, *dq_done = bars + 18;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 19);
    const long long ldo = (long long)p.heads * DH;

    if (threadIdx.x < 32) sv[threadIdx.x] = p.q_scale[threadIdx.x];
    else if (threadIdx.x < 64) sv[threadIdx.x] = p.k_scale[threadIdx.x - 32];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1);  mbar_init(&s_free[b], OB_SOFTMAX_WARPS / 2);
            mbar_init(&p_full[b], OB_SOFTMAX_WARPS / 2);  mbar_init(&ds_free[b], 1);
            mbar_init(&k_full[b], 1);  mbar_init(&k_free[b], 1);
            mbar_init(&kv_full[b], 1); mbar_init(&kv_free[b], 2);
            mbar_init(&stg_free[b], 1);
        }
        mbar_init(dq_done, 1);
        fence_barrier_init();
    }
    {
        const float* tb = p.bias_table + (long long)head * nb;
        for (int k = threadIdx.x; k < nb; k += blockDim.x)
            tabf[k] = tb[k] * LOG2E;
        if (threadIdx.x == 0) tabf[-1] = 0.f;
        for (int jb = threadIdx.x; jb < n_rows / 8; jb += blockDim.x) {
            const int j = min(jb * 8, n - 8);
            tab8[jb] = (j / p.W) * nW + (j % p.W);
        }
    }
    __syncthreads();
    // ---- resident query-side tiles: q^ (scaled), dO, D_i, lse_i
    for (int m = 0; m < n_mt; ++m) tc_load_q(qs + m * TC_M * 64, p, s, head, m * TC_M, sv, threadIdx.x, blockDim.x);
    for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
        uint4 c[4];
        float d = 0.f, l2 = INFINITY;                              // lse = +inf for padded rows -> P = dS = 0
        if (r < n) {
            const long long row = seq_row(p, s, r);
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
            l2 = p.lse[row * p.heads + head] * LOG2E;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = make_uint4(0, 0, 0, 0);
        }
        dl[r] = d;
        lse2[r] = l2;
        uint8_t* tile = dos + (r >> 7) * (TC_M * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(tile + tile_off(r & 127, j)) = c[j];
    }
    __syncthreads();
    // ---- their transposes: element (d, row r) at block r/64, row d, 16-byte chunk ((r%64)/8) ^ (d%8), slot r%8
    for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const int ch = (r & 63) >> 3;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const uint8_t* src = (which ? dos : qs) + (r >> 7) * (TC_M * 64);
            uint8_t* blk = tT + (r >> 6) * 8192 + (which ? 0 : 4096) + (r & 7) * 2;     // rows 0-31 dO^T, rows 32-63 (q^)^T
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 c = *reinterpret_cast<const uint4*>(src + tile_off(r & 127, q));
                const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int d0 = q * 8 + e * 2;
                    *reinterpret_cast<uint16_t*>(blk + d0 * 128 + ((ch ^ (d0 & 7)) << 4)) = (uint16_t)(w[e] & 0xFFFFu);
                    *reinterpret_cast<uint16_t*>(blk + (d0 + 1) * 128 + ((ch ^ ((d0 + 1) & 7)) << 4)) = (uint16_t)(w[e] >> 16);
                }
            }
        }
    }
    if (warp == OB_WARP_MMA) tmem_alloc<OB_TMEM_COLS>(tmem_ptr);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_tiles = n_kt * n_mt;

    if (warp == OB_WARP_MMA) {
        // the whole warp walks the tile stream, one elected lane issues (elect_one: descriptors stay in uniform registers)
        const bool issuer = elect_one();
        {
            const uint32_t idesc = make_idesc_bf16(TC_M, OB_NK), idesc_vk = make_idesc_bf16(TC_M, 2 * OB_NK);
            auto issue_s = [&](int t, int kt, int mt) {               // S(t), dP(t) into buffer t & 1
                const int b = t & 1;
                if (mt == 0) { mbar_wait(&k_full[kt & 1], (kt >> 1) & 1); tcgen05_fence_after(); }
                if (t >= 2) { mbar_wait(&s_free[b], ((t >> 1) - 1) & 1); tcgen05_fence_after(); }
                const uint8_t* kb = ktl + (kt & 1) * 6144;
                const uint64_t dq_ = make_umma_desc_sw64(smem_u32(qs + mt * TC_M * 64));
                const uint64_t dd = make_umma_desc_sw64(smem_u32(dos + mt * TC_M * 64));
                const uint64_t dk = make_umma_desc_sw64(smem_u32(kb));
                const uint64_t dv = make_umma_desc_sw64(smem_u32(kb + 2048));
                const uint32_t ts = tmem_base + OB_COL_S + b * OB_NK, tdp = tmem_base + OB_COL_DP + b * OB_NK;
                if (issuer) umma_f16_ss(ts, dq_, dk, idesc, 0u);
                if (issuer) umma_f16_ss(ts, dq_ + 2, dk + 2, idesc, 1u);          // second K16 step: +32 B inside the 64 B row
                if (issuer) umma_f16_ss(tdp, dd, dv, idesc, 0u);
                if (issuer) umma_f16_ss(tdp, dd + 2, dv + 2, idesc, 1u);
                if (issuer) umma_commit(&s_full[b]);
            };
            // (kt, mt) of tile t + 2, advanced alongside the loop (no integer divisions in the tile stream)
            int kt2 = 0, mt2 = 0;
            auto advance2 = [&]() { if (++mt2 == n_mt) { mt2 = 0; ++kt2; } };
            issue_s(0, kt2, mt2); advance2();
            if (n_tiles > 1) { issue_s(1, kt2, mt2); advance2(); }
            // With ONE M-tile per key tile, tile t + 2 belongs to key tile kt + 2, whose K buffer is the one tile t still
            // reads: issuing it ahead would wait on a k_full that only this thread's own k_free commit can produce.
            const bool ahead = n_mt >= 2;
            int t = 0;
            for (int kt = 0; kt < n_kt; ++kt) {
                const uint8_t* kb = ktl + (kt & 1) * 6144;
                const uint64_t dkT = make_umma_desc_sw64(smem_u32(kb + 4096));
                const uint32_t tvk = tmem_base + OB_COL_DVK + (kt & 1) * (2 * OB_NK);
                for (int mt = 0; mt < n_mt; ++mt, ++t) {
                    const int b = t & 1;
                    // S / dP of tile t + 2 first: they only need the softmax warps to have READ tile t (s_free)
                    if (ahead && t + 2 < n_tiles) { issue_s(t + 2, kt2, mt2); advance2(); }
                    mbar_wait(&p_full[b], (t >> 1) & 1);              // dS(t) in TMEM, P / dS staged
                    if (mt == 0 && kt >= 2) mbar_wait(&kv_free[kt & 1], ((kt >> 1) - 1) & 1);   // accumulators drained
                    tcgen05_fence_after();
                    // dQ[mt] += dS K^
                    const uint32_t tds = tmem_base + OB_COL_DS + b * (OB_NK / 2);
                    const uint32_t tdq = tmem_base + OB_COL_DQ + mt * DH;
#pragma unroll
                    for (int kk = 0; kk < OB_NK / 16; ++kk)
                        if (issuer) umma_f16_ts(tdq, tds + kk * 8, dkT + (uint64_t)(kk * 2), idesc, (kt > 0 || kk > 0) ? 1u : 0u);
                    if (issuer) umma_commit(&ds_free[b]);
                    // [dV^T ; dK^T][kt] += [dO^T ; Q^^T](mt) [P ; dS](t)^T    (reduction over the 128 queries of the M-tile)
                    const uint64_t a0 = make_umma_desc_sw128(smem_u32(tT) + (uint32_t)(mt * 2) * 8192u);
                    const uint64_t b0 = make_umma_desc_sw128(smem_u32(stg) + (uint32_t)(b * OB_STG));
#pragma unroll
                    for (int kk = 0; kk < TC_M / 16; ++kk) {
                        // +512 in the (addr >> 4) field = the next 8 KB block; +2 = the next 16 queries (32 B) of a row
                        const uint64_t off = (uint64_t)((kk >> 2) * 512 + (kk & 3) * 2);
                        if (issuer) umma_f16_ss(tvk, a0 + off, b0 + off, idesc_vk, (mt > 0 || kk > 0) ? 1u : 0u);
                    }
                    if (issuer) umma_commit(&stg_free[b]);
                    if (mt == n_mt - 1) { if (issuer) umma_commit(&kv_full[kt & 1]); if (issuer) umma_commit(&k_free[kt & 1]); }
                    if (!ahead && t + 2 < n_tiles) { issue_s(t + 2, kt2, mt2); advance2(); }
                }
            }
            if (issuer) umma_commit(dq_done);
        }
    } else if (warp == OB_WARP_LOAD) {
        // K^ (normalised), V and (K^)^T of key tile kt; lane = key
        for (int kt = 0; kt < n_kt; ++kt) {
            if (kt >= 2) mbar_wait(&k_free[kt & 1], ((kt >> 1) - 1) & 1);
            uint8_t* kb = ktl + (kt & 1) * 6144;
            const long long row = seq_row(p, s, kt * OB_NK + lane);
            const uint4* gk = reinterpret_cast<const uint4*>(p.k + row * p.ldkv + head * DH);
            const uint4* gv = reinterpret_cast<const uint4*>(p.v + row * p.ldkv + head * DH);
            uint4 c[4], v4[4];
            float f[32];
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c[j] = gk[j];
                v4[j] = gv[j];
                const uint32_t w[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 tt = unpack_bf16(w[e]);
                    f[j * 8 + e * 2] = tt.x; f[j * 8 + e * 2 + 1] = tt.y;
                    ss += tt.x * tt.x + tt.y * tt.y;
                }
            }
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            const int ch = lane >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int d0 = j * 8 + e * 2;
                    w[e] = pack_bf16(f[d0] * inv * sv[32 + d0], f[d0 + 1] * inv * sv[32 + d0 + 1]);
                    // (K^)^T tile: element (row d, key lane) in 64-byte rows, SWIZZLE_64B
                    *reinterpret_cast<uint16_t*>(kb + 4096 + tile_off(d0, ch) + (lane & 7) * 2) = (uint16_t)(w[e] & 0xFFFFu);
                    *reinterpret_cast<uint16_t*>(kb + 4096 + tile_off(d0 + 1, ch) + (lane & 7) * 2) = (uint16_t)(w[e] >> 16);
                }
                *reinterpret_cast<uint4*>(kb + tile_off(lane, j)) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(kb + 2048 + tile_off(lane, j)) = v4[j];
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&k_full[kt & 1]);
        }
    } else if (warp == OB_WARP_DRAIN_V || warp == OB_WARP_DRAIN_K) {
        // dV^T (TMEM lanes 0-31, columns 0-31) / dK^T (lanes 32-63, columns 32-63) of key tile kt: lane = head dim
        const bool is_k = warp == OB_WARP_DRAIN_K;
        float* trw = tr + (is_k ? 32 * 33 : 0);
        const uint32_t tsel = is_k ? (((uint32_t)32 << 16) + OB_NK) : 0u;
        for (int kt = 0; kt < n_kt; ++kt) {
            const int b2 = kt & 1;
            mbar_wait(&kv_full[b2], (kt >> 1) & 1);
            tcgen05_fence_after();
            uint32_t av[32];
            tmem_ld_32x32b_x32(tmem_base + OB_COL_DVK + b2 * (2 * OB_NK) + tsel, av);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_free[b2]);
            const long long row = seq_row(p, s, kt * OB_NK + lane);      // after the transpose: lane = key
#pragma unroll
            for (int c = 0; c < 32; ++c) trw[c * 33 + lane] = __uint_as_float(av[c]);
            __syncwarp();
            if (!is_k) {
                uint4* drow = reinterpret_cast<uint4*>(p.dv + row * p.lddkv + head * DH);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) w[e] = pack_bf16(trw[lane * 33 + j * 8 + e * 2], trw[lane * 33 + j * 8 + e * 2 + 1]);
                    drow[j] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
                // accumulated against q^ * scale * log2e -> undo log2e; k_scale and the l2norm adjoint
                const uint4* gk = reinterpret_cast<const uint4*>(p.k + row * p.ldkv + head * DH);
                float x[32], g[32];
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 c = gk[j];
                    const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 tt = unpack_bf16(w[e]);
                        x[j * 8 + e * 2] = tt.x; x[j * 8 + e * 2 + 1] = tt.y;
                        ss += tt.x * tt.x + tt.y * tt.y;
                    }
                }
                const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
                float dot = 0.f;
#pragma unroll
                for (int dd = 0; dd < 32; ++dd) {
                    g[dd] = trw[lane * 33 + dd] * LN2 * sv[32 + dd];
                    x[dd] *= inv;
                    dot += x[dd] * g[dd];
                }
                uint4* drow = reinterpret_cast<uint4*>(p.dk + row * p.lddkv + head * DH);
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
            __syncwarp();
        }
    } else {
        // Two groups of 8 softmax warps work on ALTERNATING tiles (group = tile parity = S / dP / dS / staging buffer), so
        // that the latency chain of one tile (barrier wake-up, TMEM loads, exponentials, TMEM store, staging stores, proxy
        // fence) overlaps the next tile's instead of all 16 warps walking it tile after tile (ncu: both pipes ~35 % busy).
        const int quarter = warp & 3, grp = warp >> 3, half = (warp >> 2) & 1;   // TMEM lane quarter; tile parity; 16-key half
        const int r = quarter * 32 + lane;                            // query row of the M-tile = TMEM lane
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        const int b = grp;
        // per-thread constants of the tile stream: the bias-table base of this thread's row in every M-tile and the staging
        // addresses of its keys (P row c, dS row 32 + c of the 64-query block r / 64); key k + 8 sits 8 rows further
        int base_m[OB_MAX_MT];
#pragma unroll
        for (int m = 0; m < OB_MAX_MT; ++m) base_m[m] = bias_base(p, m * TC_M + r);
        uint32_t soff[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            soff[k] = smem_u32(stg) + (uint32_t)(b * OB_STG + (r >> 6) * 8192 + (half * 16 + k) * 128 + ((((r & 63) >> 3) ^ k) << 4) + (r & 7) * 2);
        const uint32_t ts0 = tmem_base + OB_COL_S + lane_sel + b * OB_NK + half * 16;
        const uint32_t tdp0 = tmem_base + OB_COL_DP + lane_sel + b * OB_NK + half * 16;
        const uint32_t tds0 = tmem_base + OB_COL_DS + lane_sel + b * (OB_NK / 2) + half * 8;
        int kt = 0, mt = grp;
        if (mt >= n_mt) { mt -= n_mt; ++kt; }                          // n_mt == 1
#pragma unroll 1
        for (int t = grp; t < n_tiles; t += 2) {
            const int it = t >> 1;                                    // this group's tile counter = use count of buffer b
            const int tb0 = tab8[kt * (OB_NK / 8) + half * 2], tb1 = tab8[kt * (OB_NK / 8) + half * 2 + 1];
            const int i = mt * TC_M + r;
            // select chain instead of base_m[mt]: a dynamically indexed array would live in local memory
            const int base_i = mt == 0 ? base_m[0] : mt == 1 ? base_m[1] : mt == 2 ? base_m[2] : mt == 3 ? base_m[3] : base_m[4];
            const float lse2_i = lse2[i], d_i = dl[i];
            mbar_wait(&s_full[b], it & 1);
            tcgen05_fence_after();
            uint32_t vs_[16], vd_[16];
            tmem_ld_32x32b_x16(ts0, vs_);
            tmem_ld_32x32b_x16(tdp0, vd_);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[b]);                   // S / dP buffer b may be overwritten (tile t + 2)
            uint32_t pkP[8], pkD[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float* prow = tabf + (base_i - (u < 4 ? tb0 : tb1));
                const int uu = u & 3;
                const float p0 = fast_exp2(__uint_as_float(vs_[2 * u]) + prow[-2 * uu] - lse2_i);
                const float p1 = fast_exp2(__uint_as_float(vs_[2 * u + 1]) + prow[-2 * uu - 1] - lse2_i);
                pkP[u] = pack_bf16(p0, p1);
                pkD[u] = pack_bf16(p0 * (__uint_as_float(vd_[2 * u]) - d_i), p1 * (__uint_as_float(vd_[2 * u + 1]) - d_i));
            }
            if (it >= 1) {                                            // dS(t-2) consumed by its dQ MMAs
                mbar_wait(&ds_free[b], (it - 1) & 1);
                tcgen05_fence_after();
            }
            tmem_st_32x32b_x4(tds0, reinterpret_cast<const uint32_t(&)[4]>(pkD[0]));
            tmem_st_32x32b_x4(tds0 + 4, reinterpret_cast<const uint32_t(&)[4]>(pkD[4]));
            // staging buffer b consumed by the dV / dK MMAs of tile t - 2 (the MMAs of tile t - 1 read the other buffer)
            if (it >= 1) mbar_wait(&stg_free[b], (it - 1) & 1);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // keys 2u, 2u + 1 of this warp's 16: staging rows (half * 16 + key); st.shared.u16 stores the low register half
                const uint32_t a0 = soff[(2 * u) & 7] + (uint32_t)((2 * u) >> 3) * 1024u, a1 = soff[(2 * u + 1) & 7] + (uint32_t)((2 * u + 1) >> 3) * 1024u;
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a0), "h"((uint16_t)(pkP[u] & 0xFFFFu)) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a1), "h"((uint16_t)(pkP[u] >> 16)) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a0 + 4096u), "h"((uint16_t)(pkD[u] & 0xFFFFu)) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a1 + 4096u), "h"((uint16_t)(pkD[u] >> 16)) : "memory");
            }
            tmem_st_wait();
            fence_proxy_async();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            mt += 2;
            if (mt >= n_mt) { mt -= n_mt; ++kt; if (mt >= n_mt) { mt -= n_mt; ++kt; } }
        }
        const int cpart = warp >> 2;
        if (cpart == 0) {                                             // one warp per lane quarter finishes the dQ rows
            mbar_wait(dq_done, 0);
            tcgen05_fence_after();
            for (int mt = 0; mt < n_mt; ++mt) {
                const int i = mt * TC_M + r;
                uint32_t o[32];
                tmem_ld_32x32b_x32(tmem_base + OB_COL_DQ + mt * DH + lane_sel, o);
                tmem_ld_wait();
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
    if (warp == OB_WARP_MMA) {
        tcgen05_fence_after();
        tmem_dealloc<OB_TMEM_COLS>(tmem_base);
    }
}

bool tc_bwd_onepass_eligible(const AttnParams& p) {
    return p.bias_table != nullptr && p.mode == CTC_MODE_SPATIAL && p.n % OB_NK == 0 && p.W % 8 == 0 && p.n >= 64 &&
           p.n <= OB_MAX_MT * TC_M;
}

int run_tc_bwd_onepass(const AttnParams& p, cudaStream_t st) {
    const int n_rows = ((p.n + TC_M - 1) / TC_M) * TC_M;
    const ObSmem L = ob_layout(n_rows, (2 * p.H - 1) * (2 * p.W - 1));
    // up to 1023 bytes of slack for the 1024-byte alignment of the swizzled tiles; at n = 576 the layout leaves only 312
    // (dynamic shared memory starts 1024-aligned in practice; the kernel traps if the aligned layout does not fit)
    CTC_REQUIRE((size_t)L.total <= 227 * 1024, "attention one-pass backward: %d bytes of shared memory exceed 227 KB", L.total);
    const size_t smem = (size_t)L.total + 1024 <= 227 * 1024 ? (size_t)L.total + 1024 : (size_t)227 * 1024;
    static size_t configured_dev[kMaxDevices] = {};
    size_t& configured = configured_dev[current_device()];
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    attn_tc_bwd_onepass_kernel<<<dim3(p.n_seq, p.heads), OB_THREADS, smem, st>>>(p);
    CTC_LAUNCH_CHECK();
    return 0;
}

}  // namespace ctc
