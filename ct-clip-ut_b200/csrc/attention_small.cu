// Warp-autonomous attention kernels for small sequences (temporal attention of the factorised CTViT transformer,
// src/utils/ctvit.py:99-101, n = T = 24 tokens).  See attention.cu for the numerics shared by all attention kernels.
#include "attention_common.cuh"

namespace ctc {

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

bool small_warp_path(const AttnParams& p) { return p.n <= SMALL_N && p.bias_table == nullptr; }
int run_small_fwd(const AttnParams& p, cudaStream_t st) {
    const size_t smem = 256 + (size_t)SMALL_FWD_WARPS * 2 * 3 * SMALL_TILE;
    const int n_tasks = p.n_seq * p.heads;
    int grid = (n_tasks + SMALL_FWD_WARPS - 1) / SMALL_FWD_WARPS;
    if (grid > 2 * num_sms()) grid = 2 * num_sms();
    return launch_attn<attn_small_fwd_kernel>(p, dim3(grid), SMALL_FWD_WARPS * 32, smem, st);
}
int run_small_bwd(const AttnParams& p, cudaStream_t st) {
    const size_t smem = 256 + (size_t)SMALL_BWD_WARPS * 2 * 4 * SMALL_TILE;
    const int n_tasks = p.n_seq * p.heads;
    int grid = (n_tasks + SMALL_BWD_WARPS - 1) / SMALL_BWD_WARPS;
    if (grid > num_sms()) grid = num_sms();
    return launch_attn<attn_small_bwd_kernel>(p, dim3(grid), SMALL_BWD_WARPS * 32, smem, st);
}


}  // namespace ctc
