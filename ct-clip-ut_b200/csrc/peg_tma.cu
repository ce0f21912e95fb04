// PEG depthwise 3x3x3 stencil (src/utils/attention.py:55-83), bulk-asynchronous version for sm_100a.
//
// The stencil is HBM-bound on paper (28 MB in / 28 MB out per volume) but a per-thread gather of its
// 27 taps is latency-bound: every output needs 9 new global rows and a warp can only keep a few of
// them in flight.  Here the copy engine does the gathering: a persistent CTA walks a list of output
// tiles (TT frames x TH rows x all W columns x 32 channels), and ONE 5-D TMA load per tile
// (cp.async.bulk.tensor.5d, box [32 c, W, TH+2, TT+2, 1]) brings the whole halo tile into shared memory —
// the causal / same-padding zeros of the convolution come for free from TMA's out-of-bounds fill (negative
// and past-the-end coordinates).  Two smem stages: the load of tile i+1 overlaps the arithmetic of tile i.
// A warp owns one (t, h) row of the tile, a lane one channel: the 9 new values per output column are
// conflict-free LDS.32 (32 consecutive words), the 27 weights stay in registers for the CTA's lifetime
// (a CTA keeps one channel chunk), stores are 128-byte coalesced.
//
// mode / transpose select the tap geometry (see elementwise.cu: SPATIAL (dt,dh,dw) = (a-2,b-1,c-1); TEMPORAL is
// the reference's axis scramble (dt,dh,dw) = (c-1,a-2,b-1); the adjoint gathers with negated offsets).
#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {

static constexpr int PT_TT = 4, PT_TH = 4, PT_CH = 32, PT_WMAX = 24;
static constexpr int PT_THREADS = PT_TT * PT_TH * 32;

struct PegTmaArgs {
    int B, T, H, W, C;
    const float* w27;
    const float* bias;
    int mode, sign;
    float* y;
    __nv_bfloat16* yb;
    int tiles_t, tiles_h, tiles_c, groups;   // groups = gridDim.x / tiles_c: CTAs sharing a channel chunk
};

CTC_DEVINL void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2,
                            int32_t c3, int32_t c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

template <int CW, bool BF16OUT>
__global__ void __launch_bounds__(PT_THREADS, 1)
peg_tma_kernel(const __grid_constant__ CUtensorMap tmap, const PegTmaArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // pointer arithmetic on the __shared__ array itself (not a uintptr_t round trip), so that the tile reads
    // below compile to LDS rather than generic loads
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int W = CW ? CW : a.W;
    const int row_elems = W * PT_CH;                                 // one (t, h) row of the tile
    const int tile_elems = row_elems * (PT_TH + 2) * (PT_TT + 2);
    const uint32_t tile_bytes = (uint32_t)tile_elems * 4u;
    const uint32_t stage_stride = (tile_bytes + 127u) & ~127u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * stage_stride);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tl = warp / PT_TH, hl = warp % PT_TH;
    const int cc = blockIdx.x % a.tiles_c, grp = blockIdx.x / a.tiles_c;
    const int c = cc * PT_CH + lane;
    const int n_sp = a.B * a.tiles_t * a.tiles_h;                    // spatial tiles per channel chunk

    // box origin relative to the tile origin (t0, h0): the taps reach [lo, lo + 2] on each axis
    int t_lo, h_lo;
    if (a.mode == CTC_MODE_SPATIAL) { t_lo = a.sign > 0 ? -2 : 0; h_lo = -1; }
    else                            { t_lo = -1; h_lo = a.sign > 0 ? -2 : 0; }

    auto issue = [&](int sp, int s) {
        const int th = sp % a.tiles_h, tt = (sp / a.tiles_h) % a.tiles_t, b = sp / (a.tiles_h * a.tiles_t);
        mbar_arrive_expect_tx(&full[s], tile_bytes);
        tma_load_5d(smem + s * stage_stride, &tmap, &full[s], cc * PT_CH, 0, th * PT_TH + h_lo, tt * PT_TT + t_lo, b);
    };
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (grp < n_sp) issue(grp, 0);
        if (grp + a.groups < n_sp) issue(grp + a.groups, 1);
    }

    // per-thread tap table: weights wk[k9][cw] and the smem row of tap row k9 = (p, q)
    float wk[9][3];
    int roff[9];
#pragma unroll
    for (int k9 = 0; k9 < 9; ++k9) {
        const int p = k9 / 3, q = k9 % 3;
        int dt, dh;
        if (a.mode == CTC_MODE_SPATIAL) { dt = p - 2; dh = q - 1; } else { dh = p - 2; dt = q - 1; }
        const int fr = tl + a.sign * dt - t_lo, rw = hl + a.sign * dh - h_lo;     // both in [0, extent)
        roff[k9] = (fr * (PT_TH + 2) + rw) * row_elems + lane;
#pragma unroll
        for (int cw = 0; cw < 3; ++cw) {
            const int cwm = (a.sign > 0) ? cw : 2 - cw;      // the adjoint reads x[w - (cw - 1)]: mirrored w taps
            const int tap = (a.mode == CTC_MODE_SPATIAL) ? (p * 3 + q) * 3 + cwm : (p * 3 + cwm) * 3 + q;
            wk[k9][cw] = (c < a.C) ? a.w27[tap * a.C + c] : 0.f;
        }
    }
    const float bv = (a.bias && a.sign > 0 && c < a.C) ? a.bias[c] : 0.f;

    int it = 0;
    for (int sp = grp; sp < n_sp; sp += a.groups, ++it) {
        const int s = it & 1;
        mbar_wait(&full[s], (it >> 1) & 1);
        const int th = sp % a.tiles_h, tt = (sp / a.tiles_h) % a.tiles_t, b = sp / (a.tiles_h * a.tiles_t);
        const int t = tt * PT_TT + tl, h = th * PT_TH + hl;
        if (t < a.T && h < a.H && c < a.C) {
            const float* xs = reinterpret_cast<const float*>(smem + s * stage_stride);
            const long long orow = ((((long long)b * a.T + t) * a.H + h) * W) * a.C + c;
            float* yo = a.y + orow;
            __nv_bfloat16* yb = BF16OUT ? a.yb + orow : nullptr;
            float A[9], Bc[9], D[9];
#pragma unroll
            for (int k9 = 0; k9 < 9; ++k9) { A[k9] = 0.f; Bc[k9] = xs[roff[k9]]; }
#pragma unroll
            for (int w = 0; w < (CW ? CW : PT_WMAX); ++w) {
                if (w < W) {
                    if (w + 1 < W) {
#pragma unroll
                        for (int k9 = 0; k9 < 9; ++k9) D[k9] = xs[roff[k9] + (w + 1) * PT_CH];
                    } else {
#pragma unroll
                        for (int k9 = 0; k9 < 9; ++k9) D[k9] = 0.f;
                    }
                    // k9 = 7 is the (dt, dh) = (0, 0) row: Bc[7] is the residual term x[w]
                    float acc0 = Bc[7] + bv, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
                    for (int k9 = 0; k9 < 9; ++k9) {
                        acc0 = fmaf(wk[k9][0], A[k9], acc0);
                        acc1 = fmaf(wk[k9][1], Bc[k9], acc1);
                        acc2 = fmaf(wk[k9][2], D[k9], acc2);
                    }
                    const float r = acc0 + (acc1 + acc2);
                    yo[(long long)w * a.C] = r;
                    if (BF16OUT) yb[(long long)w * a.C] = __float2bfloat16(r);
#pragma unroll
                    for (int k9 = 0; k9 < 9; ++k9) { A[k9] = Bc[k9]; Bc[k9] = D[k9]; }
                }
            }
        }
        __syncthreads();                                   // every warp is done reading stage s
        if (threadIdx.x == 0 && sp + 2 * a.groups < n_sp) issue(sp + 2 * a.groups, s);
    }
}

typedef CUresult (*PFN_encodeTiled5)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// returns 0 launched, -1 not eligible (caller falls back to the register-window kernel), > 0 error
int peg_tma_launch(const float* x, int B, int T, int H, int W, int C, const float* w27, const float* bias, int mode,
                   int transpose, float* y, void* y_bf16, cudaStream_t st) {
    if (C % PT_CH != 0 || W > PT_WMAX || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return -1;
    static PFN_encodeTiled5 enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return -1;
        enc = reinterpret_cast<PFN_encodeTiled5>(p);
    }
    CUtensorMap map;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4,
                             (cuuint64_t)T * H * W * C * 4};
    cuuint32_t box[5] = {(cuuint32_t)PT_CH, (cuuint32_t)W, (cuuint32_t)(PT_TH + 2), (cuuint32_t)(PT_TT + 2), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTC_REQUIRE(r == CUDA_SUCCESS, "peg: cuTensorMapEncodeTiled (5-D fp32) failed with CUresult %d", (int)r);

    PegTmaArgs a{};
    a.B = B; a.T = T; a.H = H; a.W = W; a.C = C; a.w27 = w27; a.bias = bias; a.mode = mode;
    a.sign = transpose ? -1 : 1; a.y = y; a.yb = (__nv_bfloat16*)y_bf16;
    a.tiles_t = (T + PT_TT - 1) / PT_TT; a.tiles_h = (H + PT_TH - 1) / PT_TH; a.tiles_c = C / PT_CH;
    const int n_sp = B * a.tiles_t * a.tiles_h;
    int groups = num_sms() / a.tiles_c;
    if (groups < 1) groups = 1;
    if (groups > n_sp) groups = n_sp;
    a.groups = groups;
    const size_t tile_bytes = ((size_t)W * PT_CH * (PT_TH + 2) * (PT_TT + 2) * 4 + 127) & ~(size_t)127;
    const size_t smem = 2 * tile_bytes + 64 + 128;
    const bool cw24 = (W == 24);
#define PT_LAUNCH(CWV, BF)                                                                                        \
    do {                                                                                                          \
        static size_t configured_dev[kMaxDevices] = {};                                                                             \
    size_t& configured = configured_dev[current_device()];                                                                             \
        if (smem > configured) {                                                                                  \
            CTC_CHECK_CUDA(cudaFuncSetAttribute(peg_tma_kernel<CWV, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)smem));                                                      \
            configured = smem;                                                                                    \
        }                                                                                                         \
        peg_tma_kernel<CWV, BF><<<groups * a.tiles_c, PT_THREADS, smem, st>>>(map, a);                            \
    } while (0)
    if (cw24) { if (y_bf16) PT_LAUNCH(24, true); else PT_LAUNCH(24, false); }
    else      { if (y_bf16) PT_LAUNCH(0, true); else PT_LAUNCH(0, false); }
#undef PT_LAUNCH
    CTC_LAUNCH_CHECK();
    return 0;
}

}  // namespace ctc
