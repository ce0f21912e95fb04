// 3-D patch embedding front end (reference: src/utils/ctvit.py:44-49):
//   Rearrange 'b c (t pt) (h p1) (w p2) -> b t h w (c pt p1 p2)'  +  LayerNorm(pt*p1*p2)
// fused with the two input perturbations of the attribution methods so that perturbed volumes are
// never materialised:
//   - integrated gradients:  x' = 1 + alpha * (x - 1)          (visualizations.py:853-862)
//   - occlusion:             x'[cube] = -1                      (visualizations.py:380-381)
// HBM-bound: the fp32 volume is read exactly once — a CTA owns a group of G patches that are adjacent
// along W and fetches its [pt][p1][G*p2] box with ONE 4-D TMA load (cp.async.bulk.tensor over the
// [B, D, H, W] volume; cp.async 16-byte copies only when the geometry does not fit a tensor map) —
// staged in shared memory, normalised with fp32 two-pass statistics (near-constant "air" patches
// have rstd up to ~316, so the statistics must not be taken in bf16) and written once as the
// bf16 A-operand of the patch-embedding GEMM.
#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {

struct PatchGeom {
    int B, D, H, W, pt, p, T, Hp, Wp, P, G;  // G = patches per CTA (along W)
    long long vol_stride;
};

CTC_DEVINL float perturb(float v, int d, int y, int x, float alpha, bool has_alpha, const int* oc, float oval) {
    if (has_alpha) v = 1.0f + alpha * (v - 1.0f);
    if (oc && oc[3] > 0 && d >= oc[0] && d < oc[0] + oc[3] && y >= oc[1] && y < oc[1] + oc[4] && x >= oc[2] &&
        x < oc[2] + oc[5])
        v = oval;
    return v;
}

// Shared-memory tile: [pt*p rows][G*p floats] (the CTA's G patches side by side along W).
// Every pass walks the tile with the same "quad" mapping: thread -> (column quad cq, row group), i.e. one
// 128-bit shared-memory access per 4 voxels of ONE patch (p % 4 == 0) and no per-element index arithmetic.
struct QuadMap {
    int ncq, rgq, cq, rgi, j, c;      // column quads per row, row groups, my quad, my row group, my patch, col in patch
    bool active;
    CTC_DEVINL void init(const PatchGeom& g, int rowlen) {
        ncq = rowlen >> 2;
        rgq = blockDim.x / ncq;
        cq = threadIdx.x % ncq; rgi = threadIdx.x / ncq;
        active = rgi < rgq;
        j = (cq * 4) / g.p;
        c = cq * 4 - j * g.p;
    }
};

// Deterministic per-patch reduction of two per-thread partials (fixed shuffle tree, no floating-point atomics:
// two runs give bit-identical statistics — the VQ arg-max downstream amplifies last-bit noise).
CTC_DEVINL void patch_reduce2(const PatchGeom& g, const QuadMap& m, float a0, float a1, float* s_part, float* out0,
                              float* out1) {
    // s_part layout: [2][rgq][ncq]
    const int n = m.rgq * m.ncq;
    if (m.active) { s_part[m.rgi * m.ncq + m.cq] = a0; s_part[n + m.rgi * m.ncq + m.cq] = a1; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qpp = g.p >> 2;                             // quads per patch row
    for (int j = warp; j < g.G; j += blockDim.x >> 5) {
        float t0 = 0.f, t1 = 0.f;
        for (int i = lane; i < m.rgq * qpp; i += 32) {
            const int idx = (i / qpp) * m.ncq + j * qpp + i % qpp;
            t0 += s_part[idx]; t1 += s_part[n + idx];
        }
        t0 = warp_sum(t0); t1 = warp_sum(t1);
        if (lane == 0) { out0[j] = t0; if (out1) out1[j] = t1; }
    }
    __syncthreads();
}

// per-patch mean / rstd with the two-pass formula
CTC_DEVINL void patch_stats(const float* tile, const PatchGeom& g, const QuadMap& m, int rows, int rowlen, float eps,
                            float* s_part, float* s_mean, float* s_rstd) {
    float s = 0.f;
    if (m.active)
        for (int r = m.rgi; r < rows; r += m.rgq) {
            const float4 v = *reinterpret_cast<const float4*>(tile + r * rowlen + m.cq * 4);
            s += (v.x + v.y) + (v.z + v.w);
        }
    patch_reduce2(g, m, s, 0.f, s_part, s_mean, nullptr);
    if (threadIdx.x < g.G) s_mean[threadIdx.x] /= g.P;
    __syncthreads();
    float q = 0.f;
    if (m.active) {
        const float mean = s_mean[m.j];
        for (int r = m.rgi; r < rows; r += m.rgq) {
            const float4 v = *reinterpret_cast<const float4*>(tile + r * rowlen + m.cq * 4);
            const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    patch_reduce2(g, m, q, 0.f, s_part, s_rstd, nullptr);
    if (threadIdx.x < g.G) s_rstd[threadIdx.x] = rsqrtf(s_rstd[threadIdx.x] / g.P + eps);
    __syncthreads();
}

// async tile load: every thread issues all of its 16-byte copies back to back (coalesced along W)
CTC_DEVINL void load_tile_async(float* tile, const float* vb, const PatchGeom& g, int tp, int hp, int x0, int rows,
                                int rowlen) {
    const int vec_per_row = rowlen >> 2;
    int r = threadIdx.x / vec_per_row, v = threadIdx.x % vec_per_row;
    const int dr = blockDim.x / vec_per_row, dv = blockDim.x % vec_per_row;
    while (r < rows) {
        const int pt_i = r / g.p;
        const int d = tp * g.pt + pt_i, y = hp * g.p + (r - pt_i * g.p);
        cp_async_16(tile + r * rowlen + v * 4, vb + ((long long)d * g.H + y) * g.W + x0 + v * 4);
        v += dv; r += dr;
        if (v >= vec_per_row) { v -= vec_per_row; ++r; }
    }
}

// one elected thread issues the TMA box load of the CTA's volume tile; everybody waits on the mbarrier
CTC_DEVINL void load_tile_tma(float* tile, const CUtensorMap* tmap, uint64_t* bar, const PatchGeom& g, int b, int tp,
                              int hp, int x0, int rows, int rowlen) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)(rows * rowlen * 4));
        tma_load_4d(tile, tmap, bar, x0, hp * g.p, tp * g.pt, g.vol_stride ? b : 0);
    }
    mbar_wait(bar, 0);
}

template <bool TMA>
__global__ void __launch_bounds__(256)
patchify_ln_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ vol, PatchGeom g,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       const float* __restrict__ alpha, const int* __restrict__ occl, float oval,
                       __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(128) float tile[];
    __shared__ float s_mean[32], s_rstd[32], s_part[512];
    __shared__ uint64_t tma_bar;
    const int groups_w = g.Wp / g.G;
    int bid = blockIdx.x;
    const int gw = bid % groups_w; bid /= groups_w;
    const int hp = bid % g.Hp; bid /= g.Hp;
    const int tp = bid % g.T;
    const int b = bid / g.T;
    const int rows = g.pt * g.p;          // (pt, p1) pairs
    const int rowlen = g.G * g.p;         // floats per row in this CTA
    const float* vb = vol + (long long)b * g.vol_stride;
    const bool has_alpha = alpha != nullptr;
    const float a = has_alpha ? alpha[b] : 1.f;
    const int* oc = occl ? occl + b * 6 : nullptr;
    const int x0 = gw * rowlen;
    if (TMA) {
        load_tile_tma(tile, &tmap, &tma_bar, g, b, tp, hp, x0, rows, rowlen);
    } else {
        load_tile_async(tile, vb, g, tp, hp, x0, rows, rowlen);
        cp_async_wait_all();
        __syncthreads();
    }
    // ---- perturbations applied in shared memory (only when requested / when the cube touches this tile)
    bool hit = false;
    if (oc && oc[3] > 0) {
        const int d0 = tp * g.pt, y0 = hp * g.p;
        hit = d0 < oc[0] + oc[3] && d0 + g.pt > oc[0] && y0 < oc[1] + oc[4] && y0 + g.p > oc[1] &&
              x0 < oc[2] + oc[5] && x0 + rowlen > oc[2];
    }
    if (has_alpha || hit) {
        int r = threadIdx.x / rowlen, c = threadIdx.x % rowlen;
        const int dr = blockDim.x / rowlen, dc = blockDim.x % rowlen;
        while (r < rows) {
            const int pt_i = r / g.p;
            const int d = tp * g.pt + pt_i, y = hp * g.p + (r - pt_i * g.p);
            tile[r * rowlen + c] = perturb(tile[r * rowlen + c], d, y, x0 + c, a, has_alpha, hit ? oc : nullptr, oval);
            c += dc; r += dr;
            if (c >= rowlen) { c -= rowlen; ++r; }
        }
        __syncthreads();
    }
    QuadMap m; m.init(g, rowlen);
    patch_stats(tile, g, m, rows, rowlen, eps, s_part, s_mean, s_rstd);
    // ---- normalise + affine, 4 voxels -> 4 bf16 (8 bytes) per step
    if (m.active) {
        const long long tok = (((long long)b * g.T + tp) * g.Hp + hp) * g.Wp + (long long)gw * g.G + m.j;
        const float mean = s_mean[m.j], rstd = s_rstd[m.j];
        __nv_bfloat16* orow = out + tok * g.P;
        for (int r = m.rgi; r < rows; r += m.rgq) {
            const int e = r * g.p + m.c;
            const float4 v = *reinterpret_cast<const float4*>(tile + r * rowlen + m.cq * 4);
            const float4 gm = *reinterpret_cast<const float4*>(gamma + e);
            const float4 bt = *reinterpret_cast<const float4*>(beta + e);
            *reinterpret_cast<uint2*>(orow + e) =
                make_uint2(pack_bf16((v.x - mean) * rstd * gm.x + bt.x, (v.y - mean) * rstd * gm.y + bt.y),
                           pack_bf16((v.z - mean) * rstd * gm.z + bt.z, (v.w - mean) * rstd * gm.w + bt.w));
        }
    }
}

// Backward: dx' = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dY * gamma;  chain through
// x' = 1 + alpha (x - 1) is NOT applied: IG differentiates w.r.t. the interpolated input itself
// (visualizations.py:863,872).  The occlusion path is forward only.
template <bool TMA>
__global__ void __launch_bounds__(256)
patchify_ln_bwd_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ vol, PatchGeom g,
                       const float* __restrict__ gamma, float eps, const float* __restrict__ alpha,
                       const __nv_bfloat16* __restrict__ dy, float* __restrict__ grad, int sum_over_batch,
                       float wscale) {
    extern __shared__ __align__(128) float tile[];
    __shared__ float s_mean[32], s_rstd[32], s_part[512], s_mg[32], s_mgx[32];
    __shared__ uint64_t tma_bar;
    const int groups_w = g.Wp / g.G;
    int bid = blockIdx.x;
    const int gw = bid % groups_w; bid /= groups_w;
    const int hp = bid % g.Hp; bid /= g.Hp;
    const int tp = bid % g.T;
    const int b = bid / g.T;
    const int rows = g.pt * g.p;
    const int rowlen = g.G * g.p;
    const float* vb = vol + (long long)b * g.vol_stride;
    const bool has_alpha = alpha != nullptr;
    const float a = has_alpha ? alpha[b] : 1.f;
    const int x0 = gw * rowlen;
    const long long tok0 = (((long long)b * g.T + tp) * g.Hp + hp) * g.Wp + (long long)gw * g.G;
    __nv_bfloat16* dys = reinterpret_cast<__nv_bfloat16*>(tile + rows * rowlen);   // [G][P] bf16 (contiguous in global)
    if (!TMA) load_tile_async(tile, vb, g, tp, hp, x0, rows, rowlen);
    {
        const int n16 = g.G * g.P / 8;                      // the CTA's G patches are adjacent rows of dY
        const __nv_bfloat16* src = dy + tok0 * g.P;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) cp_async_16(dys + i * 8, src + i * 8);
    }
    if (TMA) load_tile_tma(tile, &tmap, &tma_bar, g, b, tp, hp, x0, rows, rowlen);
    cp_async_wait_all();
    __syncthreads();
    if (has_alpha) {
        for (int i = threadIdx.x; i < rows * rowlen; i += blockDim.x) tile[i] = 1.f + a * (tile[i] - 1.f);
        __syncthreads();
    }
    QuadMap m; m.init(g, rowlen);
    patch_stats(tile, g, m, rows, rowlen, eps, s_part, s_mean, s_rstd);
    // ---- per-patch sum(g) and sum(g * xhat), g = dY * gamma (dY tile is in shared memory)
    float sg = 0.f, sgx = 0.f;
    if (m.active) {
        const float mean = s_mean[m.j], rstd = s_rstd[m.j];
        for (int r = m.rgi; r < rows; r += m.rgq) {
            const int e = r * g.p + m.c;
            const float4 v = *reinterpret_cast<const float4*>(tile + r * rowlen + m.cq * 4);
            const float4 gm = *reinterpret_cast<const float4*>(gamma + e);
            const uint2 dd = *reinterpret_cast<const uint2*>(dys + m.j * g.P + e);
            const float2 d01 = unpack_bf16(dd.x), d23 = unpack_bf16(dd.y);
            const float g0 = d01.x * gm.x, g1 = d01.y * gm.y, g2 = d23.x * gm.z, g3 = d23.y * gm.w;
            sg += (g0 + g1) + (g2 + g3);
            sgx += (g0 * (v.x - mean) + g1 * (v.y - mean) + g2 * (v.z - mean) + g3 * (v.w - mean)) * rstd;
        }
    }
    patch_reduce2(g, m, sg, sgx, s_part, s_mg, s_mgx);
    // ---- dx in place of x (same tile positions), then store coalesced along W
    if (m.active) {
        const float mean = s_mean[m.j], rstd = s_rstd[m.j];
        const float invP = 1.f / g.P;
        const float mg = s_mg[m.j] * invP, mgx = s_mgx[m.j] * invP;
        for (int r = m.rgi; r < rows; r += m.rgq) {
            const int e = r * g.p + m.c;
            float4* tp4 = reinterpret_cast<float4*>(tile + r * rowlen + m.cq * 4);
            const float4 v = *tp4;
            const float4 gm = *reinterpret_cast<const float4*>(gamma + e);
            const uint2 dd = *reinterpret_cast<const uint2*>(dys + m.j * g.P + e);
            const float2 d01 = unpack_bf16(dd.x), d23 = unpack_bf16(dd.y);
            float4 o;
            o.x = rstd * (d01.x * gm.x - mg - (v.x - mean) * rstd * mgx);
            o.y = rstd * (d01.y * gm.y - mg - (v.y - mean) * rstd * mgx);
            o.z = rstd * (d23.x * gm.z - mg - (v.z - mean) * rstd * mgx);
            o.w = rstd * (d23.y * gm.w - mg - (v.w - mean) * rstd * mgx);
            *tp4 = o;
        }
    }
    __syncthreads();
    float* gb = grad + (sum_over_batch ? 0 : (long long)b * g.D * g.H * g.W);
    {
        const int vec_per_row = rowlen >> 2;
        int r = threadIdx.x / vec_per_row, v = threadIdx.x % vec_per_row;
        const int dr = blockDim.x / vec_per_row, dv = blockDim.x % vec_per_row;
        while (r < rows) {
            const int v4 = v * 4;
            const int pt_i = r / g.p;
            const int d = tp * g.pt + pt_i, y = hp * g.p + (r - pt_i * g.p);
            const float4 val = *reinterpret_cast<const float4*>(tile + r * rowlen + v4);
            float* dst = gb + ((long long)d * g.H + y) * g.W + x0 + v4;
            if (sum_over_batch) {
                atomicAdd(dst + 0, val.x * wscale); atomicAdd(dst + 1, val.y * wscale);
                atomicAdd(dst + 2, val.z * wscale); atomicAdd(dst + 3, val.w * wscale);
            } else {
                *reinterpret_cast<float4*>(dst) = val;
            }
            v += dv; r += dr;
            if (v >= vec_per_row) { v -= vec_per_row; ++r; }
        }
    }
}

static int make_geom(PatchGeom& g, long long vol_stride, int B, int D, int H, int W, int pt, int p, int bytes_per_elem) {
    CTC_REQUIRE(D % pt == 0 && H % p == 0 && W % p == 0, "patchify: volume %dx%dx%d not divisible by patch %dx%dx%d",
                D, H, W, pt, p, p);
    CTC_REQUIRE(p % 4 == 0, "patchify: patch size %d must be a multiple of 4 (128-bit loads)", p);
    g.B = B; g.D = D; g.H = H; g.W = W; g.pt = pt; g.p = p;
    g.T = D / pt; g.Hp = H / p; g.Wp = W / p; g.P = pt * p * p; g.vol_stride = vol_stride;
    // patches per CTA: largest divisor of Wp with tile <= 50 KB (four resident CTAs per SM overlap their
    // load / statistics / store phases; two 96 KB CTAs left DRAM at 35 % in ncu) and <= 32 patches
    int G = 1;
    for (int c = 1; c <= g.Wp && c <= 32; ++c)
        if (g.Wp % c == 0 && (long long)c * g.P * bytes_per_elem <= 50 * 1024 && c * p <= 1024) G = c;
    g.G = G;
    CTC_REQUIRE((long long)G * g.P * bytes_per_elem <= 200 * 1024, "patchify: patch of %d voxels does not fit shared memory", g.P);
    CTC_REQUIRE(g.P % 8 == 0, "patchify: patch volume %d must be a multiple of 8", g.P);
    return 0;
}

typedef CUresult (*PFN_encodeTiledP)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 4-D tensor map over the fp32 volume(s) [Bv, D, H, W] with box [1, pt, p, G*p]; false when the geometry does not fit
static bool make_volume_tmap(CUtensorMap* map, const float* vol, const PatchGeom& g) {
    static PFN_encodeTiledP enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return false;
        enc = reinterpret_cast<PFN_encodeTiledP>(p);
    }
    const int rowlen = g.G * g.p;
    if (rowlen > 256 || g.p > 256 || g.pt > 256 || (rowlen * 4) % 16 != 0 || (g.W * 4) % 16 != 0 ||
        (reinterpret_cast<uintptr_t>(vol) & 15) != 0 || (g.vol_stride && (g.vol_stride * 4) % 16 != 0))
        return false;
    const long long vs = g.vol_stride ? g.vol_stride : (long long)g.D * g.H * g.W;
    cuuint64_t dims[4] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.D, (cuuint64_t)(g.vol_stride ? g.B : 1)};
    cuuint64_t strides[3] = {(cuuint64_t)g.W * 4, (cuuint64_t)g.H * g.W * 4, (cuuint64_t)vs * 4};
    cuuint32_t box[4] = {(cuuint32_t)rowlen, (cuuint32_t)g.p, (cuuint32_t)g.pt, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(vol), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ctc

using namespace ctc;

extern "C" int ctc_patchify_ln_fwd(const float* volume, int64_t vol_batch_stride, int B, int D, int H, int W, int pt,
                                   int p, const float* gamma, const float* beta, float eps, const float* alpha,
                                   const int* occl, float occl_value, void* out_bf16, void* stream) {
    PatchGeom g;
    if (int e = make_geom(g, vol_batch_stride, B, D, H, W, pt, p, 4)) return e;
    const size_t smem = (size_t)g.G * g.P * 4;
    static size_t configured_dev[kMaxDevices] = {};
    
    size_t& configured = configured_dev[current_device()];
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_fwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_fwd_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    const long long grid = (long long)B * g.T * g.Hp * (g.Wp / g.G);
    CUtensorMap tmap{};
    if (make_volume_tmap(&tmap, volume, g))
        patchify_ln_fwd_kernel<true><<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(
            tmap, volume, g, gamma, beta, eps, alpha, occl, occl_value, (__nv_bfloat16*)out_bf16);
    else
        patchify_ln_fwd_kernel<false><<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(
            tmap, volume, g, gamma, beta, eps, alpha, occl, occl_value, (__nv_bfloat16*)out_bf16);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_patchify_ln_bwd(const float* volume, int64_t vol_batch_stride, int B, int D, int H, int W, int pt,
                                   int p, const float* gamma, float eps, const float* alpha, const void* dy_bf16,
                                   float* grad, int sum_over_batch, float wscale, void* stream) {
    PatchGeom g;
    if (int e = make_geom(g, vol_batch_stride, B, D, H, W, pt, p, 6)) return e;
    const size_t smem = (size_t)g.G * g.P * 6;
    static size_t configured_dev[kMaxDevices] = {};
    
    size_t& configured = configured_dev[current_device()];
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_bwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_bwd_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            (int)cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    const long long grid = (long long)B * g.T * g.Hp * (g.Wp / g.G);
    CUtensorMap tmap{};
    if (make_volume_tmap(&tmap, volume, g))
        patchify_ln_bwd_kernel<true><<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(
            tmap, volume, g, gamma, eps, alpha, (const __nv_bfloat16*)dy_bf16, grad, sum_over_batch, wscale);
    else
        patchify_ln_bwd_kernel<false><<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(
            tmap, volume, g, gamma, eps, alpha, (const __nv_bfloat16*)dy_bf16, grad, sum_over_batch, wscale);
    CTC_LAUNCH_CHECK();
    return 0;
}
