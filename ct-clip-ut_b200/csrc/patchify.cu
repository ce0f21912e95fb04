// 3-D patch embedding front end (reference: src/utils/ctvit.py:44-49):
//   Rearrange 'b c (t pt) (h p1) (w p2) -> b t h w (c pt p1 p2)'  +  LayerNorm(pt*p1*p2)
// fused with the two input perturbations of the attribution methods so that perturbed volumes are
// never materialised:
//   - integrated gradients:  x' = 1 + alpha * (x - 1)          (visualizations.py:853-862)
//   - occlusion:             x'[cube] = -1                      (visualizations.py:380-381)
// HBM-bound: the fp32 volume is read exactly once with 128-bit coalesced loads (a CTA owns a
// group of G patches that are adjacent along W, i.e. pt*p1 contiguous runs of G*p2 floats),
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

// smem tile: [pt*p rows][G*p floats]; stats per patch
__global__ void __launch_bounds__(256)
patchify_ln_fwd_kernel(const float* __restrict__ vol, PatchGeom g, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, const float* __restrict__ alpha,
                       const int* __restrict__ occl, float oval, __nv_bfloat16* __restrict__ out) {
    extern __shared__ float tile[];
    __shared__ float s_mean[32], s_rstd[32];
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
    // ---- load (coalesced along W), apply perturbations
    const int vec_per_row = rowlen >> 2;
    for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
        const int r = i / vec_per_row, v4 = (i % vec_per_row) * 4;
        const int d = tp * g.pt + r / g.p, y = hp * g.p + r % g.p;
        float4 v = *reinterpret_cast<const float4*>(vb + ((long long)d * g.H + y) * g.W + x0 + v4);
        v.x = perturb(v.x, d, y, x0 + v4 + 0, a, has_alpha, oc, oval);
        v.y = perturb(v.y, d, y, x0 + v4 + 1, a, has_alpha, oc, oval);
        v.z = perturb(v.z, d, y, x0 + v4 + 2, a, has_alpha, oc, oval);
        v.w = perturb(v.w, d, y, x0 + v4 + 3, a, has_alpha, oc, oval);
        *reinterpret_cast<float4*>(tile + r * rowlen + v4) = v;
    }
    __syncthreads();
    // ---- per-patch statistics: warps loop over patches
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < g.G; j += nwarps) {
        float s = 0.f;
        for (int e = lane; e < g.P; e += 32) s += tile[(e / g.p) * rowlen + j * g.p + e % g.p];
        const float mean = warp_sum(s) / g.P;
        float q = 0.f;
        for (int e = lane; e < g.P; e += 32) {
            const float dlt = tile[(e / g.p) * rowlen + j * g.p + e % g.p] - mean;
            q += dlt * dlt;
        }
        const float rstd = rsqrtf(warp_sum(q) / g.P + eps);
        if (lane == 0) { s_mean[j] = mean; s_rstd[j] = rstd; }
    }
    __syncthreads();
    // ---- normalise + affine, write bf16 rows (2 elements per thread, contiguous along the patch row)
    const long long tok0 = (((long long)b * g.T + tp) * g.Hp + hp) * g.Wp + (long long)gw * g.G;
    const int half = g.P >> 1;
    for (int i = threadIdx.x; i < g.G * half; i += blockDim.x) {
        const int j = i / half, e = (i % half) * 2;
        const float mean = s_mean[j], rstd = s_rstd[j];
        const int r = e / g.p, c = e % g.p;  // p is even, so e and e+1 share a tile row
        const float v0 = tile[r * rowlen + j * g.p + c], v1 = tile[r * rowlen + j * g.p + c + 1];
        const float o0 = (v0 - mean) * rstd * gamma[e] + beta[e];
        const float o1 = (v1 - mean) * rstd * gamma[e + 1] + beta[e + 1];
        *reinterpret_cast<uint32_t*>(out + (tok0 + j) * g.P + e) = pack_bf16(o0, o1);
    }
}

// Backward: dx' = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dY * gamma;  chain through
// x' = 1 + alpha (x - 1) is NOT applied: IG differentiates w.r.t. the interpolated input itself
// (visualizations.py:863,872).  Occluded voxels receive whatever gradient flows to x' (the
// occlusion path is forward only, so the two are never combined).
__global__ void __launch_bounds__(256)
patchify_ln_bwd_kernel(const float* __restrict__ vol, PatchGeom g, const float* __restrict__ gamma, float eps,
                       const float* __restrict__ alpha, const __nv_bfloat16* __restrict__ dy,
                       float* __restrict__ grad, int sum_over_batch, float wscale) {
    extern __shared__ float tile[];
    __shared__ float s_mean[32], s_rstd[32], s_mg[32], s_mgx[32];
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
    const int vec_per_row = rowlen >> 2;
    for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
        const int r = i / vec_per_row, v4 = (i % vec_per_row) * 4;
        const int d = tp * g.pt + r / g.p, y = hp * g.p + r % g.p;
        float4 v = *reinterpret_cast<const float4*>(vb + ((long long)d * g.H + y) * g.W + x0 + v4);
        if (has_alpha) {
            v.x = 1.f + a * (v.x - 1.f); v.y = 1.f + a * (v.y - 1.f);
            v.z = 1.f + a * (v.z - 1.f); v.w = 1.f + a * (v.w - 1.f);
        }
        *reinterpret_cast<float4*>(tile + r * rowlen + v4) = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const long long tok0 = (((long long)b * g.T + tp) * g.Hp + hp) * g.Wp + (long long)gw * g.G;
    for (int j = warp; j < g.G; j += nwarps) {
        float s = 0.f;
        for (int e = lane; e < g.P; e += 32) s += tile[(e / g.p) * rowlen + j * g.p + e % g.p];
        const float mean = warp_sum(s) / g.P;
        float q = 0.f;
        for (int e = lane; e < g.P; e += 32) {
            const float dlt = tile[(e / g.p) * rowlen + j * g.p + e % g.p] - mean;
            q += dlt * dlt;
        }
        const float rstd = rsqrtf(warp_sum(q) / g.P + eps);
        const __nv_bfloat16* dyr = dy + (tok0 + j) * g.P;
        float sg = 0.f, sgx = 0.f;
        for (int e = lane; e < g.P; e += 32) {
            const float gg = __bfloat162float(dyr[e]) * gamma[e];
            const float xh = (tile[(e / g.p) * rowlen + j * g.p + e % g.p] - mean) * rstd;
            sg += gg; sgx += gg * xh;
        }
        sg = warp_sum(sg) / g.P; sgx = warp_sum(sgx) / g.P;
        if (lane == 0) { s_mean[j] = mean; s_rstd[j] = rstd; s_mg[j] = sg; s_mgx[j] = sgx; }
    }
    __syncthreads();
    // overwrite the tile with dx (same element order), then store coalesced along W
    for (int i = threadIdx.x; i < g.G * g.P; i += blockDim.x) {
        const int j = i / g.P, e = i % g.P;
        const int r = e / g.p, c = e % g.p;
        const float mean = s_mean[j], rstd = s_rstd[j];
        const float xh = (tile[r * rowlen + j * g.p + c] - mean) * rstd;
        const float gg = __bfloat162float(dy[(tok0 + j) * g.P + e]) * gamma[e];
        tile[r * rowlen + j * g.p + c] = rstd * (gg - s_mg[j] - xh * s_mgx[j]);
    }
    __syncthreads();
    float* gb = grad + (sum_over_batch ? 0 : (long long)b * g.D * g.H * g.W);
    for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
        const int r = i / vec_per_row, v4 = (i % vec_per_row) * 4;
        const int d = tp * g.pt + r / g.p, y = hp * g.p + r % g.p;
        float4 v = *reinterpret_cast<const float4*>(tile + r * rowlen + v4);
        float* dst = gb + ((long long)d * g.H + y) * g.W + x0 + v4;
        if (sum_over_batch) {
            atomicAdd(dst + 0, v.x * wscale); atomicAdd(dst + 1, v.y * wscale);
            atomicAdd(dst + 2, v.z * wscale); atomicAdd(dst + 3, v.w * wscale);
        } else {
            *reinterpret_cast<float4*>(dst) = v;
        }
    }
}

static int make_geom(PatchGeom& g, long long vol_stride, int B, int D, int H, int W, int pt, int p) {
    CTC_REQUIRE(D % pt == 0 && H % p == 0 && W % p == 0, "patchify: volume %dx%dx%d not divisible by patch %dx%dx%d",
                D, H, W, pt, p, p);
    CTC_REQUIRE(p % 4 == 0, "patchify: patch size %d must be a multiple of 4 (128-bit loads)", p);
    g.B = B; g.D = D; g.H = H; g.W = W; g.pt = pt; g.p = p;
    g.T = D / pt; g.Hp = H / p; g.Wp = W / p; g.P = pt * p * p; g.vol_stride = vol_stride;
    // patches per CTA: largest divisor of Wp with tile <= 96 KB and <= 32 patches
    int G = 1;
    for (int c = 1; c <= g.Wp && c <= 32; ++c)
        if (g.Wp % c == 0 && (long long)c * g.P * 4 <= 96 * 1024) G = c;
    g.G = G;
    CTC_REQUIRE((long long)G * g.P * 4 <= 200 * 1024, "patchify: patch of %d voxels does not fit shared memory", g.P);
    return 0;
}

}  // namespace ctc

using namespace ctc;

extern "C" int ctc_patchify_ln_fwd(const float* volume, int64_t vol_batch_stride, int B, int D, int H, int W, int pt,
                                   int p, const float* gamma, const float* beta, float eps, const float* alpha,
                                   const int* occl, float occl_value, void* out_bf16, void* stream) {
    PatchGeom g;
    if (int e = make_geom(g, vol_batch_stride, B, D, H, W, pt, p)) return e;
    const size_t smem = (size_t)g.G * g.P * 4;
    static size_t configured = 0;
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const long long grid = (long long)B * g.T * g.Hp * (g.Wp / g.G);
    patchify_ln_fwd_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(
        volume, g, gamma, beta, eps, alpha, occl, occl_value, (__nv_bfloat16*)out_bf16);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_patchify_ln_bwd(const float* volume, int64_t vol_batch_stride, int B, int D, int H, int W, int pt,
                                   int p, const float* gamma, float eps, const float* alpha, const void* dy_bf16,
                                   float* grad, int sum_over_batch, float wscale, void* stream) {
    PatchGeom g;
    if (int e = make_geom(g, vol_batch_stride, B, D, H, W, pt, p)) return e;
    const size_t smem = (size_t)g.G * g.P * 4;
    static size_t configured = 0;
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(patchify_ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const long long grid = (long long)B * g.T * g.Hp * (g.Wp / g.G);
    patchify_ln_bwd_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(
        volume, g, gamma, eps, alpha, (const __nv_bfloat16*)dy_bf16, grad, sum_over_batch, wscale);
    CTC_LAUNCH_CHECK();
    return 0;
}
