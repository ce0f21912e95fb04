// Tail of the CTCLIP image tower: VQ nearest-code search + straight-through adjoint
// (src/utils/ctvit.py:115-118; vector_quantize_pytorch cosine-similarity codebook), temporal mean +
// latent projection + cosine similarity with the text latent (src/models/ctclip.py:110-127).
#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {

// ---------------------------------------------------------------------------------------------
// VQ refine: the bf16 tensor-core GEMM (epilogue ARGMAX) leaves per row, for every 128-code slice, the two best
// GROUPS of four consecutive codes (group maximum + first code of the group).  Every code of a candidate group
// whose bf16 maximum is within a rounding margin of the best is re-scored exactly in fp32 (x_fp32 . E_fp32), so
// the selected code equals the fp32 arg-max of the reference (ties -> lowest index, as torch.argmax).
// One warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_refine_kernel(const float* __restrict__ x, int R, int C, const float* __restrict__ codebook, int K,
                 const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int n_cand,
                 int* __restrict__ ind) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const float* xr = x + (long long)row * C;
    float ss = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    const float xnorm = sqrtf(warp_sum(ss));
    // bf16 operands: relative rounding 2^-9 per factor, so |score error| <= 2^-8 |x| (|e| = 1) per code;
    // two codes can move against each other, hence 2 * 2^-8 |x|.
    const float margin = xnorm * (1.0f / 128.0f);
    float best = -3.0e38f;
    for (int c = lane; c < n_cand; c += 32) best = fmaxf(best, cand_val[(long long)row * n_cand + c]);
    best = warp_max(best);
    float best_exact = -3.0e38f;
    int best_idx = 0x7fffffff;
    for (int c0 = 0; c0 < n_cand; c0 += 32) {
        const int c = c0 + lane;
        const bool mine = (c < n_cand) && (cand_val[(long long)row * n_cand + c] >= best - margin);
        unsigned mask = __ballot_sync(0xffffffffu, mine);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const int code0 = __shfl_sync(0xffffffffu, (c < n_cand) ? cand_idx[(long long)row * n_cand + c] : 0, src);
            for (int code = code0; code < min(code0 + 4, K); ++code) {
                const float* e = codebook + (long long)code * C;
                float d = 0.f;
                for (int k = lane * 4; k < C; k += 128) {
                    const float4 a = *reinterpret_cast<const float4*>(xr + k);
                    const float4 b = *reinterpret_cast<const float4*>(e + k);
                    d += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
                }
                d = warp_sum(d);
                if (d > best_exact || (d == best_exact && code < best_idx)) { best_exact = d; best_idx = code; }
            }
        }
    }
    if (lane == 0) ind[row] = best_idx;
}

// pooled[b, hw, :] = mean_t E[ind[b,t,hw]]  (+ optional gathered tokens)
__global__ void __launch_bounds__(128)
vq_gather_pool_kernel(const int* __restrict__ ind, const float* __restrict__ codebook, int B, int T, int HW, int C,
                      float* __restrict__ pooled, __nv_bfloat16* __restrict__ pooled_bf16, float* __restrict__ tokens) {
    const int bhw = blockIdx.x;
    const int b = bhw / HW, hw = bhw % HW;
    for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < T; ++t) {
            const long long r = ((long long)b * T + t) * HW + hw;
            const float4 e = *reinterpret_cast<const float4*>(codebook + (long long)ind[r] * C + c);
            acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
            if (tokens) *reinterpret_cast<float4*>(tokens + r * C + c) = e;
        }
        const float inv = 1.f / T;
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        const long long o = (long long)bhw * C + c;
        *reinterpret_cast<float4*>(pooled + o) = acc;
        if (pooled_bf16)
            *reinterpret_cast<uint2*>(pooled_bf16 + o) = make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
    }
}

// straight-through adjoint to the pre-VQ activations (one warp per token row)
__global__ void __launch_bounds__(256)
vq_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ dtokens, const float* __restrict__ x,
              int B, int T, int HW, int C, int grad_mode, float* __restrict__ dx) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * T * HW) return;
    const int hw = (int)(row % HW);
    const int b = (int)(row / ((long long)T * HW));
    const float* gp = dpooled + ((long long)b * HW + hw) * C;
    const float* xr = x + row * C;
    const float invT = 1.f / T;
    float ss = 0.f, dot = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        float4 g = *reinterpret_cast<const float4*>(gp + c);
        g.x *= invT; g.y *= invT; g.z *= invT; g.w *= invT;
        if (dtokens) {
            const float4 d = *reinterpret_cast<const float4*>(dtokens + row * C + c);
            g.x += d.x; g.y += d.y; g.z += d.z; g.w += d.w;
        }
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        dot += v.x * g.x + v.y * g.y + v.z * g.z + v.w * g.w;
    }
    ss = warp_sum(ss); dot = warp_sum(dot);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    const float proj = dot * inv * inv;   // (xh . g) / |x| * (1/|x|) applied to x
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        float4 g = *reinterpret_cast<const float4*>(gp + c);
        g.x *= invT; g.y *= invT; g.z *= invT; g.w *= invT;
        if (dtokens) {
            const float4 d = *reinterpret_cast<const float4*>(dtokens + row * C + c);
            g.x += d.x; g.y += d.y; g.z += d.z; g.w += d.w;
        }
        float4 o;
        if (grad_mode == 0) {
            o.x = (g.x - v.x * proj) * inv; o.y = (g.y - v.y * proj) * inv;
            o.z = (g.z - v.z * proj) * inv; o.w = (g.w - v.w * proj) * inv;
        } else {
            o = g;
        }
        *reinterpret_cast<float4*>(dx + row * C + c) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// latent projection: latent[b, n] = sum_l pooled[b, l] * Wv[n, l].  L = 294 912, NL = 512: a
// 302 MB bf16 weight read once -> HBM-bound.  Each CTA owns a chunk of CH columns for all NL
// outputs and up to 8 batch rows (pooled chunk staged in smem); partial sums go to
// partial[chunk, b, n] and are reduced by a second kernel (deterministic, no atomics).
// ---------------------------------------------------------------------------------------------
static constexpr int LAT_CH = 1024;
static constexpr int LAT_BMAX = 8;

__global__ void __launch_bounds__(256)
latent_proj_kernel(const float* __restrict__ pooled, const __nv_bfloat16* __restrict__ wv,
                   const __nv_bfloat16* __restrict__ wv_lo, int B, long long L, int NL, float* __restrict__ partial) {
    __shared__ float sp[LAT_BMAX][LAT_CH];
    const int chunk = blockIdx.x;
    const int b0 = blockIdx.y * LAT_BMAX;
    const int nb = min(LAT_BMAX, B - b0);
    const long long l0 = (long long)chunk * LAT_CH;
    const int len = (int)min((long long)LAT_CH, L - l0);
    for (int i = threadIdx.x; i < LAT_BMAX * LAT_CH; i += blockDim.x) {
        const int b = i / LAT_CH, l = i % LAT_CH;
        sp[b][l] = (b < nb && l < len) ? pooled[(long long)(b0 + b) * L + l0 + l] : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < NL; n += blockDim.x >> 5) {
        const __nv_bfloat16* wr = wv + (long long)n * L + l0;
        float acc[LAT_BMAX];
#pragma unroll
        for (int b = 0; b < LAT_BMAX; ++b) acc[b] = 0.f;
        for (int l = lane * 8; l < len; l += 256) {
            const uint4 w4 = *reinterpret_cast<const uint4*>(wr + l);
            const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
            float wf[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 t = unpack_bf16(ww[e]); wf[2 * e] = t.x; wf[2 * e + 1] = t.y; }
            if (wv_lo) {                                              // W = hi + lo (both bf16): ~16 mantissa bits
                const uint4 l4 = *reinterpret_cast<const uint4*>(wv_lo + (long long)n * L + l0 + l);
                const uint32_t lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 t = unpack_bf16(lw[e]); wf[2 * e] += t.x; wf[2 * e + 1] += t.y; }
            }
#pragma unroll
            for (int b = 0; b < LAT_BMAX; ++b) {
                const float4 p0 = *reinterpret_cast<const float4*>(&sp[b][l]);
                const float4 p1 = *reinterpret_cast<const float4*>(&sp[b][l + 4]);
                acc[b] += wf[0] * p0.x + wf[1] * p0.y + wf[2] * p0.z + wf[3] * p0.w + wf[4] * p1.x + wf[5] * p1.y +
                          wf[6] * p1.z + wf[7] * p1.w;
            }
        }
#pragma unroll
        for (int b = 0; b < LAT_BMAX; ++b) {
            const float v = warp_sum(acc[b]);
            if (lane == 0 && b < nb) partial[((long long)chunk * B + b0 + b) * NL + n] = v;
        }
    }
}

// Tensor-core version (NL % 64 == 0, L % 64 == 0): the SIMT kernel above stages the pooled chunk in shared memory
// and is bound by its LDS traffic (1.23 ms for 32 rows, 4 weight passes).  Here each CTA streams its [NL x 1024]
// weight chunk from HBM ONCE for up to 32 batch rows with mma.sync m16n8k16: the reduction index inside a 32-wide
// k-block is permuted (thread t owns k = 8t .. 8t+7 of both operands), so that the weight fragment is ONE 16-byte
// global load per lane and the pooled fragment two float4 loads - no shared memory, no ldmatrix.  The fp32 pooled
// row is split into bf16 hi + lo parts (two MMAs), which keeps ~16 mantissa bits of the activations.  The WEIGHT is
// split the same way when wv_lo is given (W = hi + lo, lo = bf16(W - hi); a third MMA, hi-activation x lo-weight):
// this 294 912-long dot product decides a logit of magnitude 1e-2 whose DIFFERENCES between occluded volumes are the
// attribution signal, and the 2^-9 rounding of a bf16-only weight showed up as a 4e-4 logit error with every VQ code
// identical to the reference (profiles/r02_parity.md) - the projection is 0.04 % of the FLOPs and HBM-bound either
// way.  Partials go to partial[chunk, b, n] and are reduced by latent_reduce_kernel (deterministic).
template <int MT, bool LO>
__global__ void __launch_bounds__(256, 1)
latent_proj_mma_kernel(const float* __restrict__ pooled, const __nv_bfloat16* __restrict__ wv,
                       const __nv_bfloat16* __restrict__ wv_lo, int B, long long L, int NL, float* __restrict__ partial) {
    const int chunk = blockIdx.x;
    const int b0 = blockIdx.y * (16 * MT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const long long l0 = (long long)chunk * LAT_CH;
    const int len = (int)min((long long)LAT_CH, L - l0);             // multiple of 64
    for (int nbase = 0; nbase < NL; nbase += 512) {
        const int n_w = nbase + warp * 64;                             // this warp's 8 n-tiles: n_w + 8 i + g
        if (n_w >= NL) continue;
        float acc[MT][8][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[m][i][0] = acc[m][i][1] = acc[m][i][2] = acc[m][i][3] = 0.f;
        for (int kb = 0; kb < len; kb += 32) {
            // weight fragments of the 8 n-tiles: 8 independent 16-byte loads in flight per lane
            uint4 wf[8], wl[LO ? 8 : 1];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                wf[i] = *reinterpret_cast<const uint4*>(wv + (long long)(n_w + 8 * i + g) * L + l0 + kb + 8 * t);
                if constexpr (LO)
                    wl[i] = *reinterpret_cast<const uint4*>(wv_lo + (long long)(n_w + 8 * i + g) * L + l0 + kb + 8 * t);
            }
            // pooled fragments (rows g, g+8 of each m-tile), split into bf16 hi / lo
            uint32_t ah[MT][2][4], al[MT][2][4];                      // [m-tile][k-step j][a0..a3]
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int row = b0 + 16 * m + g + 8 * r;
                    float x[8];
                    if (row < B) {
                        const float4 p0 = *reinterpret_cast<const float4*>(pooled + (long long)row * L + l0 + kb + 8 * t);
                        const float4 p1 = *reinterpret_cast<const float4*>(pooled + (long long)row * L + l0 + kb + 8 * t + 4);
                        x[0] = p0.x; x[1] = p0.y; x[2] = p0.z; x[3] = p0.w; x[4] = p1.x; x[5] = p1.y; x[6] = p1.z; x[7] = p1.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {                  // h = 0: logical k 2t,2t+1 ; h = 1: 2t+8,2t+9
                            const float u = x[4 * j + 2 * h], v = x[4 * j + 2 * h + 1];
                            const __nv_bfloat16 uh = __float2bfloat16(u), vh = __float2bfloat16(v);
                            ah[m][j][r + 2 * h] = pack_bf16(__bfloat162float(uh), __bfloat162float(vh));
                            al[m][j][r + 2 * h] = pack_bf16(u - __bfloat162float(uh), v - __bfloat162float(vh));
                        }
                }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t wr[4] = {wf[i].x, wf[i].y, wf[i].z, wf[i].w};
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        mma_bf16_16816(acc[m][i], ah[m][j], wr[2 * j], wr[2 * j + 1]);
                        mma_bf16_16816(acc[m][i], al[m][j], wr[2 * j], wr[2 * j + 1]);
                    }
                if constexpr (LO) {
                    const uint32_t lr[4] = {wl[i].x, wl[i].y, wl[i].z, wl[i].w};
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int m = 0; m < MT; ++m) mma_bf16_16816(acc[m][i], ah[m][j], lr[2 * j], lr[2 * j + 1]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = b0 + 16 * m + g + 8 * r;
                if (row >= B) continue;
                float* o = partial + ((long long)chunk * B + row) * NL + n_w + 2 * t;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float2*>(o + 8 * i) = make_float2(acc[m][i][2 * r], acc[m][i][2 * r + 1]);
            }
    }
}

__global__ void latent_reduce_kernel(const float* __restrict__ partial, int n_chunks, int total, float* __restrict__ latent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float s = 0.f;
    for (int c = 0; c < n_chunks; ++c) s += partial[(long long)c * total + i];
    latent[i] = s;
}

// dpooled[b, l] = sum_n dlatent[b, n] * Wv[n, l]; thread per 8 columns, loops over n (coalesced along l)
__global__ void __launch_bounds__(256)
latent_proj_bwd_kernel(const float* __restrict__ dlatent, const __nv_bfloat16* __restrict__ wv, int B, long long L,
                       int NL, float* __restrict__ dpooled) {
    extern __shared__ float sd[];  // [nb][NL]
    const int b0 = blockIdx.y * LAT_BMAX;
    const int nb = min(LAT_BMAX, B - b0);
    for (int i = threadIdx.x; i < nb * NL; i += blockDim.x) sd[i] = dlatent[(long long)b0 * NL + i];
    __syncthreads();
    const long long l = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (l >= L) return;
    float acc[LAT_BMAX][8];
#pragma unroll
    for (int b = 0; b < LAT_BMAX; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[b][e] = 0.f;
    for (int n = 0; n < NL; ++n) {
        const uint4 w4 = *reinterpret_cast<const uint4*>(wv + (long long)n * L + l);
        const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
        float wf[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 t = unpack_bf16(ww[e]); wf[2 * e] = t.x; wf[2 * e + 1] = t.y; }
#pragma unroll
        for (int b = 0; b < LAT_BMAX; ++b) {
            const float d = (b < nb) ? sd[b * NL + n] : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[b][e] += d * wf[e];
        }
    }
    for (int b = 0; b < nb; ++b) {
        float* o = dpooled + (long long)(b0 + b) * L + l;
        *reinterpret_cast<float4*>(o) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[b][4], acc[b][5], acc[b][6], acc[b][7]);
    }
}

// text latent: one CTA per text row; out = l2norm(Wt e)
__global__ void text_latent_kernel(const float* __restrict__ e, const float* __restrict__ wt, int DT, int NL,
                                   float* __restrict__ out) {
    extern __shared__ float sm[];  // [NL]
    __shared__ float s_norm;
    const float* er = e + (long long)blockIdx.x * DT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < NL; n += blockDim.x >> 5) {
        float acc = 0.f;
        for (int k = lane; k < DT; k += 32) acc += wt[(long long)n * DT + k] * er[k];
        acc = warp_sum(acc);
        if (lane == 0) sm[n] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        float ss = 0.f;
        for (int n = lane; n < NL; n += 32) ss += sm[n] * sm[n];
        ss = warp_sum(ss);
        if (lane == 0) s_norm = sqrtf(ss);
    }
    __syncthreads();
    for (int n = threadIdx.x; n < NL; n += blockDim.x) out[(long long)blockIdx.x * NL + n] = sm[n] / s_norm;
}

// sim[i, j] = (u_i/|u_i|) . t_j * temp ; optional normalised latents and d sim[i, i % Bt] / d u_i
__global__ void latent_sim_kernel(const float* __restrict__ latent, const float* __restrict__ text, int B, int Bt,
                                  int NL, float temp, float* __restrict__ sim, float* __restrict__ image_latents,
                                  float* __restrict__ dlatent) {
    const int i = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float s_inv;
    __shared__ float s_cos;
    const float* u = latent + (long long)i * NL;
    if (warp == 0) {
        float ss = 0.f;
        for (int n = lane; n < NL; n += 32) ss += u[n] * u[n];
        ss = warp_sum(ss);
        if (lane == 0) s_inv = 1.f / sqrtf(ss);
    }
    __syncthreads();
    const float inv = s_inv;
    for (int j = warp; j < Bt; j += blockDim.x >> 5) {
        float d = 0.f;
        for (int n = lane; n < NL; n += 32) d += u[n] * inv * text[(long long)j * NL + n];
        d = warp_sum(d);
        if (lane == 0) {
            sim[(long long)i * Bt + j] = d * temp;
            if (j == i % Bt) s_cos = d;
        }
    }
    __syncthreads();
    const float* tj = text + (long long)(i % Bt) * NL;
    for (int n = threadIdx.x; n < NL; n += blockDim.x) {
        const float un = u[n] * inv;
        if (image_latents) image_latents[(long long)i * NL + n] = un;
        if (dlatent) dlatent[(long long)i * NL + n] = temp * (tj[n] - un * s_cos) * inv;
    }
}

// dlatent_i = sum_j g_ij * temp * (t_j - un_i * cos_ij) / |u_i|
__global__ void latent_sim_bwd_kernel(const float* __restrict__ latent, const float* __restrict__ text,
                                      const float* __restrict__ gsim, int Bt, int NL, float temp,
                                      float* __restrict__ dlatent) {
    extern __shared__ float sm[];   // [Bt] cosines
    __shared__ float s_inv;
    const int i = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* u = latent + (long long)i * NL;
    if (warp == 0) {
        float ss = 0.f;
        for (int n = lane; n < NL; n += 32) ss += u[n] * u[n];
        ss = warp_sum(ss);
        if (lane == 0) s_inv = 1.f / sqrtf(ss);
    }
    __syncthreads();
    const float inv = s_inv;
    for (int j = warp; j < Bt; j += blockDim.x >> 5) {
        float d = 0.f;
        for (int n = lane; n < NL; n += 32) d += u[n] * inv * text[(long long)j * NL + n];
        d = warp_sum(d);
        if (lane == 0) sm[j] = d;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < NL; n += blockDim.x) {
        const float un = u[n] * inv;
        float acc = 0.f;
        for (int j = 0; j < Bt; ++j) acc += gsim[(long long)i * Bt + j] * (text[(long long)j * NL + n] - un * sm[j]);
        dlatent[(long long)i * NL + n] = acc * temp * inv;
    }
}

}  // namespace ctc

using namespace ctc;

extern "C" int ctc_latent_sim_bwd(const float* latent, const float* text_latents, const float* gsim, int B, int Bt,
                                  int NL, float temp, float* dlatent, void* stream) {
    latent_sim_bwd_kernel<<<B, 128, Bt * sizeof(float), (cudaStream_t)stream>>>(latent, text_latents, gsim, Bt, NL,
                                                                                temp, dlatent);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_vq_num_candidates(int K) { return gemm_argmax_candidates(K); }

extern "C" int ctc_vq_argmax(const float* x, const void* x_bf16, int R, int C, const float* codebook,
                             const void* codebook_bf16, int K, float* cand_val, int* cand_idx, int* ind,
                             void* stream) {
    CTC_REQUIRE(C % 4 == 0, "vq: C=%d must be a multiple of 4", C);
    const int n_cand = gemm_argmax_candidates(K);
    if (int e = gemm_bf16(x_bf16, C, codebook_bf16, C, nullptr, 0, R, K, C, CTC_EPI_ARGMAX, nullptr, nullptr, 0, nullptr, 0,
                          cand_val, cand_idx, CTC_GEMM_TCGEN05, (cudaStream_t)stream))
        return e;
    vq_refine_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, R, C, codebook, K, cand_val, cand_idx,
                                                                    n_cand, ind);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_vq_gather_pool(const int* ind, const float* codebook, int B, int T, int HW, int C, float* pooled,
                                  void* pooled_bf16, float* tokens, void* stream) {
    CTC_REQUIRE(C % 4 == 0, "vq: C=%d must be a multiple of 4", C);
    vq_gather_pool_kernel<<<B * HW, 128, 0, (cudaStream_t)stream>>>(ind, codebook, B, T, HW, C, pooled,
                                                                    (__nv_bfloat16*)pooled_bf16, tokens);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_vq_bwd(const float* dpooled, const float* dtokens, const float* x, int B, int T, int HW, int C,
                          int grad_mode, float* dx, void* stream) {
    CTC_REQUIRE(C % 4 == 0, "vq: C=%d must be a multiple of 4", C);
    const long long R = (long long)B * T * HW;
    vq_bwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dpooled, dtokens, x, B, T, HW, C,
                                                                            grad_mode, dx);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_latent_proj(const float* pooled, const void* wv_bf16, const void* wv_lo_bf16, int B, int64_t L, int NL,
                               float* partial, int n_chunks, float* latent, void* stream) {
    const int need = (int)((L + LAT_CH - 1) / LAT_CH);
    CTC_REQUIRE(n_chunks == need, "latent_proj: partial buffer must have %d chunks (got %d)", need, n_chunks);
    CTC_REQUIRE(L % 8 == 0, "latent_proj: L=%lld must be a multiple of 8", (long long)L);
    const bool mma_ok = NL % 64 == 0 && L % 64 == 0 && (reinterpret_cast<uintptr_t>(pooled) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(wv_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(partial) & 7) == 0 &&
                        (reinterpret_cast<uintptr_t>(wv_lo_bf16) & 15) == 0;
    const __nv_bfloat16* wh = (const __nv_bfloat16*)wv_bf16;
    const __nv_bfloat16* wlo = (const __nv_bfloat16*)wv_lo_bf16;
    cudaStream_t st = (cudaStream_t)stream;
    if (mma_ok && B > 16) {
        const dim3 grid(n_chunks, (B + 31) / 32);
        if (wlo) latent_proj_mma_kernel<2, true><<<grid, 256, 0, st>>>(pooled, wh, wlo, B, L, NL, partial);
        else latent_proj_mma_kernel<2, false><<<grid, 256, 0, st>>>(pooled, wh, wlo, B, L, NL, partial);
    } else if (mma_ok) {
        if (wlo) latent_proj_mma_kernel<1, true><<<dim3(n_chunks, 1), 256, 0, st>>>(pooled, wh, wlo, B, L, NL, partial);
        else latent_proj_mma_kernel<1, false><<<dim3(n_chunks, 1), 256, 0, st>>>(pooled, wh, wlo, B, L, NL, partial);
    } else {
        dim3 grid(n_chunks, (B + LAT_BMAX - 1) / LAT_BMAX);
        latent_proj_kernel<<<grid, 256, 0, st>>>(pooled, wh, wlo, B, L, NL, partial);
    }
    CTC_LAUNCH_CHECK();
    const int total = B * NL;
    latent_reduce_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partial, n_chunks, total, latent);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_latent_proj_bwd(const float* dlatent, const void* wv_bf16, int B, int64_t L, int NL, float* dpooled,
                                   void* stream) {
    CTC_REQUIRE(L % 8 == 0, "latent_proj_bwd: L=%lld must be a multiple of 8", (long long)L);
    dim3 grid((unsigned)((L / 8 + 255) / 256), (B + LAT_BMAX - 1) / LAT_BMAX);
    const size_t smem = (size_t)LAT_BMAX * NL * sizeof(float);
    latent_proj_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dlatent, (const __nv_bfloat16*)wv_bf16, B, L, NL,
                                                                      dpooled);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_text_latent(const float* e, const float* wt, int Bt, int DT, int NL, float* out, void* stream) {
    text_latent_kernel<<<Bt, 256, NL * sizeof(float), (cudaStream_t)stream>>>(e, wt, DT, NL, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_latent_sim(const float* latent, const float* text_latents, int B, int Bt, int NL, float temp,
                              float* sim, float* image_latents, float* dlatent, void* stream) {
    latent_sim_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(latent, text_latents, B, Bt, NL, temp, sim, image_latents,
                                                           dlatent);
    CTC_LAUNCH_CHECK();
    return 0;
}
