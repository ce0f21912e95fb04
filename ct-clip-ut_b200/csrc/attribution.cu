// Attribution-side reductions (reference: src/utils/visualizations.py):
//   attention rollout (:707-743 as driven by :800-841), raw-attention query means (:666,671),
//   Grad-CAM channel weights + CAM (:933-991), integrated-gradients combine (:878-879),
//   trilinear up-sampling with the fused rot90 (:289-293, :816).
// All HBM-bound: coalesced 128-bit accesses, warp-shuffle reductions.
#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {

// ---------------------------------------------------------------------------------------------
// spatial "rollout": the reference calls attention_rollout([P_slice]) with ONE matrix, so
//   A = mean_h P;  A /= (rowsum + 1e-8);  A += I;  A /= rowsum;  result = A @ I;  out = colsum(A)
// One CTA per slice.  Pass 1: a warp per row computes the two row normalisers; pass 2: a thread per
// column adds the rows in order.  No atomics: the result is bit-reproducible.  `probs` is
// [n_slices, heads, n, n]; with the head-fused matrix of ctc_attention_fused_probs, heads = 1.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
rollout_spatial_kernel(const float* __restrict__ probs, int heads, int n, float* __restrict__ out) {
    extern __shared__ float rn[];   // [2n]: 1/d1, 1/r2 per row
    const int s = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float invh = 1.f / heads;
    const float* P = probs + (long long)s * heads * n * n;
    for (int i = warp; i < n; i += nw) {
        float rs = 0.f;
        for (int j = lane; j < n; j += 32) {
            float a = 0.f;
            for (int h = 0; h < heads; ++h) a += P[((long long)h * n + i) * n + j];
            rs += a * invh;
        }
        rs = warp_sum(rs);
        if (lane == 0) {
            const float d1 = rs + 1e-8f;
            rn[2 * i] = d1;
            rn[2 * i + 1] = rs / d1 + 1.f;          // second normaliser: sum_j (A_ij / d1) + 1
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < n; ++i) {
            float a = 0.f;
            for (int h = 0; h < heads; ++h) a += P[((long long)h * n + i) * n + j];
            float v = a * invh / rn[2 * i];
            if (j == i) v += 1.f;
            acc += v / rn[2 * i + 1];
        }
        out[(long long)s * n + j] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// generic Visualizations.attention_rollout (visualizations.py:707-743), one layer: head fusion
// (mean / max), optional discard of the lowest weights of each row (keep the k_keep largest:
// `flat.topk(n - num_discard).min()` is the k_keep-th largest value, found exactly by rank counting),
// A /= (rowsum + 1e-8), and with the residual A += I, A /= rowsum.  One CTA per row, n <= 1024.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rollout_fuse_kernel(const float* __restrict__ attn, int heads, int n, int fusion_max, int k_keep, int use_residual,
                    float* __restrict__ out) {
    extern __shared__ float row[];            // [n] fused row, then [8] reduction scratch
    float* red = row + n;
    __shared__ float s_thr;
    const int i = blockIdx.x;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float a = fusion_max ? -3.0e38f : 0.f;
        for (int h = 0; h < heads; ++h) {
            const float v = attn[((long long)h * n + i) * n + j];
            a = fusion_max ? fmaxf(a, v) : a + v;
        }
        row[j] = fusion_max ? a : a / (float)heads;
    }
    __syncthreads();
    if (k_keep < n) {
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const float x = row[j];
            int gt = 0, ge = 0;
            for (int m = 0; m < n; ++m) { const float y = row[m]; gt += y > x; ge += y >= x; }
            if (gt < k_keep && k_keep <= ge) s_thr = x;       // every match carries the same value
        }
        __syncthreads();
        const float thr = s_thr;
        for (int j = threadIdx.x; j < n; j += blockDim.x) if (!(row[j] >= thr)) row[j] = 0.f;
        __syncthreads();
    }
    auto block_sum = [&](float v) {           // fixed-order block reduction
        v = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        return t;
    };
    float part = 0.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) part += row[j];
    const float d1 = block_sum(part) + 1e-8f;
    part = 0.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float v = row[j] / d1;
        if (use_residual && j == i) v += 1.f;
        row[j] = v;
        part += v;
    }
    const float r2 = use_residual ? block_sum(part) : 1.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) out[(long long)i * n + j] = use_residual ? row[j] / r2 : row[j];
}

// C = A @ B, fp32 row-major n x n (the `result = attn @ result` chain of the generic rollout); 32 x 32 tiles
__global__ void __launch_bounds__(1024)
matmul_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, int n, float* __restrict__ C) {
    __shared__ float sa[32][33], sb[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.y * 32 + ty, col = blockIdx.x * 32 + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < n; k0 += 32) {
        sa[ty][tx] = (row < n && k0 + tx < n) ? A[(long long)row * n + k0 + tx] : 0.f;
        sb[ty][tx] = (k0 + ty < n && col < n) ? B[(long long)(k0 + ty) * n + col] : 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) acc = fmaf(sa[ty][k], sb[k][tx], acc);
        __syncthreads();
    }
    if (row < n && col < n) C[(long long)row * n + col] = acc;
}

// dst[i] (+)= sum_b src[b, i] * scale, batch rows added in order (the IG running sum, visualizations.py:872,878)
__global__ void __launch_bounds__(256)
batch_sum_kernel(const float* __restrict__ src, int B, long long n, float scale, int accumulate, float* __restrict__ dst) {
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        float4 a = accumulate ? *reinterpret_cast<const float4*>(dst + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = 0; b < B; ++b) {
            const float4 v = *reinterpret_cast<const float4*>(src + (long long)b * n + i);
            a.x += v.x * scale; a.y += v.y * scale; a.z += v.z * scale; a.w += v.w * scale;
        }
        *reinterpret_cast<float4*>(dst + i) = a;
    }
}

// temporal rollout: chain L layers of [T, T] head-mean matrices per token (T <= 32): one warp per token,
// lane = matrix row.  result = A_L ... A_1 (result = A @ result per layer); out = colsum(result).
__global__ void __launch_bounds__(128)
rollout_temporal_kernel(const float* __restrict__ probs, int n_layers, int n_tok, int heads, int T,
                        float* __restrict__ out) {
    __shared__ float res[4][32][33];
    __shared__ float nxt[4][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tok = blockIdx.x * 4 + warp;
    if (tok >= n_tok) return;
    float (*R)[33] = res[warp];
    float (*N)[33] = nxt[warp];
    for (int j = 0; j < T; ++j) if (lane < T) R[lane][j] = (lane == j) ? 1.f : 0.f;
    __syncwarp();
    const long long layer_stride = (long long)n_tok * heads * T * T;
    for (int l = 0; l < n_layers; ++l) {
        float a[32];
        float rs = 0.f;
        if (lane < T) {
            for (int j = 0; j < T; ++j) {
                float v = 0.f;
                for (int h = 0; h < heads; ++h)
                    v += probs[l * layer_stride + (((long long)tok * heads + h) * T + lane) * T + j];
                a[j] = v / heads;
                rs += a[j];
            }
            const float d1 = rs + 1e-8f;
            float r2 = 0.f;
            for (int j = 0; j < T; ++j) { a[j] = a[j] / d1 + (j == lane ? 1.f : 0.f); r2 += a[j]; }
            for (int j = 0; j < T; ++j) a[j] /= r2;
            for (int c = 0; c < T; ++c) {
                float v = 0.f;
                for (int j = 0; j < T; ++j) v += a[j] * R[j][c];
                N[lane][c] = v;
            }
        }
        __syncwarp();
        if (lane < T) for (int c = 0; c < T; ++c) R[lane][c] = N[lane][c];
        __syncwarp();
    }
    if (lane < T) {
        float v = 0.f;
        for (int i = 0; i < T; ++i) v += R[i][lane];
        out[(long long)tok * T + lane] = v;
    }
}

// out[s, h, j] = mean_i P[s, h, i, j]; one CTA per (s, h), thread per column (coalesced rows)
__global__ void __launch_bounds__(256)
attn_colmean_kernel(const float* __restrict__ probs, int n, float* __restrict__ out) {
    const long long sh = blockIdx.x;
    const float* P = probs + sh * n * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float a = 0.f;
        for (int i = 0; i < n; ++i) a += P[(long long)i * n + j];
        out[sh * n + j] = a / n;
    }
}

// part[blk, c] = sum_r g[r, c] over this CTA's row slab; colmean_final adds the slabs in order (no atomics)
__global__ void __launch_bounds__(256)
colmean_kernel(const float* __restrict__ g, long long R, int C, float* __restrict__ part) {
    const long long rows_per = (R + gridDim.x - 1) / gridDim.x;
    const long long r0 = blockIdx.x * rows_per, r1 = min(R, r0 + rows_per);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (long long r = r0; r < r1; ++r) a += g[r * C + c];
        part[(long long)blockIdx.x * C + c] = a;
    }
}
__global__ void __launch_bounds__(256)
colmean_final_kernel(const float* __restrict__ part, int nblk, int C, float invR, float* __restrict__ w) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f;
    for (int b = 0; b < nblk; ++b) a += part[(long long)b * C + c];
    w[c] = a * invR;
}

// cam[r] = relu(sum_c (fa[r,c] - fb[r,c]) * w[c]); warp per row
__global__ void __launch_bounds__(256)
gradcam_kernel(const float* __restrict__ fa, const float* __restrict__ fb, const float* __restrict__ w, long long R,
               int C, float* __restrict__ cam) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    float a = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        float4 x = *reinterpret_cast<const float4*>(fa + row * C + c);
        if (fb) {
            const float4 y = *reinterpret_cast<const float4*>(fb + row * C + c);
            x.x -= y.x; x.y -= y.y; x.z -= y.z; x.w -= y.w;
        }
        const float4 ww = *reinterpret_cast<const float4*>(w + c);
        a += x.x * ww.x + x.y * ww.y + x.z * ww.z + x.w * ww.w;
    }
    a = warp_sum(a);
    if (lane == 0) cam[row] = fmaxf(a, 0.f);
}

// F.interpolate(mode='trilinear', align_corners=False): src = max(0, (dst + 0.5) * in/out - 0.5)
CTC_DEVINL void lerp_coord(int dst, int in, int out, int& i0, int& i1, float& w1) {
    float s = ((float)dst + 0.5f) * ((float)in / (float)out) - 0.5f;
    s = fmaxf(s, 0.f);
    i0 = min((int)s, in - 1);
    i1 = min(i0 + 1, in - 1);
    w1 = s - (float)i0;
}
// thread per 4 consecutive output voxels along the fastest output axis
__global__ void __launch_bounds__(256)
upsample_kernel(const float* __restrict__ in, int d, int h, int w, float* __restrict__ out, int D, int H, int W, int rot) {
    // un-rotated result up[z, y, x]; np.rot90(up, k=-1, axes=(1,2)) -> out[z, x, H-1-y] with shape [D, W, H]
    const int OY = rot ? W : H, OX = rot ? H : W;
    const long long total = (long long)D * OY * (OX / 4);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox4 = (int)(idx % (OX / 4)) * 4;
    const int oy = (int)((idx / (OX / 4)) % OY);
    const int z = (int)(idx / ((long long)(OX / 4) * OY));
    int z0, z1; float wz;
    lerp_coord(z, d, D, z0, z1, wz);
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int ox = ox4 + e;
        const int y = rot ? (H - 1 - ox) : oy;
        const int x = rot ? oy : ox;
        int y0, y1, x0, x1; float wy, wx;
        lerp_coord(y, h, H, y0, y1, wy);
        lerp_coord(x, w, W, x0, x1, wx);
        auto at = [&](int zz, int yy, int xx) { return in[((long long)zz * h + yy) * w + xx]; };
        const float c00 = at(z0, y0, x0) * (1.f - wx) + at(z0, y0, x1) * wx;
        const float c01 = at(z0, y1, x0) * (1.f - wx) + at(z0, y1, x1) * wx;
        const float c10 = at(z1, y0, x0) * (1.f - wx) + at(z1, y0, x1) * wx;
        const float c11 = at(z1, y1, x0) * (1.f - wx) + at(z1, y1, x1) * wx;
        const float c0 = c00 * (1.f - wy) + c01 * wy;
        const float c1 = c10 * (1.f - wy) + c11 * wy;
        r[e] = c0 * (1.f - wz) + c1 * wz;
    }
    *reinterpret_cast<float4*>(out + ((long long)z * OY + oy) * OX + ox4) = make_float4(r[0], r[1], r[2], r[3]);
}

CTC_DEVINL void atomic_min_float(float* addr, float v) {   // v >= 0 only (relu output)
    atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
}
CTC_DEVINL void atomic_max_float(float* addr, float v) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
__global__ void __launch_bounds__(256)
ig_combine_kernel(const float* __restrict__ vol, const float* __restrict__ gsum, long long n, float inv_steps,
                  float* __restrict__ ig, float* __restrict__ mm) {
    float mn = 3.0e38f, mx = 0.f;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        const float4 x = *reinterpret_cast<const float4*>(vol + i);
        const float4 g = *reinterpret_cast<const float4*>(gsum + i);
        float4 o;
        o.x = fmaxf((x.x - 1.f) * (g.x * inv_steps), 0.f); o.y = fmaxf((x.y - 1.f) * (g.y * inv_steps), 0.f);
        o.z = fmaxf((x.z - 1.f) * (g.z * inv_steps), 0.f); o.w = fmaxf((x.w - 1.f) * (g.w * inv_steps), 0.f);
        *reinterpret_cast<float4*>(ig + i) = o;
        mn = fminf(mn, fminf(fminf(o.x, o.y), fminf(o.z, o.w)));
        mx = fmaxf(mx, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
    }
    mx = warp_max(mx);
    mn = -warp_max(-mn);
    if ((threadIdx.x & 31) == 0) { atomic_min_float(mm, mn); atomic_max_float(mm + 1, mx); }
}

// global min / max of an fp32 array of arbitrary sign (mm[0] = min, mm[1] = max; caller pre-sets +inf / -inf)
CTC_DEVINL void atomic_min_any(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
CTC_DEVINL void atomic_max_any(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ x, long long n, float* __restrict__ mm) {
    float mn = 3.0e38f, mx = -3.0e38f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    mn = -warp_max(-mn);
    if ((threadIdx.x & 31) == 0) { atomic_min_any(mm, mn); atomic_max_any(mm + 1, mx); }
}

// the reference's three normalisations + optional np.rot90(k=-1, axes=(1,2)):
//   mode 0: (v-min)/(max+1e-8)   mode 1: (v-min)/(max-min+1e-8)   mode 2: v/(max+1e-8)
// out[z, x, H-1-y] = f(in[z, y, x]) when rot (out shape [D, W, H]).
__global__ void __launch_bounds__(256)
normalize_kernel(const float* __restrict__ in, int D, int H, int W, const float* __restrict__ mm, int mode, int rot,
                 float* __restrict__ out) {
    const int OY = rot ? W : H, OX = rot ? H : W;
    const long long total = (long long)D * OY * OX;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox = (int)(idx % OX), oy = (int)((idx / OX) % OY), z = (int)(idx / ((long long)OX * OY));
    const int y = rot ? (H - 1 - ox) : oy, x = rot ? oy : ox;
    const float v = in[((long long)z * H + y) * W + x];
    const float mn = mm[0], mx = mm[1];
    float o;
    if (mode == 0) o = (v - mn) / (mx + 1e-8f);
    else if (mode == 1) o = (v - mn) / (mx - mn + 1e-8f);
    else o = v / (mx + 1e-8f);
    out[idx] = o;
}

// 16-bit radix histogram of the fp32 bit patterns (non-negative values: bit order == numeric order).
// Counts (bits >> shift) & 0xffff of the elements whose bits above (shift+16) equal the prefix (shift == 0 only;
// the prefix comes from `prefix_dev` when given, so that a two-pass selection needs no host round trip).
// Exact zeros — half of a relu'd integrated-gradients map — are counted per warp with one atomic.
__global__ void __launch_bounds__(256)
hist16_kernel(const float* __restrict__ x, long long n, int shift, unsigned int prefix,
              const unsigned int* __restrict__ prefix_dev, unsigned int* __restrict__ hist) {
    if (prefix_dev) prefix = *prefix_dev;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long base = i0 - threadIdx.x % 32; base < n; base += stride) {     // warp-uniform trip count
        const long long i = base + threadIdx.x % 32;
        const bool in = i < n;
        const unsigned int b = in ? __float_as_uint(x[i]) : 0xffffffffu;
        const bool take = in && !(shift == 0 && (b >> 16) != prefix);
        const bool zero = take && ((b >> shift) & 0xffffu) == 0u;
        const unsigned int zmask = __ballot_sync(0xffffffffu, zero);
        if (zmask && (threadIdx.x % 32) == (unsigned)(__ffs(zmask) - 1)) atomicAdd(&hist[0], (unsigned)__popc(zmask));
        if (take && !zero) atomicAdd(&hist[(b >> shift) & 0xffffu], 1u);
    }
}

// Bucket of a 65536-bin histogram that holds the element of 0-based rank k (one CTA of 1024 threads, 64 bins each).
//   pass 0: k = k_imm;            writes sel[0] = bucket, rank_left = k - (elements before the bucket)
//   pass 1: k = rank_left (read); writes the selected VALUE bits (sel[0] << 16 | bucket) to out
struct KthState { unsigned long long rank_left; unsigned int hi; unsigned int pad; };
__global__ void __launch_bounds__(1024)
hist_select_kernel(const unsigned int* __restrict__ hist, long long k_imm, int pass, KthState* __restrict__ st,
                   float* __restrict__ out) {
    __shared__ unsigned long long wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long k = pass == 0 ? (unsigned long long)k_imm : st->rank_left;
    unsigned long long local = 0;
    for (int b = 0; b < 64; ++b) local += hist[tid * 64 + b];
    unsigned long long inc = local;                                   // inclusive scan inside the warp
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned long long before = inc - local;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (k >= before && k < before + local) {                          // exactly one thread
        unsigned long long c = before;
        for (int b = 0; b < 64; ++b) {
            const unsigned int h = hist[tid * 64 + b];
            if (k < c + h) {
                const unsigned int bucket = tid * 64 + b;
                if (pass == 0) { st->hi = bucket; st->rank_left = k - c; }
                else *out = __uint_as_float((st->hi << 16) | bucket);
                break;
            }
            c += h;
        }
    }
}

// np.quantile's linear interpolation between the two order statistics, in float32 like NumPy 2.x does for an fp32 array
__global__ void quantile_lerp_kernel(const float* __restrict__ a_dev, const float* __restrict__ b_dev, float gamma,
                                     float* __restrict__ out) {
    const float a = *a_dev, b = *b_dev;
    const float d = b - a;
    float r = a + d * gamma;
    if (gamma >= 0.5f) r = b - d * (1.f - gamma);
    *out = r;
}

// integrated-gradients finalisation, second half (visualizations.py:882-901):
//   n1 = (ig - min)/(max + 1e-8);  n2 = n1 >= q ? n1 : 0;  n3 = n2 ** 0.05;  out = n3 / (max(n3) + 1e-8);  rot90
__global__ void __launch_bounds__(256)
ig_finalize_kernel(const float* __restrict__ ig, int D, int H, int W, float mn, float mx, float q, float inv_m3, int rot,
                   float* __restrict__ out) {
    const int OY = rot ? W : H, OX = rot ? H : W;
    const long long total = (long long)D * OY * OX;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox = (int)(idx % OX), oy = (int)((idx / OX) % OY), z = (int)(idx / ((long long)OX * OY));
    const int y = rot ? (H - 1 - ox) : oy, x = rot ? oy : ox;
    const float n1 = (ig[((long long)z * H + y) * W + x] - mn) / (mx + 1e-8f);
    const float n2 = (n1 >= q) ? n1 : 0.f;
    out[idx] = (n2 > 0.f ? powf(n2, 0.05f) : 0.f) * inv_m3;
}

// same, with min / max / quantile read from device memory (no host round trip between the IG stages).  The
// final scale follows the reference's float32 pipeline literally: n1max = (max-min)/(max+1e-8),
// m3 = n1max >= q ? n1max ** 0.05 : 0, out = n3 / (m3 + 1e-8).
__global__ void __launch_bounds__(256)
ig_finalize_dev_kernel(const float* __restrict__ ig, int D, int H, int W, const float* __restrict__ mm,
                       const float* __restrict__ q_dev, int rot, float* __restrict__ out) {
    const int OY = rot ? W : H, OX = rot ? H : W;
    const long long total = (long long)D * OY * OX;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const float mn = mm[0], mx = mm[1], q = *q_dev;
    const float n1max = (mx - mn) / (mx + 1e-8f);
    const float m3 = (n1max >= q && n1max > 0.f) ? powf(n1max, 0.05f) : 0.f;
    const int ox = (int)(idx % OX), oy = (int)((idx / OX) % OY), z = (int)(idx / ((long long)OX * OY));
    const int y = rot ? (H - 1 - ox) : oy, x = rot ? oy : ox;
    const float n1 = (ig[((long long)z * H + y) * W + x] - mn) / (mx + 1e-8f);
    const float n2 = (n1 >= q) ? n1 : 0.f;
    out[idx] = (n2 > 0.f ? powf(n2, 0.05f) : 0.f) / (m3 + 1e-8f);
}

// occlusion heat map (visualizations.py:366-367, 390-392, 411-413): windows form a regular grid
// (d0 = i*sd, ...), so voxel (z,y,x) is covered by at most ceil(p/s)^3 of them; sum their importances
// in double (the reference accumulates float64 numpy arrays), divide by the count (0 -> 1).
// imp fp32 [nd, nh, nw]; inc uint8 [nd, nh, nw] marks windows that were evaluated (sharding may drop some).
__global__ void __launch_bounds__(256)
occlusion_heat_kernel(const float* __restrict__ imp, const unsigned char* __restrict__ inc, int nd, int nh, int nw,
                      int pd, int ph, int pw, int sd, int sh, int sw, int D, int H, int W, float* __restrict__ heat) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)D * H * W) return;
    const int x = (int)(idx % W), y = (int)((idx / W) % H), z = (int)(idx / ((long long)W * H));
    auto lo = [](int v, int p, int s) { const int t = v - p + 1; return t <= 0 ? 0 : (t + s - 1) / s; };
    const int i0 = lo(z, pd, sd), i1 = min(z / sd, nd - 1);
    const int j0 = lo(y, ph, sh), j1 = min(y / sh, nh - 1);
    const int k0 = lo(x, pw, sw), k1 = min(x / sw, nw - 1);
    double acc = 0.0;
    int cnt = 0;
    for (int i = i0; i <= i1; ++i)
        for (int j = j0; j <= j1; ++j)
            for (int k = k0; k <= k1; ++k) {
                const int w = (i * nh + j) * nw + k;
                if (inc[w]) { acc += (double)imp[w]; ++cnt; }
            }
    heat[idx] = (float)acc / (float)(cnt == 0 ? 1 : cnt);
}

}  // namespace ctc

namespace ctc {
// flags[t, h, w] = 1 iff every voxel of patch (t, h, w) equals `value` (one CTA per patch, 128-bit loads).
// An occlusion window whose patches are all already filled with the occlusion value leaves the volume
// unchanged, so its score equals the un-occluded score exactly (visualizations.py:380-390 would compute
// importance = 0 for it after a full forward).
__global__ void __launch_bounds__(128)
patch_is_constant_kernel(const float* __restrict__ vol, int D, int H, int W, int pt, int p, float value,
                         unsigned char* __restrict__ flags) {
    const int Wp = W / p, Hp = H / p;
    const int wp = blockIdx.x % Wp, hp = (blockIdx.x / Wp) % Hp, tp = blockIdx.x / (Wp * Hp);
    const int q = p >> 2;                                   // float4 per patch row
    int bad = 0;
    for (int i = threadIdx.x; i < pt * p * q; i += blockDim.x) {
        const int c4 = i % q, y = (i / q) % p, d = i / (q * p);
        const float4 v = *reinterpret_cast<const float4*>(vol + ((long long)(tp * pt + d) * H + hp * p + y) * W + wp * p + c4 * 4);
        bad |= (v.x != value) | (v.y != value) | (v.z != value) | (v.w != value);
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) flags[blockIdx.x] = bad ? 0 : 1;
}

// Zero-shot scoring: sim [B, 2P] holds (present, absent) logit pairs; out[b, j] = softmax(pair)[0] as float64.
__global__ void pair_softmax_kernel(const float* __restrict__ sim, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // softmax([a, b])[0] = 1 / (1 + e^(b-a)), evaluated in double and rounded to the reference's fp32 result type
    const double d = (double)sim[2 * i + 1] - (double)sim[2 * i];
    out[i] = (double)(float)(1.0 / (1.0 + exp(d)));
}
}  // namespace ctc

using namespace ctc;

extern "C" int ctc_pair_softmax(const float* sim, int B, int P, double* out, void* stream) {
    CTC_REQUIRE(B >= 0 && P >= 0, "pair_softmax: B=%d P=%d", B, P);
    const int n = B * P;
    if (n == 0) return 0;
    pair_softmax_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sim, n, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_patch_is_constant(const float* volume, int D, int H, int W, int pt, int p, float value,
                                     unsigned char* flags, void* stream) {
    CTC_REQUIRE(D % pt == 0 && H % p == 0 && W % p == 0 && p % 4 == 0,
                "patch_is_constant: volume %dx%dx%d / patch %dx%dx%d (p must be a multiple of 4)", D, H, W, pt, p, p);
    CTC_REQUIRE((reinterpret_cast<uintptr_t>(volume) & 15) == 0, "patch_is_constant: volume not 16-byte aligned");
    const int n = (D / pt) * (H / p) * (W / p);
    patch_is_constant_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(volume, D, H, W, pt, p, value, flags);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_minmax(const float* x, int64_t n, float* mm, void* stream) {
    minmax_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(x, n, mm);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_normalize(const float* in, int D, int H, int W, const float* mm, int mode, int rot90, float* out,
                             void* stream) {
    CTC_REQUIRE(in != out || !rot90, "normalize: rot90 cannot run in place");
    const long long total = (long long)D * H * W;
    normalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, D, H, W, mm, mode, rot90, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_hist16(const float* x, int64_t n, int shift, unsigned int prefix, unsigned int* hist, void* stream) {
    CTC_REQUIRE(shift == 0 || shift == 16, "hist16: shift must be 0 or 16");
    CTC_CHECK_CUDA(cudaMemsetAsync(hist, 0, 65536 * sizeof(unsigned int), (cudaStream_t)stream));
    hist16_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(x, n, shift, prefix, nullptr, hist);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_kth_value_ws_bytes(void) { return 65536 * 4 + (int)sizeof(KthState); }

extern "C" int ctc_kth_value(const float* x, int64_t n, int64_t k, void* ws, float* out_dev, void* stream) {
    CTC_REQUIRE(n > 0 && k >= 0 && k < n, "kth_value: k=%lld outside [0, %lld)", (long long)k, (long long)n);
    CTC_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7) == 0, "kth_value: workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* hist = reinterpret_cast<unsigned int*>(ws);
    KthState* state = reinterpret_cast<KthState*>(hist + 65536);
    for (int pass = 0; pass < 2; ++pass) {
        CTC_CHECK_CUDA(cudaMemsetAsync(hist, 0, 65536 * sizeof(unsigned int), st));
        hist16_kernel<<<148 * 8, 256, 0, st>>>(x, n, pass == 0 ? 16 : 0, 0u, pass == 0 ? nullptr : &state->hi, hist);
        CTC_LAUNCH_CHECK();
        hist_select_kernel<<<1, 1024, 0, st>>>(hist, k, pass, state, out_dev);
        CTC_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int ctc_quantile_lerp(const float* a_dev, const float* b_dev, float gamma, float* out_dev, void* stream) {
    quantile_lerp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a_dev, b_dev, gamma, out_dev);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_ig_finalize_dev(const float* ig, int D, int H, int W, const float* mm_dev, const float* q_dev,
                                   int rot90, float* out, void* stream) {
    const long long total = (long long)D * H * W;
    ig_finalize_dev_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ig, D, H, W, mm_dev, q_dev,
                                                                                             rot90, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_ig_finalize(const float* ig, int D, int H, int W, float mn, float mx, float q, float inv_m3,
                               int rot90, float* out, void* stream) {
    const long long total = (long long)D * H * W;
    ig_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ig, D, H, W, mn, mx, q, inv_m3,
                                                                                         rot90, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_occlusion_heatmap(const float* imp, const unsigned char* inc, int nd, int nh, int nw, int pd, int ph,
                                     int pw, int sd, int sh, int sw, int D, int H, int W, float* heat, void* stream) {
    const long long total = (long long)D * H * W;
    occlusion_heat_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        imp, inc, nd, nh, nw, pd, ph, pw, sd, sh, sw, D, H, W, heat);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_rollout_spatial(const float* probs, int n_slices, int heads, int n, float* out, void* stream) {
    CTC_REQUIRE(n > 0 && (size_t)n * 8 <= 48 * 1024, "rollout_spatial: n=%d outside (0, 6144]", n);
    rollout_spatial_kernel<<<n_slices, 512, 2 * n * sizeof(float), (cudaStream_t)stream>>>(probs, heads, n, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_rollout_fuse(const float* attn, int heads, int n, int fusion, int k_keep, int use_residual,
                                float* out, void* stream) {
    CTC_REQUIRE(heads > 0 && n > 0 && n <= 4096, "rollout_fuse: heads=%d n=%d", heads, n);
    CTC_REQUIRE(fusion == 0 || fusion == 1, "rollout_fuse: fusion must be 0 (mean) or 1 (max), got %d", fusion);
    CTC_REQUIRE(k_keep >= 1 && k_keep <= n, "rollout_fuse: k_keep=%d outside [1, %d]", k_keep, n);
    rollout_fuse_kernel<<<n, 256, (n + 8) * sizeof(float), (cudaStream_t)stream>>>(attn, heads, n, fusion, k_keep,
                                                                                  use_residual, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_matmul_f32(const float* A, const float* B, int n, float* C, void* stream) {
    CTC_REQUIRE(n > 0 && C != A && C != B, "matmul_f32: n=%d, output must not alias an input", n);
    dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 32);
    matmul_f32_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(A, B, n, C);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_batch_sum(const float* src, int B, int64_t n, float scale, int accumulate, float* dst, void* stream) {
    CTC_REQUIRE(B > 0 && n > 0 && n % 4 == 0, "batch_sum: B=%d n=%lld (n must be a multiple of 4)", B, (long long)n);
    CTC_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
                "batch_sum: buffers must be 16-byte aligned");
    batch_sum_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(src, B, n, scale, accumulate, dst);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_rollout_temporal(const float* probs, int n_layers, int n_tok, int heads, int T, float* out,
                                    void* stream) {
    CTC_REQUIRE(T <= 32, "rollout_temporal: T=%d exceeds 32", T);
    rollout_temporal_kernel<<<(n_tok + 3) / 4, 128, 0, (cudaStream_t)stream>>>(probs, n_layers, n_tok, heads, T, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_attn_colmean(const float* probs, int n_seq, int heads, int n, float* out, void* stream) {
    attn_colmean_kernel<<<n_seq * heads, 256, 0, (cudaStream_t)stream>>>(probs, n, out);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_colmean_ws_floats(int R, int C) { return (R < 592 ? R : 592) * C; }

extern "C" int ctc_colmean(const float* g, int R, int C, float* w, float* ws, void* stream) {
    CTC_REQUIRE(R > 0 && C > 0 && ws != nullptr, "colmean: R=%d C=%d, workspace of ctc_colmean_ws_floats(R, C) floats required", R, C);
    const int grid = R < 592 ? R : 592;
    colmean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, R, C, ws);
    CTC_LAUNCH_CHECK();
    colmean_final_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ws, grid, C, 1.f / (float)R, w);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_gradcam(const float* fa, const float* fb, const float* w, int R, int C, float* cam, void* stream) {
    CTC_REQUIRE(C % 4 == 0, "gradcam: C=%d must be a multiple of 4", C);
    gradcam_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(fa, fb, w, R, C, cam);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_upsample_trilinear(const float* in, int d, int h, int w, float* out, int D, int H, int W, int rot90,
                                      void* stream) {
    CTC_REQUIRE(H % 4 == 0 && W % 4 == 0, "upsample: H=%d, W=%d must be multiples of 4", H, W);
    const long long total = (long long)D * H * W / 4;
    upsample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, d, h, w, out, D, H, W, rot90);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_ig_combine(const float* volume, const float* gsum, int64_t n, float inv_steps, float* ig, float* mm,
                              void* stream) {
    CTC_REQUIRE(n % 4 == 0, "ig_combine: n must be a multiple of 4");
    ig_combine_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(volume, gsum, n, inv_steps, ig, mm);
    CTC_LAUNCH_CHECK();
    return 0;
}
