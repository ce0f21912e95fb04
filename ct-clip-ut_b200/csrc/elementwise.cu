// HBM-bound row / stencil kernels of the CTViT path: LayerNorm fwd/bwd, PEG depthwise 3x3x3
// stencil (+adjoint), GEGLU fwd/bwd, continuous-position-bias table.
// All are bandwidth kernels: 128-bit vector accesses, one warp per 512-wide row, warp-shuffle
// reductions, grids sized well above 148 SMs x resident CTAs.
#include <limits.h>

#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {

// ---------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row. Row (<= 4 KB) is re-read from L1 for the three passes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, int R, int C, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y_bf16,
                     float* __restrict__ y_f32, __nv_bfloat16* __restrict__ xraw) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const float* xr = x + (long long)row * C;
    float s = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        float4 o;
        o.x = (v.x - mean) * rstd * g.x; o.y = (v.y - mean) * rstd * g.y;
        o.z = (v.z - mean) * rstd * g.z; o.w = (v.w - mean) * rstd * g.w;
        if (beta) {
            const float4 b = *reinterpret_cast<const float4*>(beta + c);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        const long long off = (long long)row * C + c;
        if (y_f32) *reinterpret_cast<float4*>(y_f32 + off) = o;
        if (y_bf16) *reinterpret_cast<uint2*>(y_bf16 + off) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        if (xraw) *reinterpret_cast<uint2*>(xraw + off) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

// LayerNorm backward (input gradient only — attribution never needs weight gradients).
// dy is read as fp32, or (dy16 != nullptr) as bf16
CTC_DEVINL float4 load_dy4(const float* dy, const __nv_bfloat16* dy16, long long off) {
    if (dy16) {
        const uint2 u = *reinterpret_cast<const uint2*>(dy16 + off);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return *reinterpret_cast<const float4*>(dy + off);
}
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ x,
                     int R, int C, const float* __restrict__ gamma, float eps, float* __restrict__ out, int accumulate,
                     __nv_bfloat16* __restrict__ out_bf16) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const float* xr = x + (long long)row * C;
    const long long gro = (long long)row * C;
    float s = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    float sg = 0.f, sgx = 0.f;  // sum(g), sum(g * xhat), g = dy * gamma
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float4 d = load_dy4(dy, dy16, gro + c);
        const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
        const float g0 = d.x * gm.x, g1 = d.y * gm.y, g2 = d.z * gm.z, g3 = d.w * gm.w;
        sg += (g0 + g1) + (g2 + g3);
        sgx += g0 * (v.x - mean) * rstd + g1 * (v.y - mean) * rstd + g2 * (v.z - mean) * rstd + g3 * (v.w - mean) * rstd;
    }
    sg = warp_sum(sg) / C;
    sgx = warp_sum(sgx) / C;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float4 d = load_dy4(dy, dy16, gro + c);
        const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
        float4 o;
        o.x = rstd * (d.x * gm.x - sg - (v.x - mean) * rstd * sgx);
        o.y = rstd * (d.y * gm.y - sg - (v.y - mean) * rstd * sgx);
        o.z = rstd * (d.z * gm.z - sg - (v.z - mean) * rstd * sgx);
        o.w = rstd * (d.w * gm.w - sg - (v.w - mean) * rstd * sgx);
        const long long off = (long long)row * C + c;
        if (accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(out + off);
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        *reinterpret_cast<float4*>(out + off) = o;
        if (out_bf16)
            *reinterpret_cast<uint2*>(out_bf16 + off) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
}

// C = 512 specialisation: the row (16 floats per lane) and its gradient live in registers, so every global
// load of a row (x, dy, the accumulate target) is issued up front and nothing is re-read.  Same summation
// order as the generic kernel (bit-identical results).
template <bool ACC, bool BF16OUT, bool DY16>
__global__ void __launch_bounds__(256)
layernorm_bwd_c512_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ x,
                          int R, const float* __restrict__ gamma, float eps, float* __restrict__ out,
                          __nv_bfloat16* __restrict__ out_bf16) {
    constexpr int C = 512;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const long long base = (long long)row * C + lane * 4;
    float4 v[4], d[4], gm[4], pr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = *reinterpret_cast<const float4*>(x + base + 128 * k);
        d[k] = load_dy4(DY16 ? nullptr : dy, DY16 ? dy16 : nullptr, base + 128 * k);
        if (ACC) pr[k] = *reinterpret_cast<const float4*>(out + base + 128 * k);
        gm[k] = *reinterpret_cast<const float4*>(gamma + lane * 4 + 128 * k);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = v[k].x - mean, b = v[k].y - mean, cc = v[k].z - mean, e = v[k].w - mean;
        q += (a * a + b * b) + (cc * cc + e * e);
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float g0 = d[k].x * gm[k].x, g1 = d[k].y * gm[k].y, g2 = d[k].z * gm[k].z, g3 = d[k].w * gm[k].w;
        sg += (g0 + g1) + (g2 + g3);
        sgx += g0 * (v[k].x - mean) * rstd + g1 * (v[k].y - mean) * rstd + g2 * (v[k].z - mean) * rstd +
               g3 * (v[k].w - mean) * rstd;
    }
    sg = warp_sum(sg) / C;
    sgx = warp_sum(sgx) / C;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float4 o;
        o.x = rstd * (d[k].x * gm[k].x - sg - (v[k].x - mean) * rstd * sgx);
        o.y = rstd * (d[k].y * gm[k].y - sg - (v[k].y - mean) * rstd * sgx);
        o.z = rstd * (d[k].z * gm[k].z - sg - (v[k].z - mean) * rstd * sgx);
        o.w = rstd * (d[k].w * gm[k].w - sg - (v[k].w - mean) * rstd * sgx);
        if (ACC) { o.x += pr[k].x; o.y += pr[k].y; o.z += pr[k].z; o.w += pr[k].w; }
        *reinterpret_cast<float4*>(out + base + 128 * k) = o;
        if (BF16OUT)
            *reinterpret_cast<uint2*>(out_bf16 + base + 128 * k) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
}

// ---------------------------------------------------------------------------------------------
// PEG: depthwise 3x3x3 stencil over the token grid, channels-last.
// Kernel tap (a, b, c) (the Conv3d kernel indices over the reinterpreted (t',h',w') axes, with
// causal padding (2,0) on t' and (1,1) on h', w') maps to a canonical-grid offset:
//   SPATIAL : (dt, dh, dw) = (a-2, b-1, c-1)
//   TEMPORAL: (t',h',w') = (h, w, t)  =>  (dt, dh, dw) = (c-1, a-2, b-1)
// The adjoint (sign = -1) gathers with the negated offsets.
//
// HBM-bound target (28 MB in / 28 MB out per volume).  What actually limits a stencil like this on
// sm_100 is instruction issue: 27 FMAs per output on the fma pipe (one 3-register FFMA per 2 cycles per
// SM sub-partition) and one LSU slot per global load.  So a thread owns a channel PAIR and
//   * every load is a 64-bit LDG (a warp covers 64 consecutive channels = two 128-byte lines),
//   * every multiply-add is a packed FFMA2 (fma.rn.f32x2: two fp32 FMAs per issue slot),
//   * it walks along w with a 9-row x 3-column register window (9 new loads per output pair instead
//     of 27) rotating through 5 column slots, so two columns of loads are always in flight,
//   * the 27 weight pairs live in registers, and the 12 warps of a CTA cover a 2 (t) x 6 (h) tile of
//     the same channel chunk so that the remaining 9x row reuse is served by L1.
// ---------------------------------------------------------------------------------------------
static constexpr int PEG_TT = 2, PEG_TH = 6;
typedef unsigned long long f32x2;   // two packed fp32 (lo = even channel)

CTC_DEVINL f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
CTC_DEVINL f32x2 fadd2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
CTC_DEVINL f32x2 pack2(float lo, float hi) {
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
CTC_DEVINL float2 unpack2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// One output column.  A,B,D are the register columns holding x[w-1], x[w], x[w+1] for the 9 (dt,dh) rows;
// N is the column slot being (re)filled with x[w+3] (two columns stay in flight: w+2, w+3).
// With the 5-slot rotation fully unrolled no register moves remain, and with a compile-time channel count
// every load/store of an unrolled block is "base register + immediate".
#define PEG_STEP(A, B, D, N, WOFF, LOADCOL)                                                          \
    {                                                                                                \
        const int wc = w0 + (WOFF);                                                                  \
        if (LOADCOL) {                                                                               \
            _Pragma("unroll") for (int k9 = 0; k9 < 9; ++k9)                                         \
                N[k9] = *reinterpret_cast<const f32x2*>(rp[k9] + (long long)(wc + 3) * CS);          \
        } else {                                                                                     \
            _Pragma("unroll") for (int k9 = 0; k9 < 9; ++k9) N[k9] = 0ull;                           \
        }                                                                                            \
        f32x2 acc = fadd2(B[7], bv), acc1 = 0ull, acc2 = 0ull;                                       \
        _Pragma("unroll") for (int k9 = 0; k9 < 9; ++k9) {                                           \
            acc = ffma2(wk[k9][0], A[k9], acc);                                                      \
            acc1 = ffma2(wk[k9][1], B[k9], acc1);                                                    \
            acc2 = ffma2(wk[k9][2], D[k9], acc2);                                                    \
        }                                                                                            \
        acc = fadd2(acc, fadd2(acc1, acc2));                                                         \
        *reinterpret_cast<f32x2*>(yo + (long long)wc * CS) = acc;                                    \
        if (BF16OUT) {                                                                               \
            const float2 af = unpack2(acc);                                                          \
            *reinterpret_cast<uint32_t*>(yb + (long long)wc * CS) = pack_bf16(af.x, af.y);           \
        }                                                                                            \
    }

// CC, CW = compile-time channel count / row length (0: runtime); SIGN = +1 forward, -1 adjoint;
// BF16OUT = also emit a bf16 copy.
// FRAMES (forward, spatial mode only): the stencil runs over a COMPACT list of T output frames whose three
// causal source frames (dt = -2,-1,0) come from a table: frame_src[f*3 + k] >= 0 selects frame v of `x`,
// v < 0 selects frame (-1 - v) of `xb`, INT_MIN is the causal zero pad.  This is what lets an occlusion window
// recompute only the frames its cube can reach (the causal stencil widens the changed set by two frames per
// layer) and read every other frame from the cached baseline activations.
template <int CC, int CW, int SIGN, bool BF16OUT, bool FRAMES>
__global__ void __launch_bounds__(PEG_TT * PEG_TH * 32)
peg_kernel(const float* __restrict__ x, int B, int T, int H, int W, int C, const float* __restrict__ w27,
           const float* __restrict__ bias, int mode, float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16,
           const float* __restrict__ xb, const int* __restrict__ frame_src) {
    constexpr int sign = SIGN;
    const int CS = CC ? CC : C;
    if (CW) W = CW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + lane * 2;
    const int h = blockIdx.y * PEG_TH + (warp % PEG_TH);
    const int tiles_t = (T + PEG_TT - 1) / PEG_TT;
    const int t = (blockIdx.z % tiles_t) * PEG_TT + warp / PEG_TH;
    const int b = blockIdx.z / tiles_t;
    if (h >= H || t >= T || c >= CS) return;
    // weights wk[k9][cw] for the 9 non-w taps x 3 w taps; rows outside the grid alias the centre row with
    // zero weights, so that every load of the main loop is unconditional
    f32x2 wk[9][3];
    const float* rp[9];
    const long long frame = (long long)H * W * CS;
    const float* fbase[3] = {nullptr, nullptr, nullptr};
    if (FRAMES) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int v = frame_src[t * 3 + k];
            fbase[k] = (v == INT_MIN) ? nullptr : (v >= 0 ? x + v * frame : xb + (long long)(-1 - v) * frame);
        }
    }
    const float* self = FRAMES ? fbase[2] + (long long)h * W * CS + c
                               : x + ((((long long)b * T + t) * H + h) * W) * CS + c;
#pragma unroll
    for (int k9 = 0; k9 < 9; ++k9) {
        const int p = k9 / 3, q = k9 % 3;
        int dt, dh;
        if (FRAMES || mode == CTC_MODE_SPATIAL) { dt = p - 2; dh = q - 1; } else { dh = p - 2; dt = q - 1; }
        const int tt = t + sign * dt, hh = h + sign * dh;
        const bool valid = FRAMES ? (fbase[p] != nullptr && hh >= 0 && hh < H)
                                  : (tt >= 0 && tt < T && hh >= 0 && hh < H);
#pragma unroll
        for (int cw = 0; cw < 3; ++cw) {
            // the adjoint reads x[w - (cw - 1)]: store its w-taps mirrored so that the step body is identical
            const int cwm = (SIGN > 0) ? cw : 2 - cw;
            const int tap = (FRAMES || mode == CTC_MODE_SPATIAL) ? (p * 3 + q) * 3 + cwm : (p * 3 + cwm) * 3 + q;
            wk[k9][cw] = valid ? *reinterpret_cast<const f32x2*>(w27 + tap * CS + c) : 0ull;
        }
        if (FRAMES) rp[k9] = valid ? fbase[p] + (long long)hh * W * CS + c : self;
        else rp[k9] = valid ? x + ((((long long)b * T + tt) * H + hh) * W) * CS + c : self;
    }
    const f32x2 bv = (bias && sign > 0) ? *reinterpret_cast<const f32x2*>(bias + c) : 0ull;
    const long long out_off = FRAMES ? ((long long)t * H + h) * W * CS + c : (self - x);
    float* yo = y + out_off;
    __nv_bfloat16* yb = BF16OUT ? y_bf16 + out_off : nullptr;
    // k9 = 7 is the (dt, dh) = (0, 0) row: B[7] is the residual term x[w] (its weights are always valid)
    f32x2 c0[9], c1[9], c2[9], c3[9], c4[9];
#pragma unroll
    for (int k9 = 0; k9 < 9; ++k9) {
        c0[k9] = 0ull;                                          // x[-1]
        c1[k9] = *reinterpret_cast<const f32x2*>(rp[k9]);
        c2[k9] = (W > 1) ? *reinterpret_cast<const f32x2*>(rp[k9] + (long long)1 * CS) : 0ull;
        c3[k9] = (W > 2) ? *reinterpret_cast<const f32x2*>(rp[k9] + (long long)2 * CS) : 0ull;
        c4[k9] = 0ull;
    }
    int w0 = 0;
#pragma unroll
    for (; w0 + 5 + 3 <= W; w0 += 5) {          // every column w0+3 .. w0+7 is inside the row: unconditional loads
        PEG_STEP(c0, c1, c2, c4, 0, true)
        PEG_STEP(c1, c2, c3, c0, 1, true)
        PEG_STEP(c2, c3, c4, c1, 2, true)
        PEG_STEP(c3, c4, c0, c2, 3, true)
        PEG_STEP(c4, c0, c1, c3, 4, true)
    }
#pragma unroll
    for (; w0 < W; w0 += 5) {                   // tail block(s): guard loads and stores
#define PEG_TAIL(A, B, D, N, WOFF)                                     \
        if (w0 + (WOFF) < W) {                                         \
            if (w0 + (WOFF) + 3 < W) PEG_STEP(A, B, D, N, WOFF, true)  \
            else PEG_STEP(A, B, D, N, WOFF, false)                     \
        }
        PEG_TAIL(c0, c1, c2, c4, 0)
        PEG_TAIL(c1, c2, c3, c0, 1)
        PEG_TAIL(c2, c3, c4, c1, 2)
        PEG_TAIL(c3, c4, c0, c2, 3)
        PEG_TAIL(c4, c0, c1, c3, 4)
#undef PEG_TAIL
    }
}
#undef PEG_STEP

// ---------------------------------------------------------------------------------------------
// Frame plumbing of the occlusion fast path: gather whole frames from two sources, overwrite token rows.
// ---------------------------------------------------------------------------------------------
// out frame f = (src[f] >= 0) ? a[src[f]] : b[-1 - src[f]]      (frame = frame_elems fp32, multiple of 4)
__global__ void __launch_bounds__(256)
frames_gather_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const int* __restrict__ src,
                     long long frame_vec4, float4* __restrict__ out) {
    const int f = blockIdx.y;
    const int v = src[f];
    const float4* s = (v >= 0) ? a + (long long)v * frame_vec4 : b + (long long)(-1 - v) * frame_vec4;
    float4* o = out + (long long)f * frame_vec4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < frame_vec4;
         i += (long long)gridDim.x * blockDim.x)
        o[i] = s[i];
}
// x[rows[r], :] = value[:]
__global__ void rows_fill_kernel(float* __restrict__ x, const int* __restrict__ rows, int n_rows, int C,
                                 const float* __restrict__ value) {
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    float* xr = x + (long long)rows[r] * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) xr[c] = value[c];
}

// ---------------------------------------------------------------------------------------------
// GEGLU
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(const __nv_bfloat16* __restrict__ u, long long R, int F, __nv_bfloat16* __restrict__ h) {
    const int f8 = F >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * f8) return;
    const long long r = idx / f8;
    const int f = (int)(idx % f8) * 8;
    // grouped layout: value column f lives at 64*(f/32) + f%32, its gate 32 columns further
    const long long uo = r * 2 * F + 64 * (f / 32) + (f % 32);
    const uint4 xa = *reinterpret_cast<const uint4*>(u + uo);
    const uint4 ga = *reinterpret_cast<const uint4*>(u + uo + 32);
    const uint32_t xs[4] = {xa.x, xa.y, xa.z, xa.w}, gs[4] = {ga.x, ga.y, ga.z, ga.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 xv = unpack_bf16(xs[i]), gv = unpack_bf16(gs[i]);
        o[i] = pack_bf16(gelu_erf(gv.x) * xv.x, gelu_erf(gv.y) * xv.y);
    }
    *reinterpret_cast<uint4*>(h + r * F + f) = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ dh, long long R, int F,
                 __nv_bfloat16* __restrict__ du) {
    const int f8 = F >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * f8) return;
    const long long r = idx / f8;
    const int f = (int)(idx % f8) * 8;
    const long long uo = r * 2 * F + 64 * (f / 32) + (f % 32);
    const uint4 xa = *reinterpret_cast<const uint4*>(u + uo);
    const uint4 ga = *reinterpret_cast<const uint4*>(u + uo + 32);
    const uint4 da = *reinterpret_cast<const uint4*>(dh + r * F + f);
    const uint32_t xs[4] = {xa.x, xa.y, xa.z, xa.w}, gs[4] = {ga.x, ga.y, ga.z, ga.w}, ds[4] = {da.x, da.y, da.z, da.w};
    uint32_t ox[4], og[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 xv = unpack_bf16(xs[i]), gv = unpack_bf16(gs[i]), dv = unpack_bf16(ds[i]);
        ox[i] = pack_bf16(gelu_erf(gv.x) * dv.x, gelu_erf(gv.y) * dv.y);
        og[i] = pack_bf16(xv.x * gelu_erf_grad(gv.x) * dv.x, xv.y * gelu_erf_grad(gv.y) * dv.y);
    }
    *reinterpret_cast<uint4*>(du + uo) = make_uint4(ox[0], ox[1], ox[2], ox[3]);
    *reinterpret_cast<uint4*>(du + uo + 32) = make_uint4(og[0], og[1], og[2], og[3]);
}

// ---------------------------------------------------------------------------------------------
// Continuous position bias table: one CTA per relative offset, MLP 2 -> dim -> dim -> heads.
// ---------------------------------------------------------------------------------------------
__global__ void cpb_table_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                 const float* __restrict__ w1, const float* __restrict__ b1,
                                 const float* __restrict__ w2, const float* __restrict__ b2, int dim, int heads,
                                 int H, int W, float* __restrict__ table) {
    extern __shared__ float sm[];
    float* h1 = sm;
    float* h2 = sm + dim;
    const int nW = 2 * W - 1;
    const int off = blockIdx.x;
    const int dh = off / nW - (H - 1), dw = off % nW - (W - 1);
    auto logd = [](int d) { return d == 0 ? 0.f : (d > 0 ? 1.f : -1.f) * logf(fabsf((float)d) + 1.f); };
    const float p0 = logd(dh), p1 = logd(dw);
    for (int j = threadIdx.x; j < dim; j += blockDim.x) {
        const float v = w0[j * 2] * p0 + w0[j * 2 + 1] * p1 + b0[j];
        h1[j] = v > 0.f ? v : 0.1f * v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < dim; j += blockDim.x) {
        float v = b1[j];
        for (int k = 0; k < dim; ++k) v += w1[(long long)j * dim + k] * h1[k];
        h2[j] = v > 0.f ? v : 0.1f * v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int hd = warp; hd < heads; hd += blockDim.x >> 5) {
        float v = 0.f;
        for (int k = lane; k < dim; k += 32) v += w2[(long long)hd * dim + k] * h2[k];
        v = warp_sum(v);
        if (lane == 0) table[(long long)hd * (2 * H - 1) * nW + off] = v + b2[hd];
    }
}

}  // namespace ctc

using namespace ctc;

extern "C" int ctc_layernorm_fwd(const float* x, int R, int C, const float* gamma, const float* beta, float eps,
                                 void* y_bf16, float* y_f32, void* xraw_bf16, void* stream) {
    CTC_REQUIRE(C % 4 == 0 && R > 0, "layernorm_fwd: C=%d must be a multiple of 4, R=%d > 0", C, R);
    layernorm_fwd_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        x, R, C, gamma, beta, eps, (__nv_bfloat16*)y_bf16, y_f32, (__nv_bfloat16*)xraw_bf16);
    CTC_LAUNCH_CHECK();
    return 0;
}

static int layernorm_bwd_launch(const float* dy, const __nv_bfloat16* dy16, const float* x, int R, int C,
                                const float* gamma, float eps, float* out, int accumulate, void* out_bf16, void* stream) {
    CTC_REQUIRE(C % 4 == 0 && R > 0, "layernorm_bwd: C=%d must be a multiple of 4, R=%d > 0", C, R);
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* ob = (__nv_bfloat16*)out_bf16;
    const bool al = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                      reinterpret_cast<uintptr_t>(gamma)) & 15) == 0 && (reinterpret_cast<uintptr_t>(ob) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(dy16) & 7) == 0;
    if (C == 512 && al) {
        const int grid = (R + 7) / 8;
#define CTC_LNB(ACC, BF, D16) layernorm_bwd_c512_kernel<ACC, BF, D16><<<grid, 256, 0, st>>>(dy, dy16, x, R, gamma, eps, out, ob)
        if (dy16) {
            if (accumulate) { if (ob) CTC_LNB(true, true, true); else CTC_LNB(true, false, true); }
            else { if (ob) CTC_LNB(false, true, true); else CTC_LNB(false, false, true); }
        } else {
            if (accumulate) { if (ob) CTC_LNB(true, true, false); else CTC_LNB(true, false, false); }
            else { if (ob) CTC_LNB(false, true, false); else CTC_LNB(false, false, false); }
        }
#undef CTC_LNB
    } else {
        layernorm_bwd_kernel<<<(R + 7) / 8, 256, 0, st>>>(dy, dy16, x, R, C, gamma, eps, out, accumulate, ob);
    }
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_layernorm_bwd(const float* dy, const float* x, int R, int C, const float* gamma, float eps,
                                 float* out, int accumulate, void* out_bf16, void* stream) {
    return layernorm_bwd_launch(dy, nullptr, x, R, C, gamma, eps, out, accumulate, out_bf16, stream);
}

extern "C" int ctc_layernorm_bwd_bf16(const void* dy_bf16, const float* x, int R, int C, const float* gamma, float eps,
                                      float* out, int accumulate, void* out_bf16, void* stream) {
    CTC_REQUIRE(dy_bf16 != nullptr && C % 4 == 0, "layernorm_bwd_bf16: dy missing or C=%d not a multiple of 4", C);
    return layernorm_bwd_launch(nullptr, (const __nv_bfloat16*)dy_bf16, x, R, C, gamma, eps, out, accumulate, out_bf16, stream);
}

extern "C" int ctc_peg(const float* x, int B, int T, int H, int W, int C, const float* w27, const float* bias,
                       int mode, int transpose, float* y, void* y_bf16, void* stream) {
    CTC_REQUIRE(x != y, "peg: in-place stencil is not supported");
    CTC_REQUIRE(C % 2 == 0, "peg: C=%d must be even (threads own channel pairs)", C);
    CTC_REQUIRE(mode == CTC_MODE_SPATIAL || (T == H && H == W),
                "peg: temporal mode reinterprets (h,w,t) as (t,h,w) and needs T==H==W (got %d,%d,%d)", T, H, W);
    {   // bulk-asynchronous halo-tile kernel (peg_tma.cu) whenever the geometry allows it
        const int e = peg_tma_launch(x, B, T, H, W, C, w27, bias, mode, transpose, y, y_bf16, (cudaStream_t)stream);
        if (e >= 0) return e;
    }
    const int tiles_t = (T + PEG_TT - 1) / PEG_TT;
    CTC_REQUIRE((long long)B * tiles_t <= 65535, "peg: batch %d too large for one launch", B);
    dim3 grid((C + 63) / 64, (H + PEG_TH - 1) / PEG_TH, B * tiles_t);
    const int threads = PEG_TT * PEG_TH * 32;
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* yb = (__nv_bfloat16*)y_bf16;
#define PEG_LAUNCH(CC, CW, SG, BF) \
    peg_kernel<CC, CW, SG, BF, false><<<grid, threads, 0, st>>>(x, B, T, H, W, C, w27, bias, mode, y, yb, nullptr, nullptr)
    if (C == 512 && W == 24) {     // the CTViT geometry: every offset of the unrolled walk is an immediate
        if (!transpose) { if (yb) PEG_LAUNCH(512, 24, 1, true); else PEG_LAUNCH(512, 24, 1, false); }
        else            { if (yb) PEG_LAUNCH(512, 24, -1, true); else PEG_LAUNCH(512, 24, -1, false); }
    } else {
        if (!transpose) { if (yb) PEG_LAUNCH(0, 0, 1, true); else PEG_LAUNCH(0, 0, 1, false); }
        else            { if (yb) PEG_LAUNCH(0, 0, -1, true); else PEG_LAUNCH(0, 0, -1, false); }
    }
#undef PEG_LAUNCH
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_peg_frames(const float* x_changed, const float* x_base, const int* frame_src, int F, int H, int W,
                              int C, const float* w27, const float* bias, float* y, void* stream) {
    CTC_REQUIRE(F > 0 && C % 2 == 0, "peg_frames: F=%d must be positive and C=%d even", F, C);
    CTC_REQUIRE(y != x_changed && y != x_base, "peg_frames: in-place stencil is not supported");
    const int tiles_t = (F + PEG_TT - 1) / PEG_TT;
    CTC_REQUIRE(tiles_t <= 65535, "peg_frames: %d frames are too many for one launch", F);
    dim3 grid((C + 63) / 64, (H + PEG_TH - 1) / PEG_TH, tiles_t);
    const int threads = PEG_TT * PEG_TH * 32;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 512 && W == 24)
        peg_kernel<512, 24, 1, false, true><<<grid, threads, 0, st>>>(x_changed, 1, F, H, W, C, w27, bias,
                                                                       CTC_MODE_SPATIAL, y, nullptr, x_base, frame_src);
    else
        peg_kernel<0, 0, 1, false, true><<<grid, threads, 0, st>>>(x_changed, 1, F, H, W, C, w27, bias,
                                                                   CTC_MODE_SPATIAL, y, nullptr, x_base, frame_src);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_frames_gather(const float* a, const float* b, const int* src, int F, int64_t frame_elems,
                                 float* out, void* stream) {
    CTC_REQUIRE(F > 0 && frame_elems > 0 && frame_elems % 4 == 0, "frames_gather: F=%d, frame_elems=%lld (multiple of 4)",
                F, (long long)frame_elems);
    CTC_REQUIRE(F <= 65535, "frames_gather: %d frames are too many for one launch", F);
    const long long v4 = frame_elems / 4;
    const int bx = (int)((v4 + 255) / 256 < 64 ? (v4 + 255) / 256 : 64);
    frames_gather_kernel<<<dim3(bx, F), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), src, v4, reinterpret_cast<float4*>(out));
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_rows_fill(float* x, const int* rows, int n_rows, int C, const float* value, void* stream) {
    CTC_REQUIRE(n_rows > 0 && C > 0, "rows_fill: n_rows=%d, C=%d", n_rows, C);
    rows_fill_kernel<<<n_rows, 128, 0, (cudaStream_t)stream>>>(x, rows, n_rows, C, value);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_geglu_fwd(const void* u, int R, int F, void* h, void* stream) {
    CTC_REQUIRE(F % 32 == 0, "geglu: F=%d must be a multiple of 32 (grouped [32 value | 32 gate] layout)", F);
    const long long total = (long long)R * (F / 8);
    geglu_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)u, R, F, (__nv_bfloat16*)h);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_geglu_bwd(const void* u, const void* dh, int R, int F, void* du, void* stream) {
    CTC_REQUIRE(F % 32 == 0, "geglu: F=%d must be a multiple of 32 (grouped [32 value | 32 gate] layout)", F);
    const long long total = (long long)R * (F / 8);
    geglu_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)u, (const __nv_bfloat16*)dh, R, F, (__nv_bfloat16*)du);
    CTC_LAUNCH_CHECK();
    return 0;
}

extern "C" int ctc_cpb_table(const float* w0, const float* b0, const float* w1, const float* b1, const float* w2,
                             const float* b2, int dim, int heads, int H, int W, float* table, void* stream) {
    const int n_off = (2 * H - 1) * (2 * W - 1);
    cpb_table_kernel<<<n_off, 256, 2 * dim * sizeof(float), (cudaStream_t)stream>>>(w0, b0, w1, b1, w2, b2, dim,
                                                                                    heads, H, W, table);
    CTC_LAUNCH_CHECK();
    return 0;
}
