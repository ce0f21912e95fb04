// CT volume preprocessing in front of the patch embedding (reference: src/utils/preprocess.py:84-151, model_type
// "ctclip"): HU rescale (slope * x + intercept), trilinear resample to the target spacing
// (F.interpolate(size=..., mode='trilinear', align_corners=False)), clamp to [-1000, 1000], / 1000, centre crop /
// symmetric pad with -1 to the model's input box — ONE kernel, one read of the raw scan and one write of the
// [D, H, W] fp32 volume the patch embedding consumes (the reference materialises five intermediate copies on the CPU).
//
// The raw scan is addressed through explicit element strides for its logical (H0, W0, D0) axes, so the NIfTI file
// order (first axis fastest) is read in place.  HBM-bound: a CTA owns a 32 (h) x 32 (w) output tile of one output
// slice; values are computed with the thread index running along the raw scan's fastest axis (h) and transposed
// through shared memory so that the output rows (w fastest) are written coalesced.
#include "common.cuh"
#include "ctc_internal.h"

namespace ctc {

struct PreGeom {
    int H0, W0, D0;               // raw logical sizes
    long long sH, sW, sD;         // raw element strides
    int Hn, Wn, Dn;               // resampled sizes
    int offH, offW, offD;         // resampled index = output index + off (crop: +start, pad: -pad_before)
    int H, W, D;                  // output sizes
    float slope, intercept, pad_value;
};

// src = max(0, (dst + 0.5) * in/out - 0.5)   (align_corners=False, size given: scale = in / out in fp32)
CTC_DEVINL void pre_coord(int dst, int in, int out, int& i0, int& i1, float& w1) {
    float s = ((float)dst + 0.5f) * ((float)in / (float)out) - 0.5f;
    s = fmaxf(s, 0.f);
    i0 = min((int)s, in - 1);
    i1 = min(i0 + 1, in - 1);
    w1 = s - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256)
preprocess_ct_kernel(const T* __restrict__ raw, const PreGeom g, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int d = blockIdx.z;
    const int h_base = blockIdx.y * 32, w_base = blockIdx.x * 32;
    const int zi = d + g.offD;
    const bool z_in = zi >= 0 && zi < g.Dn;
    int z0 = 0, z1 = 0; float wz = 0.f;
    if (z_in) pre_coord(zi, g.D0, g.Dn, z0, z1, wz);
    const int th = threadIdx.x & 31;                  // along h (the raw file's fastest axis)
    for (int tw = threadIdx.x >> 5; tw < 32; tw += 8) {
        const int h = h_base + th, w = w_base + tw;
        float v = g.pad_value;
        const int yi = h + g.offH, xi = w + g.offW;
        if (z_in && h < g.H && w < g.W && yi >= 0 && yi < g.Hn && xi >= 0 && xi < g.Wn) {
            int y0, y1, x0, x1; float wy, wx;
            pre_coord(yi, g.H0, g.Hn, y0, y1, wy);
            pre_coord(xi, g.W0, g.Wn, x0, x1, wx);
            // HU transform of the eight corners exactly as the reference does it (fp32 multiply, then add)
            auto hu = [&](int y, int x, int z) {
                const float r = (float)raw[y * g.sH + x * g.sW + z * g.sD];
                return __fadd_rn(__fmul_rn(g.slope, r), g.intercept);
            };
            // separable evaluation, innermost (w) axis first, as ATen's generic N-d linear kernel does
            const float a00 = (1.f - wx) * hu(y0, x0, z0) + wx * hu(y0, x1, z0);
            const float a01 = (1.f - wx) * hu(y1, x0, z0) + wx * hu(y1, x1, z0);
            const float a10 = (1.f - wx) * hu(y0, x0, z1) + wx * hu(y0, x1, z1);
            const float a11 = (1.f - wx) * hu(y1, x0, z1) + wx * hu(y1, x1, z1);
            const float b0 = (1.f - wy) * a00 + wy * a01;
            const float b1 = (1.f - wy) * a10 + wy * a11;
            v = (1.f - wz) * b0 + wz * b1;
            v = fminf(fmaxf(v, -1000.f), 1000.f) / 1000.f;
        }
        tile[tw][th] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31;                  // along w (the output's fastest axis)
    for (int ty = threadIdx.x >> 5; ty < 32; ty += 8) {
        const int h = h_base + ty, w = w_base + tx;
        if (h < g.H && w < g.W) out[((long long)d * g.H + h) * g.W + w] = tile[tx][ty];
    }
}

static void axis_offset(int n, int target, int& off) {
    // crop_and_pad (preprocess.py:39-82): larger -> centre crop from (n - target) // 2; smaller -> pad_before = total // 2
    off = (n > target) ? (n - target) / 2 : -((target - n) / 2);
}

}  // namespace ctc

using namespace ctc;

extern "C" int ctc_preprocess_ct(const void* raw, int raw_dtype, int H0, int W0, int D0, int64_t sH, int64_t sW,
                                 int64_t sD, float slope, float intercept, double z_spacing, double xy_spacing,
                                 double target_z, double target_xy, int D, int H, int W, float pad_value, float* out,
                                 int* resampled_dhw, void* stream) {
    CTC_REQUIRE(H0 > 0 && W0 > 0 && D0 > 0 && D > 0 && H > 0 && W > 0, "preprocess: empty volume");
    CTC_REQUIRE(z_spacing > 0 && xy_spacing > 0 && target_z > 0 && target_xy > 0, "preprocess: spacings must be positive");
    PreGeom g{};
    g.H0 = H0; g.W0 = W0; g.D0 = D0; g.sH = sH; g.sW = sW; g.sD = sD;
    // resize_array (preprocess.py:20-37): new_shape[i] = int(original_shape[i] * (current_spacing[i] / target_spacing[i]))
    g.Dn = (int)((double)D0 * (z_spacing / target_z));
    g.Hn = (int)((double)H0 * (xy_spacing / target_xy));
    g.Wn = (int)((double)W0 * (xy_spacing / target_xy));
    CTC_REQUIRE(g.Dn > 0 && g.Hn > 0 && g.Wn > 0, "preprocess: resampled volume is empty (%d, %d, %d)", g.Dn, g.Hn, g.Wn);
    axis_offset(g.Hn, H, g.offH); axis_offset(g.Wn, W, g.offW); axis_offset(g.Dn, D, g.offD);
    g.H = H; g.W = W; g.D = D; g.slope = slope; g.intercept = intercept; g.pad_value = pad_value;
    if (resampled_dhw) { resampled_dhw[0] = g.Dn; resampled_dhw[1] = g.Hn; resampled_dhw[2] = g.Wn; }
    CTC_REQUIRE(D <= 65535, "preprocess: output depth %d too large for one launch", D);
    dim3 grid((W + 31) / 32, (H + 31) / 32, D);
    cudaStream_t st = (cudaStream_t)stream;
    switch (raw_dtype) {
        case 0: preprocess_ct_kernel<float><<<grid, 256, 0, st>>>((const float*)raw, g, out); break;
        case 1: preprocess_ct_kernel<short><<<grid, 256, 0, st>>>((const short*)raw, g, out); break;
        case 2: preprocess_ct_kernel<double><<<grid, 256, 0, st>>>((const double*)raw, g, out); break;
        default: CTC_REQUIRE(false, "preprocess: raw_dtype %d (0 = float32, 1 = int16, 2 = float64)", raw_dtype);
    }
    CTC_LAUNCH_CHECK();
    return 0;
}
