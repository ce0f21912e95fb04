// Head-fused attention probabilities for the attention-map methods (north-star item 2: "optionally emit
// head-averaged attention probabilities for rollout").
//
// The reference keeps `attn [b, heads, n, n]` of every layer alive through forward hooks
// (src/utils/visualizations.py:153-186; 255 MB fp32 per spatial layer and volume) and reduces it on the host:
//   rollout        visualizations.py:720-741  ->  needs the HEAD-FUSED matrix (mean or max over heads)
//   raw attention  visualizations.py:666,671  ->  needs, per head, the mean over the QUERY axis
// This kernel recomputes the probabilities of one (sequence, 32-query block) from the saved q / k / row
// log-sum-exp, head by head, and emits only those two reductions: the fused matrix [n_seq, n, n] (31.9 MB
// per spatial layer) and per-head column sums; the per-head probabilities never reach HBM.  The arithmetic
// of a score is the forward kernel's (bf16-rounded l2-normalised q^ * scale and k^, fp32 accumulation, fp32
// bias), so p = exp2(s - lse) sums to one over a row.  Everything is reduced in a fixed order (no atomics).
// SIMT on purpose: 8.2 GFLOP per spatial layer is ~0.1 ms of FMA throughput and the pass is bound by the
// 32 MB it writes.
#include "attention_common.cuh"

namespace ctc {

static constexpr int FQB = 32;       // query rows per CTA
static constexpr int KSTR = 17;      // k^ row stride in 32-bit words (16 words of bf16 pairs + 1 pad: conflict-free)

__global__ void __launch_bounds__(256)
attn_fused_probs_kernel(const AttnParams p, int fusion_max, float* __restrict__ fused, float* __restrict__ colpart,
                        int n_qb) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint32_t* kt = reinterpret_cast<uint32_t*>(sm);                         // [n][KSTR] bf16 pairs of k^
    float* csum = reinterpret_cast<float*>(kt + (size_t)p.n * KSTR);        // [8 warps][n]
    float* bias = csum + 8 * p.n;                                           // [(2H-1)(2W-1)] (spatial only)
    float* acc = bias + (p.bias_table ? (2 * p.H - 1) * (2 * p.W - 1) : 0); // [FQB][n] fused rows (if requested)
    const int s = blockIdx.x, qb = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = tid >> 3, c = tid & 7;                                    // query row of the block, key lane
    const int i = qb * FQB + r;
    const bool row_ok = i < p.n;
    const int nW = 2 * p.W - 1;
    const int nb = (2 * p.H - 1) * nW;
    const int base_i = row_ok && p.bias_table ? (i / p.W + p.H - 1) * nW + (i % p.W + p.W - 1) : 0;
    const float inv_heads = 1.f / p.heads;

    for (int head = 0; head < p.heads; ++head) {
        __syncthreads();                                                    // previous head's tiles are consumed
        // ---- k^ = bf16(l2norm(k) * k_scale) for all keys of this head (attention.py:155-156)
        for (int j = tid; j < p.n; j += blockDim.x) {
            const uint4* g = reinterpret_cast<const uint4*>(p.k + seq_row(p, s, j) * p.ldkv + head * DH);
            float f[32];
            float ss = 0.f;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint4 q4 = g[v];
                const uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 t = unpack_bf16(w[e]);
                    f[v * 8 + e * 2] = t.x; f[v * 8 + e * 2 + 1] = t.y;
                    ss += t.x * t.x + t.y * t.y;
                }
            }
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
            for (int d = 0; d < 16; ++d)
                kt[j * KSTR + d] = pack_bf16(f[2 * d] * inv * p.k_scale[2 * d], f[2 * d + 1] * inv * p.k_scale[2 * d + 1]);
        }
        if (p.bias_table)
            for (int k = tid; k < nb; k += blockDim.x) bias[k] = p.bias_table[(long long)head * nb + k] * LOG2E;
        // ---- this thread's query row: q^ = bf16(l2norm(q) * q_scale * scale * log2e), held as fp32 registers
        float qv[32];
        float lse2 = 0.f;
        if (row_ok) {
            const uint4* g = reinterpret_cast<const uint4*>(p.q + seq_row(p, s, i) * p.ldq + head * DH);
            float ss = 0.f;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint4 q4 = g[v];
                const uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 t = unpack_bf16(w[e]);
                    qv[v * 8 + e * 2] = t.x; qv[v * 8 + e * 2 + 1] = t.y;
                    ss += t.x * t.x + t.y * t.y;
                }
            }
            const float inv = (p.scale * LOG2E) / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
            for (int d = 0; d < 16; ++d) {
                const float2 t = unpack_bf16(pack_bf16(qv[2 * d] * inv * p.q_scale[2 * d],
                                                       qv[2 * d + 1] * inv * p.q_scale[2 * d + 1]));
                qv[2 * d] = t.x; qv[2 * d + 1] = t.y;
            }
            lse2 = p.lse[seq_row(p, s, i) * p.heads + head] * LOG2E;
        } else {
#pragma unroll
            for (int d = 0; d < 32; ++d) qv[d] = 0.f;
        }
        __syncthreads();
        // ---- probabilities of (row i, keys c, c+8, ...)
        for (int j0 = 0; j0 < p.n; j0 += 8) {
            const int j = j0 + c;
            float pr = 0.f;
            if (row_ok && j < p.n) {
                float sc = p.bias_table ? bias[base_i - ((j / p.W) * nW + (j % p.W))] : 0.f;
                const uint32_t* kr = kt + j * KSTR;
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    const float2 kk = unpack_bf16(kr[d]);
                    sc = fmaf(qv[2 * d], kk.x, sc);
                    sc = fmaf(qv[2 * d + 1], kk.y, sc);
                }
                pr = fast_exp2(sc - lse2);
                if (fused) {
                    float* a = acc + r * p.n + j;
                    if (fusion_max) *a = head == 0 ? pr : fmaxf(*a, pr);
                    else *a = head == 0 ? pr * inv_heads : fmaf(pr, inv_heads, *a);
                }
            }
            if (colpart) {
                // sum over the 4 query rows this warp holds for key j (lanes c, c+8, c+16, c+24)
                float v = pr;
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (lane < 8 && j < p.n) csum[warp * p.n + j] = v;
            }
        }
        if (colpart) {
            __syncthreads();
            for (int j = tid; j < p.n; j += blockDim.x) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += csum[w * p.n + j];         // fixed order
                colpart[(((long long)s * p.heads + head) * n_qb + qb) * p.n + j] = v;
            }
        }
    }
    if (fused) {
        __syncthreads();
        const int rows = min(FQB, p.n - qb * FQB);
        float* dst = fused + ((long long)s * p.n + (long long)qb * FQB) * p.n;
        for (int e = tid; e < rows * p.n; e += blockDim.x) dst[e] = acc[e];
    }
}

// colmean[s, h, j] = (sum over query blocks, in order) / n
__global__ void __launch_bounds__(256)
colpart_reduce_kernel(const float* __restrict__ colpart, int n_qb, int n, long long total, float* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over (s*heads, j)
    if (idx >= total) return;
    const long long sh = idx / n;
    const int j = (int)(idx - sh * n);
    float v = 0.f;
    for (int q = 0; q < n_qb; ++q) v += colpart[(sh * n_qb + q) * n + j];
    out[idx] = v / (float)n;
}

}  // namespace ctc

using namespace ctc;

extern "C" int ctc_attention_fused_probs(const void* q, int64_t ldq, const void* k, int64_t ldkv, const float* lse,
                                         int B, int T, int H, int W, int heads, const float* q_scale,
                                         const float* k_scale, float scale, const float* bias_table, int mode,
                                         int fusion, float* fused, float* colmean, float* colpart_ws, void* stream) {
    AttnParams p{};
    p.bias_table = bias_table;
    if (int e = fill_params(p, B, T, H, W, heads, mode, 8)) return e;
    CTC_REQUIRE(fusion == 0 || fusion == 1, "attention_fused_probs: fusion must be 0 (mean) or 1 (max), got %d", fusion);
    CTC_REQUIRE(fused || colmean, "attention_fused_probs: nothing to emit");
    CTC_REQUIRE(!colmean || colpart_ws, "attention_fused_probs: column means need the [n_seq, heads, ceil(n/32), n] workspace");
    p.q = (const __nv_bfloat16*)q; p.ldq = ldq; p.k = (const __nv_bfloat16*)k; p.v = nullptr; p.ldkv = ldkv;
    p.q_scale = q_scale; p.k_scale = k_scale; p.scale = scale; p.lse = const_cast<float*>(lse);
    const int n_qb = (p.n + FQB - 1) / FQB;
    const size_t nb = bias_table ? (size_t)(2 * H - 1) * (2 * W - 1) : 0;
    const size_t smem = (size_t)p.n * KSTR * 4 + 8 * (size_t)p.n * 4 + nb * 4 + (fused ? (size_t)FQB * p.n * 4 : 0);
    CTC_REQUIRE(smem <= 220 * 1024, "attention_fused_probs: sequence length %d does not fit shared memory", p.n);
    static size_t configured_dev[kMaxDevices] = {};
    
    size_t& configured = configured_dev[current_device()];
    if (smem > configured) {
        CTC_CHECK_CUDA(cudaFuncSetAttribute(attn_fused_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaStream_t st = (cudaStream_t)stream;
    attn_fused_probs_kernel<<<dim3(p.n_seq, n_qb), 256, smem, st>>>(p, fusion, fused, colmean ? colpart_ws : nullptr, n_qb);
    CTC_LAUNCH_CHECK();
    if (colmean) {
        const long long total = (long long)p.n_seq * heads * p.n;
        colpart_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(colpart_ws, n_qb, p.n, total, colmean);
        CTC_LAUNCH_CHECK();
    }
    return 0;
}
