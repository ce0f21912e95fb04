/* ctclip_b200 — C ABI of the B200-native CT-CLIP-UT attribution hot path.
 *
 * The reference (injardav/CT-CLIP-UT) is pure Python/PyTorch and has NO plugin / FFI / operator
 * registry: its "boundary" is the Python module surface `models.ctclip.CTCLIP`,
 * `utils.ctvit.CTViT`, `utils.visualizations.Visualizations`.  This header is therefore the
 * boundary a maintainer would bind *underneath* those modules (ctypes stub: INTEGRATION.md);
 * each entry point cites the reference code it replaces (paths relative to the upstream repo).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory (torch tensors); the library
 *     allocates nothing persistent;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), returns 0 on
 *     success, non-zero on error with the message available from ctc_last_error();
 *   - sm_100a only: there is no CPU path and no other-architecture path;
 *   - token-space matrices are row-major [R, C] with R = B*T*H*W rows in canonical (b,t,h,w)
 *     order; the "temporal" view of the reference ('(b h w) t d', ctvit.py:99) is never
 *     materialised — kernels that care about sequence order take a mode flag instead;
 *   - bf16 matrices are `uint16_t`-sized elements (torch.bfloat16), fp32 otherwise.
 */
#ifndef CTCLIP_B200_H_
#define CTCLIP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTC_VERSION 100

/* GEMM epilogues */
enum { CTC_EPI_BF16 = 0, CTC_EPI_F32 = 1, CTC_EPI_ARGMAX = 2, CTC_EPI_GEGLU = 3, CTC_EPI_GEGLU_BWD = 4 };
/* GEMM implementation: tcgen05 is the product path (CTA pairs / cta_group::2 for the shapes where they measure
 * faster, single CTAs otherwise); TCGEN05_1CTA / TCGEN05_PAIR force one kernel and SIMT is a plain comparator -
 * all three for tests and A/B measurements. */
enum { CTC_GEMM_TCGEN05 = 0, CTC_GEMM_SIMT = 1, CTC_GEMM_TCGEN05_1CTA = 2, CTC_GEMM_TCGEN05_PAIR = 3 };
/* Flag OR-ed into `impl`: the rows of B (= output channels) have been permuted inside every 32-row group with
 * ctc_gemm_row_perm when the weight was packed, which lets the epilogue store straight from the tcgen05.ld.16x256b
 * register layout without a shared-memory staging pass.  Results are bit-identical to the
 * unflagged call on the unpermuted B.  N must be a multiple of 32; fp32 outputs / residual / bias 32-byte aligned. */
enum { CTC_GEMM_BPERM = 0x100 };
/* sequence mode of the factorised transformer (ctvit.py:94-101) */
enum { CTC_MODE_SPATIAL = 0, CTC_MODE_TEMPORAL = 1 };

int ctc_version(void);
const char* ctc_last_error(void);
/* 0 if the current device is compute capability 10.x; error otherwise (no fallback). */
int ctc_device_check(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long ctc_launch_count(void);

/* C[M,N] = A[M,K] * B[N,K]^T, bf16 operands (row strides lda/ldb elements), fp32 accumulate.
 * epi BF16: out bf16 [M,ldc]; F32: out fp32 = acc (+bias[N]) (+resid fp32 [M,ldr], may alias out).
 * epi GEGLU (FeedForward, attention.py:38-49): the N output columns are 64-wide groups [32 value | 32 gate]
 *   (weight rows interleaved by the caller); out = gelu(gate)*value bf16 [M, N/2]; aux (optional, stride ldaux)
 *   receives, in the same grouped layout, the ADJOINT FACTORS the backward pass needs, bf16 [M, N]:
 *   [32 a | 32 b] with a = gelu(gate) = dh/dvalue and b = value * gelu'(gate) = dh/dgate.
 * epi GEGLU_BWD: acc = dh [M,N]; aux = those factors bf16 [M, 2N]; out = du bf16 [M, 2N] = [a*dh | b*dh], the
 *   gradient w.r.t. the pre-activation in the grouped layout (two multiplies per element: fits the epilogue).
 * Replaces every nn.Linear / einsum on the path: attention.py:47,49,142,182; ctvit.py:50. */
int ctc_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldc, int M, int N,
                  int K, int epi, const float* bias, const float* resid, int64_t ldr, void* aux, int64_t ldaux,
                  int impl, void* stream);
/* perm32[a] (host, 32 ints) = the output channel, within its group of 32, whose weights belong in row a of that
 * group of B for CTC_GEMM_BPERM: Bperm[G*32 + a] = B[G*32 + perm32[a]]. */
int ctc_gemm_row_perm(int* perm32);

/* 3-D patchify + LayerNorm(P) (ctvit.py:44-49): volume fp32 [B,1,D,H,W] -> bf16 [B*T*H*W, P]
 * (P = pt*p*p, element order pt,p1,p2).  Fused on load, so perturbed volumes are never
 * materialised: IG interpolation x' = 1 + alpha[b]*(x-1) (visualizations.py:862) when alpha != NULL,
 * then occlusion cube fill (visualizations.py:380-381) when occl != NULL: occl[b] = {d0,h0,w0,pd,ph,pw}
 * (pd<=0 disables) with fill value occl_value.  vol_batch_stride = elements between successive volumes
 * (0 = the same volume for every b: windows / alpha steps of ONE volume). */
int ctc_patchify_ln_fwd(const float* volume, int64_t vol_batch_stride, int B, int D, int H, int W, int pt, int p,
                        const float* gamma, const float* beta, float eps, const float* alpha, const int* occl,
                        float occl_value, void* out_bf16, void* stream);
/* Backward of the above w.r.t. the (perturbed) voxels: dY bf16 [R,P] -> grad fp32 [B,1,D,H,W]
 * (or, with sum_over_batch=1, accumulated (+=) into ONE [D,H,W] buffer scaled by wscale: the
 * integrated-gradients running sum of visualizations.py:872,878). */
int ctc_patchify_ln_bwd(const float* volume, int64_t vol_batch_stride, int B, int D, int H, int W, int pt, int p,
                        const float* gamma, float eps, const float* alpha, const void* dy_bf16, float* grad,
                        int sum_over_batch, float wscale, void* stream);

/* LayerNorm over the last dim (attention.py:27-34, 46; ctvit.py:51).  x fp32 [R,C].
 * Any of y_bf16 / y_f32 / xraw_bf16 (= bf16(x), the un-normalised k/v input of attention.py:138) may be NULL. */
int ctc_layernorm_fwd(const float* x, int R, int C, const float* gamma, const float* beta, float eps,
                      void* y_bf16, float* y_f32, void* xraw_bf16, void* stream);
/* dx = LN'(x)·(dy*gamma); out[r] = (accumulate ? out[r] : 0) + dx; optional bf16 copy of out. */
int ctc_layernorm_bwd(const float* dy, const float* x, int R, int C, const float* gamma, float eps, float* out,
                      int accumulate, void* out_bf16, void* stream);
/* The same with the incoming gradient in bf16 (as a GEMM's bf16 epilogue leaves it: half the bytes of the pass). */
int ctc_layernorm_bwd_bf16(const void* dy_bf16, const float* x, int R, int C, const float* gamma, float eps, float* out,
                           int accumulate, void* out_bf16, void* stream);

/* PEG (attention.py:55-83): y = x + dwconv3d(pad(x)) + bias, causal padding (2,0) on the first
 * grid axis.  mode TEMPORAL reproduces the reference's axis scramble: the flat
 * '(b h w) t' tensor is reinterpreted as (t',h',w') = (h,w,t).  w27 is the depthwise weight
 * transposed to [27, C].  transpose=1 computes the adjoint (input gradient; bias ignored). */
int ctc_peg(const float* x, int B, int T, int H, int W, int C, const float* w27, const float* bias, int mode,
            int transpose, float* y, void* y_bf16, void* stream);

/* Occlusion fast path (visualizations.py:335-392 evaluates one full forward per window; a window only
 * changes the tokens of its cube, and the spatial transformer's only cross-frame operator is the CAUSAL
 * depthwise stencil, so the changed frame set grows by two frames per layer and every other frame equals
 * the cached baseline).  ctc_peg_frames is ctc_peg (forward, SPATIAL) over a compact list of F output
 * frames: frame_src int32 [F,3] names the source frame for dt = -2,-1,0: v >= 0 -> frame v of x_changed,
 * v < 0 -> frame (-1 - v) of x_base, INT_MIN -> causal zero padding.  y fp32 [F, H, W, C].
 * ctc_frames_gather: out[f] = src[f] >= 0 ? a[src[f]] : b[-1 - src[f]] (frames of frame_elems fp32).
 * ctc_rows_fill: x[rows[r], :] = value[:] (the embedding of a fully occluded patch). */
int ctc_peg_frames(const float* x_changed, const float* x_base, const int* frame_src, int F, int H, int W, int C,
                   const float* w27, const float* bias, float* y, void* stream);
int ctc_frames_gather(const float* a, const float* b, const int* src, int F, int64_t frame_elems, float* out,
                      void* stream);
int ctc_rows_fill(float* x, const int* rows, int n_rows, int C, const float* value, void* stream);

/* Cosine-similarity attention core (attention.py:144-180): per (sequence, head)
 * softmax(scale * l2norm(q)*q_scale . l2norm(k)*k_scale + bias) v.  q bf16 [R, heads*32] (stride ldq),
 * k/v bf16 (stride ldkv; v = k + heads*32 columns).  SPATIAL: sequences are the (b,t) slices of
 * H*W tokens, bias_table fp32 [heads, (2H-1)*(2W-1)] is the relative-position bias (attention.py:230-277)
 * indexed by (dh+H-1)*(2W-1)+(dw+W-1); TEMPORAL: sequences are the T tokens of each (b,h,w), no bias.
 * Writes o bf16 [R, heads*32] and lse fp32 [R, heads] (natural-log row log-sum-exp). */
int ctc_attention_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int B, int T, int H,
                      int W, int heads, const float* q_scale, const float* k_scale, float scale,
                      const float* bias_table, int mode, void* o, float* lse, void* stream);
/* The same forward for the SPATIAL sequences on tcgen05 / TMEM (S = Q K^T and O = P V as tcgen05.mma with P fed
 * from tensor memory).  score_bound must bound every attention score: scale*max|q_scale|*max|k_scale| + max|bias|
 * (ctc_attention_score_bound writes it to a device float); it replaces the running row maximum of the softmax.
 * Needs H*W % 64 == 0, W % 8 == 0, H*W <= 640, 0 < score_bound < 43; otherwise use ctc_attention_fwd. */
int ctc_attention_fwd_tc(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int B, int T, int H,
                         int W, int heads, const float* q_scale, const float* k_scale, float scale,
                         const float* bias_table, float score_bound, void* o, float* lse, void* stream);
/* Spatial forward kernel: compute every other exponential of the softmax with an FMA-pipe polynomial instead of
 * MUFU.EX2 (the FA4 split).  Off by default - it measured 21 % slower here; returns the previous setting. */
int ctc_attention_set_exp2_poly(int on);
/* Which kernels ctc_attention_bwd runs for spatial sequences: mode 2 (default) = dQ, dK and dV in ONE pass on
 * tcgen05 / TMEM (attention_tc_bwd.cu; 1 010 us against 1 245 us for the two mma.sync kernels at batch 8; same result to
 * bf16 rounding, bit-reproducible) where the geometry allows (bias table, n % 32 == 0, 64 <= n <= 640, W % 8 == 0),
 * mode 0 = the two mma.sync kernels (every other geometry always takes them), mode 1 = tcgen05 dQ + mma.sync dK/dV
 * (4 % slower than mode 0).  Returns the previous setting (NOT a status code). */
int ctc_attention_set_tc_bwd(int mode);
int ctc_attention_score_bound(const float* q_scale, const float* k_scale, float scale, const float* bias_table,
                              int heads, int H, int W, float* bound_dev, void* stream);
/* Input gradients of the above (through softmax, l2norm and q/k scales). dq bf16 [R,heads*32] (lddq),
 * dk/dv bf16 (lddkv). delta_ws fp32 [R, heads] scratch. */
int ctc_attention_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* o,
                      const void* d_o, const float* lse, int B, int T, int H, int W, int heads,
                      const float* q_scale, const float* k_scale, float scale, const float* bias_table, int mode,
                      void* dq, int64_t lddq, void* dk, void* dv, int64_t lddkv, float* delta_ws, void* stream);
/* Materialise the attention probabilities the reference's Attention.forward returns
 * (attention.py:174-175): probs fp32 [n_seq, heads, n, n]. */
int ctc_attention_probs(const void* q, int64_t ldq, const void* k, int64_t ldkv, const float* lse, int B, int T,
                        int H, int W, int heads, const float* q_scale, const float* k_scale, float scale,
                        const float* bias_table, int mode, float* probs, void* stream);
/* What the attention-map methods actually consume of those probabilities, emitted WITHOUT materialising them
 * (the reference keeps attn [b,heads,n,n] of every layer alive through hooks, visualizations.py:153-186):
 *   fused   fp32 [n_seq, n, n]      head fusion of visualizations.py:722-725: fusion 0 = mean, 1 = max   (or NULL)
 *   colmean fp32 [n_seq, heads, n]  per head, the mean over the QUERY axis (visualizations.py:666,671)  (or NULL)
 * colpart_ws fp32 [n_seq, heads, ceil(n/32), n] scratch (needed with colmean).  Deterministic (no atomics). */
int ctc_attention_fused_probs(const void* q, int64_t ldq, const void* k, int64_t ldkv, const float* lse, int B, int T,
                              int H, int W, int heads, const float* q_scale, const float* k_scale, float scale,
                              const float* bias_table, int mode, int fusion, float* fused, float* colmean,
                              float* colpart_ws, void* stream);

/* Stand-alone GEGLU (attention.py:38-41) on the grouped layout of the fused epilogues: u bf16 [R, 2*F] in
 * 64-wide groups [32 value | 32 gate] -> h = gelu(gate)*value bf16 [R,F]; the product path uses the fused
 * GEMM epilogues instead, these remain as comparators. */
int ctc_geglu_fwd(const void* u, int R, int F, void* h, void* stream);
int ctc_geglu_bwd(const void* u, const void* dh, int R, int F, void* du, void* stream);

/* Continuous position bias table (attention.py:259-277), evaluated once for the (2H-1)*(2W-1)
 * distinct offsets instead of the (HW)^2 pairs the reference recomputes every forward. */
int ctc_cpb_table(const float* w0, const float* b0, const float* w1, const float* b1, const float* w2,
                  const float* b2, int dim, int heads, int H, int W, float* table, void* stream);

/* VQ nearest-code search (ctvit.py:117-118; vector_quantize_pytorch cosine codebook):
 * cand_val/cand_idx [R, ctc_vq_num_candidates(K)] scratch = the two best groups of 4 consecutive codes of every
 * 128-code slice of the bf16 score GEMM (ctc_gemm_bf16 epi ARGMAX is run internally); every code of a group within
 * a rounding margin of the best is re-scored in fp32 against the fp32 codebook, so the arg-max is exact w.r.t. x.
 * ind int32 [R]. */
int ctc_vq_num_candidates(int K);
int ctc_vq_argmax(const float* x, const void* x_bf16, int R, int C, const float* codebook,
                  const void* codebook_bf16, int K, float* cand_val, int* cand_idx, int* ind, void* stream);
/* pooled[b,hw,c] = mean_t E[ind[b,t,hw]][c] (ctclip.py:111) fp32 (+bf16 copy); tokens fp32 [R,C] optional. */
int ctc_vq_gather_pool(const int* ind, const float* codebook, int B, int T, int HW, int C, float* pooled,
                       void* pooled_bf16, float* tokens, void* stream);
/* Straight-through backward to the pre-VQ activations: g[b,t,hw,:] = dpooled[b,hw,:]/T (+ dtokens if given);
 * grad_mode 0 (ste_l2norm): dx = (g - xh (xh.g))/|x| ; 1 (ste_raw): dx = g. */
int ctc_vq_bwd(const float* dpooled, const float* dtokens, const float* x, int B, int T, int HW, int C,
               int grad_mode, float* dx, void* stream);

/* latent[b,:] = pooled[b,:] @ Wv^T (ctclip.py:116), Wv bf16 [NL, L]; partial fp32 [chunks, B, NL] scratch.
 * wv_lo_bf16 (optional, same shape): the bf16 rounding residual of the weight, Wv = hi + lo with lo = bf16(Wv - hi),
 * for a ~16-bit-mantissa product (the logit is a 294 912-long dot product whose differences are the signal). */
int ctc_latent_proj(const float* pooled, const void* wv_bf16, const void* wv_lo_bf16, int B, int64_t L, int NL,
                    float* partial, int n_chunks, float* latent, void* stream);
/* dpooled[b,:] = dlatent[b,:] @ Wv (fp32 [B, L]) */
int ctc_latent_proj_bwd(const float* dlatent, const void* wv_bf16, int B, int64_t L, int NL, float* dpooled,
                        void* stream);
/* text_latent = l2norm(Wt @ e) (ctclip.py:115,119); e fp32 [Bt, DT], Wt fp32 [NL, DT] */
int ctc_text_latent(const float* e, const float* wt, int Bt, int DT, int NL, float* out, void* stream);
/* sim[i,j] = l2norm(latent_i).text_j * temp (ctclip.py:120,127); image_latents (normalised) optional;
 * dlatent (optional) = d sim[i, i mod Bt] / d latent_i. */
int ctc_latent_sim(const float* latent, const float* text_latents, int B, int Bt, int NL, float temp, float* sim,
                   float* image_latents, float* dlatent, void* stream);

/* General adjoint of ctc_latent_sim: dlatent[i,:] = sum_j gsim[i,j] * d sim[i,j] / d latent_i (gsim fp32 [B,Bt]). */
int ctc_latent_sim_bwd(const float* latent, const float* text_latents, const float* gsim, int B, int Bt, int NL,
                       float temp, float* dlatent, void* stream);

/* Attention-rollout reductions (visualizations.py:707-743 as used at :800-841).
 * spatial: probs fp32 [n_slices, heads, n, n] -> out[n_slices, n] = colsum(rownorm(rownorm(mean_h P)+I)).
 * temporal: probs of all L layers, each [n_tok, heads, T, T] (layer stride = n_tok*heads*T*T) -> out[n_tok, T]. */
int ctc_rollout_spatial(const float* probs, int n_slices, int heads, int n, float* out, void* stream);
int ctc_rollout_temporal(const float* probs, int n_layers, int n_tok, int heads, int T, float* out, void* stream);
/* The generic Visualizations.attention_rollout (visualizations.py:707-743), one layer at a time:
 * ctc_rollout_fuse: attn fp32 [heads, n, n] -> A fp32 [n, n]: head fusion (0 mean / 1 max); keep only the k_keep
 *   largest weights of each row (k_keep = n - int(n * discard_ratio); k_keep = n keeps all); A /= rowsum + 1e-8;
 *   with use_residual A += I, A /= rowsum.
 * ctc_matmul_f32: C = A @ B, fp32 [n, n] row-major (result = attn @ result, :741). */
int ctc_rollout_fuse(const float* attn, int heads, int n, int fusion, int k_keep, int use_residual, float* out,
                     void* stream);
int ctc_matmul_f32(const float* A, const float* B, int n, float* C, void* stream);
/* raw attention (visualizations.py:666,671): out[s, h, j] = mean_i P[s,h,i,j] */
int ctc_attn_colmean(const float* probs, int n_seq, int heads, int n, float* out, void* stream);

/* Grad-CAM (visualizations.py:933-991): w[c] = mean_r grad[r,c]; cam[r] = relu(sum_c (fa[r,c]-fb[r,c]) w[c])
 * (feature = difference of two saved residual streams; fb may be NULL). */
int ctc_colmean_ws_floats(int R, int C);
int ctc_colmean(const float* g, int R, int C, float* w, float* ws, void* stream);
int ctc_gradcam(const float* fa, const float* fb, const float* w, int R, int C, float* cam, void* stream);

/* Trilinear upsample, align_corners=False (visualizations.py:289-293), optional fused
 * np.rot90(k=-1, axes=(1,2)) (e.g. :816): in fp32 [d,h,w] -> out fp32 [D,H,W] (or [D,W,H] rotated). */
int ctc_upsample_trilinear(const float* in, int d, int h, int w, float* out, int D, int H, int W, int rot90,
                           void* stream);
/* Integrated-gradients finalisation, first half (visualizations.py:878-879):
 * ig = relu((x - 1) * gsum * inv_steps), plus global min/max into mm[2] (mm pre-set to {+inf,-inf}). */
int ctc_ig_combine(const float* volume, const float* gsum, int64_t n, float inv_steps, float* ig, float* mm,
                   void* stream);

/* dst[i] = (accumulate ? dst[i] : 0) + scale * sum_b src[b, i], batch rows added in order: the running sum over
 * the alpha steps of integrated gradients (visualizations.py:872,878) without floating-point atomics. */
int ctc_batch_sum(const float* src, int B, int64_t n, float scale, int accumulate, float* dst, void* stream);

/* CT preprocessing in front of the patch embedding (src/utils/preprocess.py:84-151, model_type "ctclip"), fused:
 * HU rescale slope*x+intercept, trilinear resample (align_corners=False) from (z_spacing, xy_spacing, xy_spacing)
 * to (target_z, target_xy, target_xy), clamp [-1000,1000] / 1000, centre crop / symmetric pad with pad_value to
 * [D,H,W].  raw: device array of logical shape [H0,W0,D0] (the reference's NIfTI axis order) with element strides
 * (sH,sW,sD); raw_dtype 0 = float32, 1 = int16, 2 = float64.  out fp32 [D,H,W]; resampled_dhw (host, optional)
 * receives the intermediate resampled shape. */
int ctc_preprocess_ct(const void* raw, int raw_dtype, int H0, int W0, int D0, int64_t sH, int64_t sW, int64_t sD,
                      float slope, float intercept, double z_spacing, double xy_spacing, double target_z,
                      double target_xy, int D, int H, int W, float pad_value, float* out, int* resampled_dhw,
                      void* stream);
/* Global min/max of an fp32 array into mm[2] (caller pre-sets {+inf, -inf}). */
int ctc_minmax(const float* x, int64_t n, float* mm, void* stream);
/* Zero-shot scoring (CTClipInference.py:133-145, 171-180): sim fp32 [B, 2P] with the "There is X." logit at
 * column 2j and the "There is no X." logit at 2j+1; out float64 [B, P] = softmax over each pair, first entry. */
int ctc_pair_softmax(const float* sim, int B, int P, double* out, void* stream);
/* flags uint8 [D/pt, H/p, W/p]: 1 iff every voxel of the patch equals `value`.  An occlusion window made only of
 * such patches leaves the volume unchanged: its score is the un-occluded score (visualizations.py:380-390). */
int ctc_patch_is_constant(const float* volume, int D, int H, int W, int pt, int p, float value, unsigned char* flags,
                          void* stream);
/* The reference's normalisations: mode 0 (v-min)/(max+1e-8) [raw attention :674, Grad-CAM :946,
 * IG :882]; mode 1 (v-min)/(max-min+1e-8) [rollout :812, occlusion :414]; mode 2 v/(max+1e-8) [IG :893];
 * optional fused np.rot90(k=-1, axes=(1,2)): in [D,H,W] -> out [D,W,H]. */
int ctc_normalize(const float* in, int D, int H, int W, const float* mm, int mode, int rot90, float* out, void* stream);
/* One 16-bit radix pass over fp32 bit patterns (non-negative data) for the exact device-side
 * np.quantile(., 0.90) of visualizations.py:886: shift 16 = high half, shift 0 = low half of the
 * elements whose high half equals prefix.  hist uint32 [65536] (zeroed by the call). */
int ctc_hist16(const float* x, int64_t n, int shift, unsigned int prefix, unsigned int* hist, void* stream);
/* k-th smallest (0-based) element of a non-negative fp32 array, exact, two radix passes, result written to
 * out_dev: no host round trip.  ws: ctc_kth_value_ws_bytes() bytes of 8-byte aligned scratch. */
int ctc_kth_value_ws_bytes(void);
int ctc_kth_value(const float* x, int64_t n, int64_t k, void* ws, float* out_dev, void* stream);
/* np.quantile's "linear" interpolation between two order statistics in float32: out = lerp(a, b, gamma). */
int ctc_quantile_lerp(const float* a_dev, const float* b_dev, float gamma, float* out_dev, void* stream);
/* Integrated-gradients finalisation, second half (visualizations.py:882-901): normalise, zero below the
 * quantile q, ** 0.05, divide by the new max (inv_m3 = 1/(max+1e-8)), optional rot90. */
int ctc_ig_finalize(const float* ig, int D, int H, int W, float mn, float mx, float q, float inv_m3, int rot90,
                    float* out, void* stream);
/* The same with min/max (mm_dev[2]) and the quantile (q_dev) read from device memory. */
int ctc_ig_finalize_dev(const float* ig, int D, int H, int W, const float* mm_dev, const float* q_dev, int rot90,
                        float* out, void* stream);
/* Occlusion heat map (visualizations.py:366-367, 390-392, 411-413) from per-window importances on the regular
 * window grid [nd,nh,nw] (window i starts at i*stride); inc uint8 marks evaluated windows; heat = sum/count. */
int ctc_occlusion_heatmap(const float* imp, const unsigned char* inc, int nd, int nh, int nw, int pd, int ph, int pw,
                          int sd, int sh, int sw, int D, int H, int W, float* heat, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCLIP_B200_H_ */
