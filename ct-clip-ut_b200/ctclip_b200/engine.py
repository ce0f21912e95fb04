"""Kernel orchestration of the CTViT / CTCLIP image tower: forward and input-gradient backward.

Mirrors, step by step, CTViT.forward (src/utils/ctvit.py:105-125), CTViT.encode (:88-103),
Transformer.forward (src/utils/attention.py:322-336) and CTCLIP.forward (src/models/ctclip.py:99-129),
but every step is one C-ABI call into the sm_100a library (see include/ctclip_b200.h).  torch is
used for buffer allocation only.

Row order: all token-space matrices are [R, C] with R = B*T*H*W in canonical (b, t, h, w) order —
the spatial '(b t) (h w) d' view *is* this order and the temporal '(b h w) t d' view is addressed by
stride inside the attention / PEG kernels (mode flag), so no rearrange is ever materialised.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import EPI_BF16, EPI_F32, MODE_SPATIAL, MODE_TEMPORAL, call, stream_ptr
from .plan import Config, LayerWeights, Plan

LN_EPS = 1e-5


@dataclass
class LayerCtx:
    x1: torch.Tensor = None      # fp32 stream after PEG (input of the attention LayerNorm)
    x2: torch.Tensor = None      # fp32 stream after attention (input of the FF LayerNorm)
    q: torch.Tensor = None       # bf16 [R, inner]
    kv: torch.Tensor = None      # bf16 [R, 2*inner]
    o: torch.Tensor = None       # bf16 [R, inner]
    lse: torch.Tensor = None     # fp32 [R, heads]
    u: torch.Tensor = None       # bf16 [R, 2*FP]  GEGLU adjoint factors [gelu(gate) | value*gelu'(gate)] (tcgen05 path)
                                 #                 or the FF pre-activation [value | gate] (SIMT comparator path)


@dataclass
class Ctx:
    """Everything the backward pass and the attribution methods need from one forward."""
    B: int = 0
    T: int = 0
    volume: torch.Tensor = None
    vol_stride: int = 0
    alpha: Optional[torch.Tensor] = None
    x_lin: torch.Tensor = None                 # patch-embedding Linear output (input of LayerNorm(dim))
    x_in: torch.Tensor = None                  # tokens entering the spatial transformer
    spatial_in: List[torch.Tensor] = field(default_factory=list)   # keep_stream: input of each spatial layer
    spatial: List[LayerCtx] = field(default_factory=list)
    temporal: List[LayerCtx] = field(default_factory=list)
    x_s_last: torch.Tensor = None              # stream leaving the last spatial layer (input of norm_out)
    x_s_out: torch.Tensor = None               # spatial norm_out output = input of temporal layer 0
    x_t_last: torch.Tensor = None
    x_pre_vq: torch.Tensor = None              # temporal norm_out output (fp32)
    indices: torch.Tensor = None               # int32 [R]
    pooled: torch.Tensor = None                # fp32 [B, HW*C]
    latent: torch.Tensor = None                # fp32 [B, NL] (un-normalised)
    image_latents: torch.Tensor = None
    sim: torch.Tensor = None                   # fp32 [B, Bt]
    dlatent: torch.Tensor = None
    text_latents: torch.Tensor = None
    tokens: Optional[torch.Tensor] = None      # fp32 [R, C] quantised tokens (optional)
    # gradient captures for Grad-CAM (visualizations.py:140-218): grads of the residual stream
    grads: Dict[str, torch.Tensor] = field(default_factory=dict)


@dataclass
class OcclusionCache:
    """Baseline activations an occlusion sweep re-uses for every window (Engine.occlusion_baseline)."""
    T: int = 0
    spatial_in: List[torch.Tensor] = field(default_factory=list)   # fp32 [T*HW, C]: input of each spatial layer
    x_s_out: torch.Tensor = None                                   # fp32 [T*HW, C]: spatial norm_out output
    e_mask: torch.Tensor = None                                    # fp32 [C]: token embedding of a fully occluded patch
    sim: torch.Tensor = None                                       # fp32 [1, Bt]: un-occluded logits


INT_MIN = -2 ** 31


def occlusion_frame_tables(cubes, cube_shape, T: int, H: int, n_layers: int):
    """Host-side index tables of Engine.forward_occluded for a batch of occlusion cubes.

    cubes int [Wn,3] = (t0,h0,w0) token coordinates, cube_shape = (nt,nh,nw).  The changed frames of window w
    entering spatial layer l are t0 .. t0+n_l-1 with n_0 = nt and n_{l+1} = min(n_l + 2, T - t0): the causal
    stencil reads frames t-2..t, so a changed frame t' reaches t', t'+1, t'+2 (attention.py:55-83,
    ctvit.py:60-61).  Compact buffers hold the windows' changed frames back to back.  Returns
    ([src0, rows, peg_src_0 .. peg_src_{L-1}, full], [F_0 .. F_{L-1}]) as int64 arrays:
      src0      [Wn*nt]     baseline frames to copy as layer-0 input, encoded -1 - t
      rows      [Wn*nt*nh*nw] token rows of the compact layer-0 input covered by the cube
      peg_src_l [F_l*3]     per output frame the source of dt = -2,-1,0: >= 0 compact frame of the previous
                            buffer, < 0 baseline frame -1 - t, INT_MIN causal zero padding
      full      [Wn*T]      per (window, t): compact frame of the last buffer or baseline -1 - t"""
    nt, nh, nw = cube_shape
    HW = H * H
    cubes = np.asarray(cubes, dtype=np.int64).reshape(-1, 3)
    Wn = len(cubes)
    t0 = cubes[:, 0]
    n_prev = np.full(Wn, nt, dtype=np.int64)
    off_prev = np.arange(Wn, dtype=np.int64) * nt
    tables = [(-1 - (t0[:, None] + np.arange(nt)[None, :])).reshape(-1)]
    jt, jh, jw = np.meshgrid(np.arange(nt), np.arange(nh), np.arange(nw), indexing="ij")
    rows = ((off_prev[:, None] + jt.reshape(1, -1)) * HW + (cubes[:, 1:2] + jh.reshape(1, -1)) * H
            + cubes[:, 2:3] + jw.reshape(1, -1)).reshape(-1)
    tables.append(rows)
    layer_F = []
    for _ in range(n_layers):
        n_out = np.minimum(n_prev + 2, T - t0)
        off_out = np.concatenate([[0], np.cumsum(n_out)[:-1]])
        F = int(n_out.sum())
        widx = np.repeat(np.arange(Wn), n_out)
        i = np.arange(F) - off_out[widx]                       # frame offset inside the window's changed set
        src = np.empty((F, 3), dtype=np.int64)
        for k, dt in enumerate((-2, -1, 0)):
            ip = i + dt                                        # offset of the source frame relative to t0
            tt = t0[widx] + ip
            in_prev = (ip >= 0) & (ip < n_prev[widx])
            src[:, k] = np.where(tt < 0, INT_MIN, np.where(in_prev, off_prev[widx] + ip, -1 - tt))
        tables.append(src.reshape(-1))
        layer_F.append(F)
        n_prev, off_prev = n_out, off_out
    tt = np.arange(T)[None, :] - t0[:, None]                                           # [Wn, T]
    full = np.where((tt >= 0) & (tt < n_prev[:, None]), off_prev[:, None] + tt, -1 - np.arange(T)[None, :])
    tables.append(full.reshape(-1))
    return tables, layer_F


class Engine:
    def __init__(self, plan: Plan, gemm_impl: int = _lib.GEMM_TCGEN05):
        self.plan = plan
        self.cfg = plan.cfg
        self.dev = plan.device
        self.gemm_impl = gemm_impl
        self.attn_tc = True          # spatial attention forward on tcgen05/TMEM (ctc_attention_fwd_tc)
        self._row_cap = 0

    # ------------------------------------------------------------------ helpers
    def _empty(self, *shape, dtype=torch.float32):
        """Uninitialised device buffer.  While `_row_cap` is set (occlusion fast path: the number of changed frames
        varies from batch to batch), 2-D buffers are carved out of a `_row_cap`-row allocation so that the caching
        allocator sees the same block sizes every batch instead of a fresh cudaMalloc per new shape."""
        cap = self._row_cap
        if cap and len(shape) == 2 and shape[0] <= cap:
            return torch.empty(cap, shape[1], dtype=dtype, device=self.dev)[:shape[0]]
        return torch.empty(*shape, dtype=dtype, device=self.dev)

    def gemm(self, a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, epi: int, bias=None, resid=None, aux=None):
        """out = a[M,K] @ w[N,K]^T with the epilogue `epi` (+bias)(+resid); `aux`: GEGLU pre-activation buffer"""
        M, K = a.shape
        N = w.shape[0]
        assert w.shape[1] == K and out.shape[0] == M
        call("ctc_gemm_bf16", a, a.stride(0), w, w.stride(0), out, out.stride(0), M, N, K, epi, bias, resid,
             resid.stride(0) if resid is not None else 0, aux, aux.stride(0) if aux is not None else 0,
             self.gemm_impl | self.plan.gemm_flags.get(w.data_ptr(), 0), stream_ptr())
        return out

    def layernorm(self, x, g, b, y_bf16=None, y_f32=None, xraw=None):
        R, C = x.shape
        call("ctc_layernorm_fwd", x, R, C, g, b, LN_EPS, y_bf16, y_f32, xraw, stream_ptr())

    def layernorm_bwd(self, dy, x, g, out, accumulate, out_bf16=None):
        """dy fp32 or bf16 (the GEMM that produced it wrote bf16: half the bytes of this HBM-bound pass)"""
        R, C = x.shape
        name = "ctc_layernorm_bwd_bf16" if dy.dtype == torch.bfloat16 else "ctc_layernorm_bwd"
        call(name, dy, x, R, C, g, LN_EPS, out, int(accumulate), out_bf16, stream_ptr())

    # ------------------------------------------------------------------ text tower tail
    def text_latents(self, text_embeds: torch.Tensor) -> torch.Tensor:
        """l2norm(to_text_latent(e)) — ctclip.py:115,119.  text_embeds fp32 [Bt, dim_text]."""
        e = text_embeds.to(self.dev, torch.float32).contiguous()
        out = self._empty(e.shape[0], self.cfg.dim_latent)
        call("ctc_text_latent", e, self.plan.wt, e.shape[0], e.shape[1], self.cfg.dim_latent, out, stream_ptr())
        return out

    # ------------------------------------------------------------------ one transformer layer
    def _layer_fwd(self, x0, lw: LayerWeights, B, T, mode, save: bool, frames=None) -> Tuple[torch.Tensor, LayerCtx]:
        """One transformer layer.  `frames` = (x_base, frame_src int32 [F,3]) runs the PEG stencil over a compact
        frame list (occlusion fast path): x0 holds the changed frames of the previous layer, the output has
        B*T = F frames."""
        cfg = self.cfg
        C = x0.shape[1]
        H = W = cfg.hw
        R = B * T * H * W
        inner, FP = cfg.inner, cfg.ff_pad
        bf = torch.bfloat16
        lc = LayerCtx()
        # x = peg(x) + x                                                   attention.py:325
        x1 = self._empty(R, C)
        if frames is None:
            assert x0.shape[0] == R
            call("ctc_peg", x0, B, T, H, W, C, lw.w27, lw.peg_bias, mode, 0, x1, None, stream_ptr())
        else:
            assert mode == MODE_SPATIAL and B == 1
            call("ctc_peg_frames", x0, frames[0], frames[1], T, H, W, C, lw.w27, lw.peg_bias, x1, stream_ptr())
        # attention: q from LayerNorm(x), k/v from the RAW x                attention.py:138-142
        xn, xraw = self._empty(R, C, dtype=bf), self._empty(R, C, dtype=bf)
        self.layernorm(x1, lw.ln_g, lw.ln_b, y_bf16=xn, xraw=xraw)
        q = self.gemm(xn, lw.wq, self._empty(R, inner, dtype=bf), EPI_BF16)
        kv = self.gemm(xraw, lw.wkv, self._empty(R, 2 * inner, dtype=bf), EPI_BF16)
        o, lse = self._empty(R, inner, dtype=bf), self._empty(R, cfg.heads)
        if (mode == MODE_SPATIAL and self.attn_tc and (H * W) % 64 == 0 and W % 8 == 0 and H * W <= 640
                and 0.0 < getattr(lw, "score_bound", 0.0) < 43.0):
            # tcgen05 / TMEM kernel with the fixed-shift softmax (scores are bounded by lw.score_bound)
            call("ctc_attention_fwd_tc", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, B, T, H, W, cfg.heads,
                 lw.q_scale, lw.k_scale, cfg.attn_scale, self.plan.bias_table, lw.score_bound, o, lse, stream_ptr())
        else:
            call("ctc_attention_fwd", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, B, T, H, W, cfg.heads,
                 lw.q_scale, lw.k_scale, cfg.attn_scale, self.plan.bias_table if mode == MODE_SPATIAL else None, mode,
                 o, lse, stream_ptr())
        # x = to_out(attn) + x                                             attention.py:182, 328
        x2 = self.gemm(o, lw.wout, self._empty(R, C), EPI_F32, resid=x1)
        # x = ff(x) + x                                                    attention.py:43-51, 334
        xn2 = xn  # reuse
        self.layernorm(x2, lw.ff_ln_w, lw.ff_ln_b, y_bf16=xn2)
        # Linear(dim, 2*inner) + GEGLU fused in the GEMM epilogue; the adjoint factors u = [gelu(gate) | value *
        # gelu'(gate)] are only written when a backward pass will follow
        u = self._empty(R, 2 * FP, dtype=bf) if save else None
        hff = self._empty(R, FP, dtype=bf)
        if self.gemm_impl != _lib.GEMM_SIMT:
            self.gemm(xn2, lw.w1, hff, _lib.EPI_GEGLU, aux=u)
        else:   # SIMT comparator path (tests): plain GEMM + stand-alone GEGLU
            u = self.gemm(xn2, lw.w1, self._empty(R, 2 * FP, dtype=bf), EPI_BF16)
            call("ctc_geglu_fwd", u, R, FP, hff, stream_ptr())
        x3 = self.gemm(hff, lw.w2, self._empty(R, C), EPI_F32, resid=x2)
        if save:
            lc.x1, lc.x2, lc.q, lc.kv, lc.o, lc.lse, lc.u = x1, x2, q, kv, o, lse, u
        return x3, lc

    # ------------------------------------------------------------------ forward
    def forward(self, volume: torch.Tensor, text_latents: Optional[torch.Tensor], batch: Optional[int] = None,
                alpha: Optional[torch.Tensor] = None, occl: Optional[torch.Tensor] = None,
                occl_value: float = -1.0, save: bool = False, want_tokens: bool = False,
                keep_attn: bool = False, keep_stream: bool = False, stop_after_patch_emb: bool = False) -> Ctx:
        """volume fp32 [Bv, 1, D, H, W] (Bv == batch, or Bv == 1 shared by all `batch` rows: windows /
        alpha steps of one volume).  alpha fp32 [batch] (IG interpolation), occl int32 [batch, 6]
        (occlusion cubes).  `save` keeps what backward needs; `keep_attn` keeps q/kv/lse only (rollout);
        `keep_stream` keeps the input of every spatial layer and the spatial output (occlusion baseline cache)."""
        cfg, pl = self.cfg, self.plan
        if not (volume.is_cuda and volume.dtype == torch.float32 and volume.is_contiguous()):
            raise RuntimeError(f"ctclip_b200: the volume must be a contiguous fp32 CUDA tensor (there is no CPU path); "
                               f"got device={volume.device}, dtype={volume.dtype}, contiguous={volume.is_contiguous()}")
        if volume.dim() != 5 or volume.shape[1] != 1:
            raise ValueError(f"ctclip_b200: expected a volume of shape [B, 1, D, H, W], got {tuple(volume.shape)}")
        Bv, _, D, Hv, Wv = volume.shape
        B = batch or Bv
        if Bv not in (1, B):
            raise ValueError(f"ctclip_b200: {Bv} volumes for a batch of {B} rows (must be 1 or {B})")
        if Hv != cfg.image_size or Wv != cfg.image_size or D % cfg.temporal_patch_size or D == 0:
            raise ValueError(f"ctclip_b200: volume {D}x{Hv}x{Wv} does not tile into {cfg.temporal_patch_size}x"
                             f"{cfg.patch_size}x{cfg.patch_size} patches of a {cfg.image_size}-pixel model")
        T, H = D // cfg.temporal_patch_size, cfg.hw
        R, C, P = B * T * H * H, cfg.dim, cfg.patch_dim
        bf = torch.bfloat16
        ctx = Ctx(B=B, T=T, volume=volume, vol_stride=(D * Hv * Wv if Bv == B and B > 1 else 0), alpha=alpha)
        keep = save or keep_attn

        # to_patch_emb: patchify + LN(P) -> Linear(P, dim) + bias -> LN(dim)      ctvit.py:44-52
        a_pe = self._empty(R, P, dtype=bf)
        call("ctc_patchify_ln_fwd", volume, ctx.vol_stride, B, D, Hv, Wv, cfg.temporal_patch_size, cfg.patch_size,
             pl.pe_ln1_w, pl.pe_ln1_b, LN_EPS, alpha, occl, occl_value, a_pe, stream_ptr())
        x_lin = self.gemm(a_pe, pl.pe_w, self._empty(R, C), EPI_F32, bias=pl.pe_b)
        del a_pe
        x = self._empty(R, C)
        self.layernorm(x_lin, pl.pe_ln2_w, pl.pe_ln2_b, y_f32=x)
        if save:
            ctx.x_lin, ctx.x_in = x_lin, x
        if stop_after_patch_emb:
            ctx.x_in = x
            return ctx

        # spatial transformer over '(b t) (h w) d'                               ctvit.py:94-96
        for lw in pl.spatial:
            if keep_stream:
                ctx.spatial_in.append(x)
            x, lc = self._layer_fwd(x, lw, B, T, MODE_SPATIAL, keep)
            ctx.spatial.append(lc)
        xs = self._empty(R, C)
        self.layernorm(x, pl.spatial_norm_g, pl.spatial_norm_b, y_f32=xs)
        if save:
            ctx.x_s_last = x
        if save or keep_stream:
            ctx.x_s_out = xs
        return self._forward_tail(ctx, xs, text_latents, save, keep, want_tokens)

    def _forward_tail(self, ctx: Ctx, xs: torch.Tensor, text_latents: Optional[torch.Tensor], save: bool, keep: bool,
                      want_tokens: bool) -> Ctx:
        """Temporal transformer, VQ and CLIP head on the spatial transformer's output xs fp32 [R, C]."""
        cfg, pl = self.cfg, self.plan
        B, T, H = ctx.B, ctx.T, cfg.hw
        R, C = xs.shape
        bf = torch.bfloat16
        x = xs
        # temporal transformer over '(b h w) t d'                                ctvit.py:99-101
        for lw in pl.temporal:
            x, lc = self._layer_fwd(x, lw, B, T, MODE_TEMPORAL, keep)
            ctx.temporal.append(lc)
        xt, xt_bf = self._empty(R, C), self._empty(R, C, dtype=bf)
        self.layernorm(x, pl.temporal_norm_g, pl.temporal_norm_b, y_f32=xt, y_bf16=xt_bf)
        if save:
            ctx.x_t_last = x
        ctx.x_pre_vq = xt

        # VQ (cosine codebook, arg-max)                                           ctvit.py:115-118
        K = cfg.codebook_size
        n_cand = _lib.vq_num_candidates(K)
        cand_val = self._empty(R, n_cand)
        cand_idx = self._empty(R, n_cand, dtype=torch.int32)
        ind = self._empty(R, dtype=torch.int32)
        call("ctc_vq_argmax", xt, xt_bf, R, C, pl.codebook, pl.codebook_bf16, K, cand_val, cand_idx, ind, stream_ptr())
        ctx.indices = ind
        # tokens.mean(dim=1) -> view(B, -1) -> to_visual_latent -> l2norm -> sim   ctclip.py:110-127
        HW = H * H
        pooled = self._empty(B, HW * C)
        tokens = self._empty(R, C) if want_tokens else None
        call("ctc_vq_gather_pool", ind, pl.codebook, B, T, HW, C, pooled, None, tokens, stream_ptr())
        ctx.pooled, ctx.tokens = pooled, tokens
        if text_latents is None:      # CTViT.forward alone (ctvit.py:105-125): stop after the VQ
            return ctx
        L, NL = HW * C, cfg.dim_latent
        n_chunks = (L + 1023) // 1024
        partial = self._empty(n_chunks, B, NL)
        latent = self._empty(B, NL)
        call("ctc_latent_proj", pooled, pl.wv_bf16, pl.wv_lo_bf16, B, L, NL, partial, n_chunks, latent, stream_ptr())
        Bt = text_latents.shape[0]
        sim, il = self._empty(B, Bt), self._empty(B, NL)
        dlat = self._empty(B, NL) if save else None
        call("ctc_latent_sim", latent, text_latents, B, Bt, NL, pl.temp_exp, sim, il, dlat, stream_ptr())
        ctx.latent, ctx.image_latents, ctx.sim, ctx.dlatent = latent, il, sim, dlat
        ctx.text_latents = text_latents
        return ctx

    # ------------------------------------------------------------------ occlusion fast path
    def occlusion_baseline(self, volume: torch.Tensor, text_latents: torch.Tensor, fill: float = -1.0
                           ) -> OcclusionCache:
        """Un-occluded forward of ONE volume that keeps what every occlusion window can re-use: the input of
        each spatial layer, the spatial output, and the embedding of a patch filled with `fill`
        (LayerNorm of a constant patch is exactly its bias, so that embedding is position independent)."""
        cfg = self.cfg
        assert volume.shape[0] == 1
        ctx = self.forward(volume, text_latents, keep_stream=True)
        D = volume.shape[2]
        tp, ps = cfg.temporal_patch_size, cfg.patch_size
        win = torch.tensor([[0, 0, 0, tp, ps, ps]], dtype=torch.int32, device=self.dev)
        pe = self.forward(volume, None, batch=1, occl=win, occl_value=fill, stop_after_patch_emb=True)
        return OcclusionCache(T=D // tp, spatial_in=ctx.spatial_in, x_s_out=ctx.x_s_out, e_mask=pe.x_in[0].clone(),
                              sim=ctx.sim)

    def forward_occluded(self, cache: OcclusionCache, cubes, cube_shape, text_latents: torch.Tensor) -> Ctx:
        """Forward of len(cubes) occluded copies of the cached volume.  cubes: int [Wn, 3] token coordinates
        (t0, h0, w0) of each occlusion cube, cube_shape = (nt, nh, nw) tokens (visualizations.py:380-381 with
        token-aligned windows).  Identical arithmetic to forward(occl=...), but only the frames a cube can reach
        are recomputed in the spatial transformer: its one cross-frame operator is the causal PEG stencil
        (attention.py:55-83, ctvit.py:60-61), so after layer l the changed frames are t0 .. t0+nt-1+2(l+1);
        every other frame is read from the baseline cache.  The temporal transformer, VQ and head run in full."""
        cfg, pl = self.cfg, self.plan
        T, H = cache.T, cfg.hw
        HW, C = H * H, cfg.dim
        nt, nh, nw = cube_shape
        cubes = np.asarray(cubes, dtype=np.int64).reshape(-1, 3)
        Wn = len(cubes)
        t0 = cubes[:, 0]
        assert (t0 >= 0).all() and (t0 + nt <= T).all() and (cubes[:, 1:] >= 0).all()
        assert (cubes[:, 1] + nh <= H).all() and (cubes[:, 2] + nw <= H).all()
        # ---- host-side frame tables for the whole batch, one upload
        tables, layer_F = occlusion_frame_tables(cubes, cube_shape, T, H, len(pl.spatial))
        rows = tables[1]
        sizes = [len(a) for a in tables]
        packed = torch.from_numpy(np.concatenate(tables).astype(np.int32)).to(self.dev, non_blocking=True)
        views, o = [], 0
        for n in sizes:
            views.append(packed[o:o + n])
            o += n
        # ---- changed input frames: baseline tokens with the cube's rows replaced by the occluded-patch embedding
        self._row_cap = Wn * min(T, nt + 2 * len(pl.spatial)) * HW      # most changed frames a batch can have
        try:
            x = self._empty(Wn * nt * HW, C)
            call("ctc_frames_gather", None, cache.spatial_in[0], views[0], Wn * nt, HW * C, x, stream_ptr())
            call("ctc_rows_fill", x, views[1], len(rows), C, cache.e_mask, stream_ptr())
            ctx = Ctx(B=Wn, T=T)
            for l, lw in enumerate(pl.spatial):
                x, _ = self._layer_fwd(x, lw, 1, layer_F[l], MODE_SPATIAL, False,
                                       frames=(cache.spatial_in[l], views[2 + l]))
            xs_c = self._empty(x.shape[0], C)
            self.layernorm(x, pl.spatial_norm_g, pl.spatial_norm_b, y_f32=xs_c)
        finally:
            self._row_cap = 0
        xs = self._empty(Wn * T * HW, C)
        call("ctc_frames_gather", xs_c, cache.x_s_out, views[-1], Wn * T, HW * C, xs, stream_ptr())
        del x, xs_c
        return self._forward_tail(ctx, xs, text_latents, False, False, False)

    # ------------------------------------------------------------------ backward
    def _layer_bwd(self, dx3, dx3_bf, lc: LayerCtx, lw: LayerWeights, B, T, mode, capture: Optional[dict], tag: str):
        """dx3: fp32 grad of the stream leaving the layer (overwritten), dx3_bf its bf16 copy.
        Returns (dx0 fp32, dx0 bf16): grad of the stream entering the layer."""
        cfg = self.cfg
        R, C = dx3.shape
        H = W = cfg.hw
        inner, FP = cfg.inner, cfg.ff_pad
        bf = torch.bfloat16
        if capture is not None:
            capture[tag + "_ff"] = dx3.clone()          # d/d(ff output)  == grad of x3
        # ---- FeedForward
        # du = GEGLU'(u) * (dx3 @ W2): the adjoint is fused into the dh GEMM's epilogue.  The forward saved the two
        # adjoint factors instead of the pre-activation, so the epilogue does two multiplies per element (round 1
        # evaluated erf / erf' there and was epilogue-bound: 779 us fused against 160 + 243 us for GEMM + pass).
        du = self._empty(R, 2 * FP, dtype=bf)
        if self.gemm_impl != _lib.GEMM_SIMT:
            self.gemm(dx3_bf, lw.w2_t, du, _lib.EPI_GEGLU_BWD, aux=lc.u)
        else:   # SIMT comparator path (tests): lc.u holds the pre-activation
            dh = self.gemm(dx3_bf, lw.w2_t, self._empty(R, FP, dtype=bf), EPI_BF16)
            call("ctc_geglu_bwd", lc.u, dh, R, FP, du, stream_ptr())
            del dh
        dxn2 = self.gemm(du, lw.w1_t, self._empty(R, C, dtype=bf), EPI_BF16)     # consumed once, by the LN adjoint
        del du
        dx2, dx2_bf = dx3, dx3_bf
        self.layernorm_bwd(dxn2, lc.x2, lw.ff_ln_w, dx2, True, dx2_bf)
        if capture is not None:
            capture[tag + "_attn"] = dx2.clone()        # d/d(attention module output) == grad of x2
        # ---- attention
        d_o = self.gemm(dx2_bf, lw.wout_t, self._empty(R, inner, dtype=bf), EPI_BF16)
        dq = self._empty(R, inner, dtype=bf)
        dkv = self._empty(R, 2 * inner, dtype=bf)
        delta = self._empty(R, cfg.heads)
        call("ctc_attention_bwd", lc.q, inner, lc.kv, lc.kv.data_ptr() + inner * 2, 2 * inner, lc.o, d_o, lc.lse,
             B, T, H, W, cfg.heads, lw.q_scale, lw.k_scale, cfg.attn_scale,
             self.plan.bias_table if mode == MODE_SPATIAL else None, mode, dq, inner, dkv,
             dkv.data_ptr() + inner * 2, 2 * inner, delta, stream_ptr())
        dx1 = self.gemm(dkv, lw.wkv_t, dx2, EPI_F32, resid=dx2)           # k/v read the raw stream
        dxn = self.gemm(dq, lw.wq_t, dxn2, EPI_BF16)
        self.layernorm_bwd(dxn, lc.x1, lw.ln_g, dx1, True, None)
        # ---- PEG adjoint
        dx0, dx0_bf = self._empty(R, C), dx2_bf
        call("ctc_peg", dx1, B, T, H, W, C, lw.w27, None, mode, 1, dx0, dx0_bf, stream_ptr())
        return dx0, dx0_bf

    def backward(self, ctx: Ctx, grad_out: Optional[torch.Tensor] = None, sum_over_batch: bool = False,
                 capture_grads: bool = False, to_input: bool = True, gsim: Optional[torch.Tensor] = None
                 ) -> Optional[torch.Tensor]:
        """Input gradient of sum_b sim[b, b % Bt] (the `sim[rank, rank].backward()` of
        visualizations.py:580,786,868,921) w.r.t. the (interpolated) voxels; with `gsim` fp32 [B, Bt]
        the gradient of sum_ij gsim[i,j] * sim[i,j] instead.
        sum_over_batch: accumulate all batch rows into `grad_out` [D,H,W] (IG partial sum, +=).
        capture_grads: keep the residual-stream gradients Grad-CAM reads."""
        cfg, pl = self.cfg, self.plan
        B, T = ctx.B, ctx.T
        H = cfg.hw
        HW, C = H * H, cfg.dim
        R = B * T * HW
        bf = torch.bfloat16
        cap = ctx.grads if capture_grads else None
        L, NL = HW * C, cfg.dim_latent
        dpooled = self._empty(B, L)
        if gsim is not None:
            Bt = ctx.text_latents.shape[0]
            g = gsim.to(self.dev, torch.float32).contiguous()
            assert g.shape == (B, Bt)
            call("ctc_latent_sim_bwd", ctx.latent, ctx.text_latents, g, B, Bt, NL, pl.temp_exp, ctx.dlatent,
                 stream_ptr())
        call("ctc_latent_proj_bwd", ctx.dlatent, pl.wv_bf16, B, L, NL, dpooled, stream_ptr())
        dxt = self._empty(R, C)
        call("ctc_vq_bwd", dpooled, None, ctx.x_pre_vq, B, T, HW, C, 0 if cfg.vq_grad_mode == "ste_l2norm" else 1,
             dxt, stream_ptr())
        if cap is not None:
            # gradient at the VQ output (visualizations.py:140-150): dtokens[b,t,hw,:] = dpooled[b,hw,:]/T
            cap["vq"] = (dpooled.view(B, 1, HW, C) / T).expand(B, T, HW, C).reshape(R, C)
        del dpooled
        dx, dx_bf = self._empty(R, C), self._empty(R, C, dtype=bf)
        self.layernorm_bwd(dxt, ctx.x_t_last, pl.temporal_norm_g, dx, False, dx_bf)
        del dxt
        for i in reversed(range(len(pl.temporal))):
            dx, dx_bf = self._layer_bwd(dx, dx_bf, ctx.temporal[i], pl.temporal[i], B, T, MODE_TEMPORAL, cap,
                                        f"temporal{i}")
        d2, d2_bf = self._empty(R, C), dx_bf
        self.layernorm_bwd(dx, ctx.x_s_last, pl.spatial_norm_g, d2, False, d2_bf)
        dx, dx_bf = d2, d2_bf
        for i in reversed(range(len(pl.spatial))):
            dx, dx_bf = self._layer_bwd(dx, dx_bf, ctx.spatial[i], pl.spatial[i], B, T, MODE_SPATIAL, cap,
                                        f"spatial{i}")
        if not to_input:
            return None
        # patch embedding adjoint
        dlin, dlin_bf = self._empty(R, C), dx_bf
        self.layernorm_bwd(dx, ctx.x_lin, pl.pe_ln2_w, dlin, False, dlin_bf)
        da = self.gemm(dlin_bf, pl.pe_w_t, self._empty(R, cfg.patch_dim, dtype=bf), EPI_BF16)
        _, _, D, Hv, Wv = ctx.volume.shape
        if not sum_over_batch:
            if grad_out is None:
                grad_out = self._empty(B, 1, D, Hv, Wv)
            call("ctc_patchify_ln_bwd", ctx.volume, ctx.vol_stride, B, D, Hv, Wv, cfg.temporal_patch_size,
                 cfg.patch_size, pl.pe_ln1_w, LN_EPS, ctx.alpha, da, grad_out, 0, 1.0, stream_ptr())
            return grad_out
        # IG running sum: per-row gradients, then the rows are added in order by one kernel (no floating-point atomics:
        # the sum does not depend on CTA scheduling and is bit-reproducible).  2 x B x 221 MB of extra traffic per
        # batch of B steps, ~1 % of the batch's time.
        if grad_out is None:
            grad_out = torch.zeros(D, Hv, Wv, device=self.dev)
        rows = self._empty(B, 1, D, Hv, Wv)
        call("ctc_patchify_ln_bwd", ctx.volume, ctx.vol_stride, B, D, Hv, Wv, cfg.temporal_patch_size, cfg.patch_size,
             pl.pe_ln1_w, LN_EPS, ctx.alpha, da, rows, 0, 1.0, stream_ptr())
        call("ctc_batch_sum", rows, B, D * Hv * Wv, 1.0, 1, grad_out, stream_ptr())
        return grad_out

    # ------------------------------------------------------------------ attention probabilities
    def attention_fused(self, ctx: Ctx, kind: str, layer: int, fused: bool = False, colmean: bool = False,
                        fusion: str = "mean") -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """The two reductions of `attn` (attention.py:174) the attention-map methods consume, emitted by the kernel
        without materialising the per-head probabilities: the head-fused matrix (visualizations.py:722-725; spatial
        [B*T, HW, HW], temporal [B*HW, T, T]) and / or the per-head mean over the query axis (:666,671; spatial
        [B*T, heads, HW], temporal [B*HW, heads, T])."""
        cfg = self.cfg
        H = cfg.hw
        spatial = kind == "spatial"
        lc = (ctx.spatial if spatial else ctx.temporal)[layer]
        lw = (self.plan.spatial if spatial else self.plan.temporal)[layer]
        n = H * H if spatial else ctx.T
        n_seq = ctx.B * ctx.T if spatial else ctx.B * H * H
        fm = self._empty(n_seq, n, n) if fused else None
        cm = self._empty(n_seq, cfg.heads, n) if colmean else None
        ws = self._empty(n_seq, cfg.heads, (n + 31) // 32, n) if colmean else None
        call("ctc_attention_fused_probs", lc.q, cfg.inner, lc.kv, 2 * cfg.inner, lc.lse, ctx.B, ctx.T, H, H, cfg.heads,
             lw.q_scale, lw.k_scale, cfg.attn_scale, self.plan.bias_table if spatial else None,
             MODE_SPATIAL if spatial else MODE_TEMPORAL, {"mean": 0, "max": 1}[fusion], fm, cm, ws, stream_ptr())
        return fm, cm

    def attention_probs(self, ctx: Ctx, kind: str, layer: int) -> torch.Tensor:
        """Materialise what Attention.forward returns as `attn` (attention.py:174): spatial ->
        [B*T, heads, HW, HW], temporal -> [B*HW, heads, T, T] (fp32)."""
        cfg = self.cfg
        H = cfg.hw
        spatial = kind == "spatial"
        lc = (ctx.spatial if spatial else ctx.temporal)[layer]
        lw = (self.plan.spatial if spatial else self.plan.temporal)[layer]
        n = H * H if spatial else ctx.T
        n_seq = ctx.B * ctx.T if spatial else ctx.B * H * H
        probs = self._empty(n_seq, cfg.heads, n, n)
        call("ctc_attention_probs", lc.q, cfg.inner, lc.kv, 2 * cfg.inner, lc.lse, ctx.B, ctx.T, H, H, cfg.heads,
             lw.q_scale, lw.k_scale, cfg.attn_scale, self.plan.bias_table if spatial else None,
             MODE_SPATIAL if spatial else MODE_TEMPORAL, probs, stream_ptr())
        return probs
