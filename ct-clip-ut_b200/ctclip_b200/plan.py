"""Weight plan: the reference state dict (keys of SURVEY §8b) repacked once for the sm_100a kernels.

Only layout work happens here (casts to bf16, transposes for the input-gradient GEMMs, zero padding
of the 1365-wide FeedForward inner dimension to a TMA-legal 1408, the 27-tap depthwise weights as
[27, C]); the one piece of arithmetic — the continuous-position-bias MLP, which the reference
re-evaluates for all 576² pairs on every forward (attention.py:259-277) — is evaluated once for the
47² distinct offsets by the `ctc_cpb_table` kernel.
"""
from __future__ import annotations

import os

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib


@dataclass(frozen=True)
class Config:
    """Hyper-parameters of CTViT / CTCLIP (src/inference_ctclip.py:21-39)."""
    dim: int = 512
    codebook_size: int = 8192
    image_size: int = 480
    patch_size: int = 20
    temporal_patch_size: int = 10
    spatial_depth: int = 4
    temporal_depth: int = 4
    dim_head: int = 32
    heads: int = 8
    dim_text: int = 768
    dim_latent: int = 512
    attn_scale: float = 8.0
    vq_grad_mode: str = "ste_l2norm"

    @property
    def hw(self) -> int:
        return self.image_size // self.patch_size

    @property
    def patch_dim(self) -> int:
        return self.temporal_patch_size * self.patch_size ** 2

    @property
    def ff_inner(self) -> int:
        return int(4 * (2 / 3) * self.dim)          # attention.py:43

    @property
    def ff_pad(self) -> int:
        return (self.ff_inner + 127) // 128 * 128    # 1365 -> 1408: 16-byte rows for TMA, 128-wide tiles

    @property
    def inner(self) -> int:
        return self.dim_head * self.heads


class LayerWeights:
    __slots__ = ("w27", "peg_bias", "ln_g", "ln_b", "q_scale", "k_scale", "wq", "wkv", "wout", "wq_t", "wkv_t",
                 "wout_t", "ff_ln_w", "ff_ln_b", "w1", "w2", "w1_t", "w2_t", "score_bound")


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).contiguous()


class Plan:
    """Device-resident, kernel-ready weights."""

    def _pack(self, w: torch.Tensor) -> torch.Tensor:
        """bf16 B operand [N, K] of a GEMM.  When N is a multiple of 32 its rows are permuted inside every group of 32
        (ctc_gemm_row_perm) and the tensor is registered in `gemm_flags`, so that Engine.gemm selects the staging-free
        direct epilogue for it; the result of the GEMM is the same either way."""
        t = _bf16(w)
        if t.shape[0] % 32 == 0 and self.direct_epilogue:
            if self._perm is None:
                self._perm = torch.tensor(_lib.gemm_row_perm(), device=t.device, dtype=torch.long)
            t = t.view(t.shape[0] // 32, 32, t.shape[1])[:, self._perm, :].reshape(t.shape).contiguous()
            self.gemm_flags[t.data_ptr()] = _lib.GEMM_BPERM
        return t

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: Config, device: Optional[torch.device] = None,
                 direct_epilogue: Optional[bool] = None):
        _lib.require_device()
        # CTC_GEMM_DIRECT=0: keep the weights unpermuted and use the staged GEMM epilogues (A/B measurement aid)
        self.direct_epilogue = (os.environ.get("CTC_GEMM_DIRECT", "1") != "0") if direct_epilogue is None else direct_epilogue
        self.gemm_flags: Dict[int, int] = {}
        self._perm: Optional[torch.Tensor] = None
        if cfg.dim_head != 32:
            raise RuntimeError("ctclip_b200: the attention kernels are specialised for dim_head == 32 "
                               "(src/inference_ctclip.py:29)")
        self.cfg = cfg
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        sd = {k: v.detach().to(dev) for k, v in state_dict.items() if isinstance(v, torch.Tensor)}
        f32 = lambda k: sd[k].float().contiguous()
        vt = "visual_transformer."
        C, F, FP = cfg.dim, cfg.ff_inner, cfg.ff_pad

        # patch embedding (ctvit.py:44-52)
        self.pe_ln1_w, self.pe_ln1_b = f32(vt + "to_patch_emb.1.weight"), f32(vt + "to_patch_emb.1.bias")
        w = f32(vt + "to_patch_emb.2.weight")                      # [C, P]
        self.pe_w, self.pe_w_t = self._pack(w), self._pack(w.t())
        self.pe_b = f32(vt + "to_patch_emb.2.bias")
        self.pe_ln2_w, self.pe_ln2_b = f32(vt + "to_patch_emb.3.weight"), f32(vt + "to_patch_emb.3.bias")

        # continuous position bias table [heads, (2H-1)(2W-1)]
        p = vt + "spatial_rel_pos_bias.net."
        hw = cfg.hw
        self.bias_table = torch.empty(cfg.heads, (2 * hw - 1) ** 2, device=dev, dtype=torch.float32)
        _lib.call("ctc_cpb_table", f32(p + "0.0.weight"), f32(p + "0.0.bias"), f32(p + "1.0.weight"),
                  f32(p + "1.0.bias"), f32(p + "2.weight"), f32(p + "2.bias"), C, cfg.heads, hw, hw,
                  self.bias_table, _lib.stream_ptr())

        self.spatial = [self._layer(sd, f"{vt}enc_spatial_transformer.layers.{i}.") for i in range(cfg.spatial_depth)]
        self.temporal = [self._layer(sd, f"{vt}enc_temporal_transformer.layers.{i}.") for i in range(cfg.temporal_depth)]
        # bound on every spatial attention score, scale*max|q_scale|*max|k_scale| + max|bias| (attention.py:160-172:
        # q, k are l2-normalised): lets the tcgen05 attention kernel use a fixed softmax shift instead of a running max
        bound = torch.empty(1, device=dev, dtype=torch.float32)
        for lw in self.spatial:
            _lib.call("ctc_attention_score_bound", lw.q_scale, lw.k_scale, cfg.attn_scale, self.bias_table, cfg.heads, hw,
                      hw, bound, _lib.stream_ptr())
            lw.score_bound = float(bound.item())
        self.spatial_norm_g = f32(vt + "enc_spatial_transformer.norm_out.gamma")
        self.spatial_norm_b = f32(vt + "enc_spatial_transformer.norm_out.beta")
        self.temporal_norm_g = f32(vt + "enc_temporal_transformer.norm_out.gamma")
        self.temporal_norm_b = f32(vt + "enc_temporal_transformer.norm_out.beta")

        cb = f32(vt + "vq._codebook.embed")[0].contiguous()       # [K, C]
        self.codebook, self.codebook_bf16 = cb, _bf16(cb)

        wv = f32("to_visual_latent.weight")                         # [NL, L]
        self.wv_bf16 = _bf16(wv)
        # rounding residual of the bf16 weight: the forward projection uses hi + lo (see latent_proj_mma_kernel)
        self.wv_lo_bf16 = _bf16(wv - self.wv_bf16.float())
        del wv
        self.wt = f32("to_text_latent.weight")                     # [NL, DT]
        self.temp_exp = float(sd["temperature"].float().exp())
        torch.cuda.synchronize(dev)

    def _layer(self, sd, p: str) -> LayerWeights:
        cfg = self.cfg
        C, F, FP = cfg.dim, cfg.ff_inner, cfg.ff_pad
        f32 = lambda k: sd[p + k].float().contiguous()
        lw = LayerWeights()
        lw.w27 = f32("0.dsconv.weight").reshape(C, 27).t().contiguous()      # [27, C], tap = (a*3+b)*3+c
        lw.peg_bias = f32("0.dsconv.bias")
        lw.ln_g, lw.ln_b = f32("1.norm.gamma"), f32("1.norm.beta")
        lw.q_scale, lw.k_scale = f32("1.q_scale"), f32("1.k_scale")
        wq, wkv, wout = f32("1.to_q.weight"), f32("1.to_kv.weight"), f32("1.to_out.weight")
        lw.wq, lw.wkv, lw.wout = self._pack(wq), self._pack(wkv), self._pack(wout)
        lw.wq_t, lw.wkv_t, lw.wout_t = self._pack(wq.t()), self._pack(wkv.t()), self._pack(wout.t())
        lw.ff_ln_w, lw.ff_ln_b = f32("3.0.weight"), f32("3.0.bias")
        w1, w2 = f32("3.1.weight"), f32("3.4.weight")                          # [2F, C], [C, F]
        # GEGLU halves (attention.py:40: x, gate = chunk(2)), zero-padded to FP and interleaved in 64-row groups
        # [32 value rows | 32 gate rows] so that a GEMM epilogue thread holds value_j and gate_j of the same j
        val = torch.zeros(FP, C, device=w1.device)
        gate = torch.zeros(FP, C, device=w1.device)
        val[:F], gate[:F] = w1[:F], w1[F:]
        w1p = torch.stack([val.view(FP // 32, 32, C), gate.view(FP // 32, 32, C)], dim=1).reshape(2 * FP, C)
        w2p = torch.zeros(C, FP, device=w2.device)
        w2p[:, :F] = w2
        lw.w1, lw.w2 = self._pack(w1p), self._pack(w2p)
        lw.w1_t, lw.w2_t = self._pack(w1p.t()), self._pack(w2p.t())
        return lw
