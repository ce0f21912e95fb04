"""Host-side mirror of the reference module interface (same class names, constructor keywords,
sub-module names and state-dict keys — SURVEY §8b), backed by the sm_100a engine.

  reference                                   here
  src/utils/attention.py   LayerNorm, PEG, Attention, FeedForward, ContinuousPositionBias, Transformer
                                              -> parameter containers with identical keys/shapes
  src/utils/ctvit.py       CTViT              -> CTViT   (forward = fused CUDA path up to the VQ)
  src/models/ctclip.py     CTCLIP             -> CTCLIP  (forward returns the same 5-tuple)

The per-block `forward`s of the containers are intentionally not implemented: the fused path never
executes module-by-module (that is the point), and a silent PyTorch fallback is not allowed here.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch
import torch.distributed as dist
from torch import nn

from . import _lib
from .engine import Ctx, Engine
from .plan import Config, Plan


def _no_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__}.forward is not executed on the fused sm_100a path; call "
                       "CTViT.forward / CTCLIP.forward (there is no per-module PyTorch fallback)")


class LayerNorm(nn.Module):
    """attention.py:27-34 — learnable gamma, zero beta buffer."""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))

    forward = _no_forward


class PEG(nn.Module):
    """attention.py:55-83"""

    def __init__(self, dim, causal=False):
        super().__init__()
        self.causal = causal
        self.dsconv = nn.Conv3d(dim, dim, 3, groups=dim)

    forward = _no_forward


class Attention(nn.Module):
    """attention.py:87-124 (self-attention instance: num_null_kv = 0)."""

    def __init__(self, dim, dim_context=None, dim_head=64, heads=8, causal=False, num_null_kv=0,
                 norm_context=True, dropout=0.0, scale=8):
        super().__init__()
        self.heads, self.causal, self.scale = heads, causal, scale
        inner = dim_head * heads
        dim_context = dim_context or dim
        self.norm = LayerNorm(dim)
        self.context_norm = LayerNorm(dim_context) if norm_context else nn.Identity()
        self.num_null_kv = num_null_kv
        self.null_kv = nn.Parameter(torch.randn(heads, 2 * num_null_kv, dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim_context, inner * 2, bias=False)
        self.q_scale = nn.Parameter(torch.ones(dim_head))
        self.k_scale = nn.Parameter(torch.ones(dim_head))
        self.to_out = nn.Linear(inner, dim, bias=False)

    forward = _no_forward


def FeedForward(dim, mult=4, dropout=0.0):
    """attention.py:43-51 — indices 0 (LayerNorm), 1 (Linear), 4 (Linear) carry parameters."""
    inner = int(mult * (2 / 3) * dim)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner * 2, bias=False), nn.Identity(),
                         nn.Identity(), nn.Linear(inner, dim, bias=False))


class ContinuousPositionBias(nn.Module):
    """attention.py:230-257"""

    def __init__(self, *, dim, heads, num_dims=2, layers=2, log_dist=True, cache_rel_pos=False):
        super().__init__()
        self.net = nn.ModuleList([nn.Sequential(nn.Linear(num_dims, dim), nn.LeakyReLU(0.1))])
        for _ in range(layers - 1):
            self.net.append(nn.Sequential(nn.Linear(dim, dim), nn.LeakyReLU(0.1)))
        self.net.append(nn.Linear(dim, heads))

    forward = _no_forward


class Transformer(nn.Module):
    """attention.py:281-320 — layers[i] = [PEG, Attention, None, FeedForward], norm_out."""

    def __init__(self, dim, *, depth, dim_head=64, heads=8, ff_mult=4, peg=False, peg_causal=False,
                 attn_dropout=0.0, ff_dropout=0.0, **_):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _i in range(depth):
            self.layers.append(nn.ModuleList([
                PEG(dim=dim, causal=peg_causal) if peg else None,
                Attention(dim=dim, dim_head=dim_head, heads=heads, dropout=attn_dropout),
                None,
                FeedForward(dim=dim, mult=ff_mult, dropout=ff_dropout)]))
        self.norm_out = LayerNorm(dim)

    forward = _no_forward


class _Codebook(nn.Module):
    def __init__(self, dim, codebook_size):
        super().__init__()
        self.register_buffer("initted", torch.tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed", nn.functional.normalize(torch.randn(1, codebook_size, dim), dim=-1))
        self.register_buffer("embed_avg", torch.zeros(1, codebook_size, dim))


class VectorQuantize(nn.Module):
    """Parameter container for `vector_quantize_pytorch.VectorQuantize(dim, codebook_size,
    use_cosine_sim=True)` (ctvit.py:66): state-dict key `_codebook.embed [1, K, dim]`."""

    def __init__(self, dim, codebook_size, **_):
        super().__init__()
        self._codebook = _Codebook(dim, codebook_size)

    forward = _no_forward


class CTViT(nn.Module):
    """Drop-in for src/utils/ctvit.py:9-125."""

    def __init__(self, dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
                 spatial_depth=4, temporal_depth=4, dim_head=64, heads=8, channels=1, attn_dropout=0.0,
                 ff_dropout=0.0, model_type="ctclip"):
        super().__init__()
        if channels != 1 or model_type != "ctclip":
            raise RuntimeError("ctclip_b200.CTViT implements the CT-CLIP path only (channels=1, model_type='ctclip')")
        self.model_type = model_type
        self.image_size, self.patch_size, self.temporal_patch_size = image_size, patch_size, temporal_patch_size
        self.patch_height = self.patch_width = image_size // patch_size
        self.dim, self.codebook_size, self.heads, self.dim_head = dim, codebook_size, heads, dim_head
        self.spatial_depth, self.temporal_depth = spatial_depth, temporal_depth
        self.vq_grad_mode = "ste_l2norm"
        self.spatial_rel_pos_bias = ContinuousPositionBias(dim=dim, heads=heads)
        p2 = channels * patch_size ** 2
        self.to_patch_emb_first_frame = nn.Sequential(nn.Identity(), nn.LayerNorm(p2), nn.Linear(p2, dim),
                                                      nn.LayerNorm(dim))
        P = p2 * temporal_patch_size
        self.to_patch_emb = nn.Sequential(nn.Identity(), nn.LayerNorm(P), nn.Linear(P, dim), nn.LayerNorm(dim))
        kw = dict(dim=dim, dim_head=dim_head, heads=heads, attn_dropout=attn_dropout, ff_dropout=ff_dropout,
                  peg=True, peg_causal=True)
        self.enc_spatial_transformer = Transformer(depth=spatial_depth, **kw)
        self.enc_temporal_transformer = Transformer(depth=temporal_depth, **kw)
        self.vq = VectorQuantize(dim=dim, codebook_size=codebook_size)
        self._standalone_engine: Optional[Engine] = None

    def load(self, path, strict=False):
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(f"Model state file not found at: {path}")
        try:
            self.load_state_dict(torch.load(str(path)), strict)
            print(f"Successfully loaded state dictionary from: {path}")
        except Exception as e:
            raise RuntimeError(f"Failed to load state dictionary from {path}: {e}")

    def load_state_dict(self, *a, **k):
        self._standalone_engine = None
        return super().load_state_dict(*a, **k)

    def config(self, dim_text=768, dim_latent=512) -> Config:
        return Config(dim=self.dim, codebook_size=self.codebook_size, image_size=self.image_size,
                      patch_size=self.patch_size, temporal_patch_size=self.temporal_patch_size,
                      spatial_depth=self.spatial_depth, temporal_depth=self.temporal_depth, dim_head=self.dim_head,
                      heads=self.heads, dim_text=dim_text, dim_latent=dim_latent, vq_grad_mode=self.vq_grad_mode)

    def forward(self, image, return_only_codebook_ids=False):
        """image [B,1,D,H,W] -> tokens [B, D/pt, H/p, W/p, dim] (ctvit.py:105-125); forward only."""
        if self._standalone_engine is None:
            sd = {"visual_transformer." + k: v for k, v in self.state_dict().items()}
            dev = image.device
            if dev.type != "cuda":
                raise RuntimeError(f"ctclip_b200: CTViT runs on sm_100a CUDA devices only (no CPU path); got {dev}")
            sd["to_text_latent.weight"] = torch.zeros(1, 1, device=dev)
            sd["to_visual_latent.weight"] = torch.zeros(1, self.patch_height * self.patch_width * self.dim, device=dev)
            sd["temperature"] = torch.zeros((), device=dev)
            self._standalone_engine = Engine(Plan(sd, self.config(1, 1), dev))
        ctx = self._standalone_engine.forward(image.float().contiguous(), None, want_tokens=True)
        B, T, H = ctx.B, ctx.T, self.patch_height
        if return_only_codebook_ids:
            return ctx.indices.long().view(B, T, H, H)
        return ctx.tokens.view(B, T, H, H, self.dim)


class _SimFunction(torch.autograd.Function):
    """sim = f(image); the input gradient is produced by the hand-written backward kernels."""

    @staticmethod
    def forward(fctx, image, engine: Engine, text_latents, want_tokens: bool, holder: dict):
        need_grad = image.requires_grad
        ctx = engine.forward(image.detach().float().contiguous(), text_latents, save=need_grad,
                             want_tokens=want_tokens)
        fctx.engine, fctx.ectx = engine, ctx
        holder["ctx"] = ctx
        # Fresh output tensors: returning ctx.sim itself would close a reference cycle through the autograd node
        # (sim -> grad_fn -> fctx.ectx -> ctx.sim) that keeps ~1.4 GB of saved activations per volume alive until a
        # cyclic GC pass, and the caching allocator then grows by a full context every step.
        sim, il = ctx.sim.clone(), ctx.image_latents.clone()
        fctx.mark_non_differentiable(il)
        return sim, il

    @staticmethod
    def backward(fctx, gsim, _gil):
        if fctx.ectx is None:
            raise RuntimeError("CTCLIP: the saved activations were released by the first backward pass "
                               "(a second backward through the same forward is not supported)")
        grad = fctx.engine.backward(fctx.ectx, gsim=gsim)
        fctx.ectx = None                 # saved activations are released with the last outside reference to the Ctx
        return grad, None, None, None, None


class CTCLIP(nn.Module):
    """Drop-in for src/models/ctclip.py:44-129."""

    def __init__(self, *, text_encoder, image_encoder, dim_text, dim_image, dim_latent, temperature_init=1.0):
        super().__init__()
        self.text_transformer = text_encoder
        self.visual_transformer = image_encoder
        self.to_text_latent = nn.Linear(dim_text, dim_latent, bias=False)
        self.to_visual_latent = nn.Linear(dim_image, dim_latent, bias=False)
        self.temperature = nn.Parameter(torch.tensor(temperature_init))
        self.dim_text, self.dim_latent = dim_text, dim_latent
        self.return_image_tokens = True
        self._engine: Optional[Engine] = None
        self.last_ctx: Optional[Ctx] = None

    # ------------------------------------------------------------------ state
    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def load(self, path, strict=False):
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(f"Model state file not found at: {path}")
        try:
            sd = torch.load(str(path), map_location=torch.device("cuda" if torch.cuda.is_available() else "cpu"))
            self.load_state_dict(sd, strict)
            print(f"Successfully loaded state dictionary from: {path}")
        except Exception as e:
            raise RuntimeError(f"Failed to load state dictionary from {path}: {e}")

    def engine(self, device=None) -> Engine:
        """Kernel-ready weights are packed on first use (and re-packed after load_state_dict)."""
        if self._engine is None:
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            if dev.type != "cuda":
                raise RuntimeError(f"ctclip_b200: CTCLIP runs on sm_100a CUDA devices only (no CPU path); got {dev}")
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith("text_transformer.")}
            cfg = self.visual_transformer.config(self.dim_text, self.dim_latent)
            self._engine = Engine(Plan(sd, cfg, dev))
        return self._engine

    def gather_features(self, features):
        """The reference all-gathers latents across ranks (ctclip.py:90-97, 123-124) although every
        attribution caller reads only sim[rank, rank]; no communication is issued here."""
        return features

    # ------------------------------------------------------------------ forward
    def forward(self, text_inputs, image_inputs, text_embeds=None):
        """Returns (sim_matrix, image_latents, text_latents, temperature.exp(), image_tokens) — ctclip.py:99-129."""
        eng = self.engine(image_inputs.device)
        self.last_ctx = None             # drop the previous call's saved activations before allocating this call's
        if text_inputs:
            text_output = self.text_transformer(**text_inputs).last_hidden_state[:, 0, :]
        else:
            text_output = text_embeds
        text_latents = eng.text_latents(text_output.detach())
        holder: dict = {}
        sim, image_latents = _SimFunction.apply(image_inputs, eng, text_latents, self.return_image_tokens, holder)
        ctx = holder["ctx"]
        self.last_ctx = ctx
        tokens = None
        if ctx.tokens is not None:
            H = self.visual_transformer.patch_height
            tokens = ctx.tokens.view(ctx.B, ctx.T, H, H, -1)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # keep the reference's indexing contract sim[rank, rank] (visualizations.py:375) without the gather
            w, r, (B, Bt) = dist.get_world_size(), dist.get_rank(), sim.shape
            full = sim.new_zeros(w * B, w * Bt)
            full[r * B:(r + 1) * B, r * Bt:(r + 1) * Bt] = sim
            sim = full
        return sim, image_latents, text_latents, self.temperature.exp(), tokens
