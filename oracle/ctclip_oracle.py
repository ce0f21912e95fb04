"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A plain-PyTorch fp32 restatement of the CT-CLIP-UT hot path (CTViT 3-D encoder
forward/backward + the numerics of the five attribution methods).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
``ct-clip-ut_b200/`` never does (it fails loudly without its CUDA library).

Every function cites the reference file:line it follows (paths relative to the
upstream repo root, i.e. ``/root/reference``).

Parity pinning
--------------
The reference ships no tests or golden vectors.  This oracle is
pinned instead against *the reference itself executed in the build container*:
``tests/golden/make_golden.py`` imports the upstream ``utils/attention.py``,
``utils/ctvit.py``, ``models/ctclip.py`` and ``utils/visualizations.py``
unmodified (with import stubs for the absent I/O packages), runs them on CPU on
seeded inputs and stores the outputs as fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those fixtures.

One piece stays **parity unpinned**: the vector-quantiser is the third-party
``vector_quantize_pytorch`` (lucidrains), not vendored and not version-pinned by
the reference.  Its forward value (``E[argmax cos]``) is version independent and
is restated here from the published algorithm; its backward Jacobian differs
between releases, so ``vq_grad_mode`` selects one of the two published
straight-through variants (default ``ste_l2norm``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# configuration                                                               #
# --------------------------------------------------------------------------- #
@dataclass(frozen=True)
class CTConfig:
    """Hyper-parameters of `CTViT(...)` / `CTCLIP(...)` as wired in
    src/inference_ctclip.py:21-39."""

    dim: int = 512
    codebook_size: int = 8192
    image_size: int = 480
    patch_size: int = 20
    temporal_patch_size: int = 10
    spatial_depth: int = 4
    temporal_depth: int = 4
    dim_head: int = 32
    heads: int = 8
    depth_voxels: int = 240          # D of the input volume
    dim_text: int = 768
    dim_latent: int = 512
    attn_scale: float = 8.0          # attention.py:97 `scale = 8`
    vq_grad_mode: str = "ste_l2norm"  # or "ste_raw"

    @property
    def t(self) -> int:
        return self.depth_voxels // self.temporal_patch_size

    @property
    def h(self) -> int:
        return self.image_size // self.patch_size

    @property
    def w(self) -> int:
        return self.image_size // self.patch_size

    @property
    def n_tokens(self) -> int:
        return self.t * self.h * self.w

    @property
    def patch_dim(self) -> int:
        return self.temporal_patch_size * self.patch_size * self.patch_size

    @property
    def ff_inner(self) -> int:
        # attention.py:43  inner_dim = int(mult * (2 / 3) * dim), mult = 4
        return int(4 * (2 / 3) * self.dim)

    @property
    def dim_image(self) -> int:
        # models/ctclip.py:110-112 : mean over t, flatten (h, w, c)
        return self.h * self.w * self.dim


FULL = CTConfig()
# A tiny configuration with the same structure (t == h == w is required by the
# temporal-PEG axis scramble).
TINY = CTConfig(dim=64, codebook_size=128, image_size=24, patch_size=4,
                temporal_patch_size=2, spatial_depth=2, temporal_depth=2,
                dim_head=32, heads=2, depth_voxels=12, dim_text=48, dim_latent=32)


# --------------------------------------------------------------------------- #
# seeded synthetic weights / inputs                              #
# --------------------------------------------------------------------------- #
def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound


def fitted_codebook(path=None) -> Tensor:
    """The fitted-codebook fixture (tests/golden/make_fitted_codebook.py): l2norm(int8 rows * per-row scale),
    fp32 [1, K, dim].  Decoding is element-wise + one row norm, so every host gets the same bits."""
    from pathlib import Path
    path = path or Path(__file__).resolve().parent.parent / "tests" / "golden" / "fitted_codebook.npz"
    z = np.load(path)
    cb = torch.from_numpy(z["q"].astype(np.float32)) * torch.from_numpy(z["scale"])[:, None]
    return (cb / cb.double().norm(dim=-1, keepdim=True).float())[None].contiguous()


def init_state_dict(cfg: CTConfig = FULL, seed: int = 42, codebook: str = "random") -> Dict[str, Tensor]:
    """Random-init weights under the reference's state-dict keys
    (shapes per attention.py:45-50,59,112-124, ctvit.py:37-66,
    ctclip.py:62-68).  Scales mimic the PyTorch default initialisers; LayerNorm
    gains / biases and q/k scales are perturbed so that no term is trivially
    the identity.  codebook="fitted" (benchmark configuration only) replaces the random
    unit-vector codebook by the k-means fit of the encoder's own outputs stored in
    tests/golden/fitted_codebook.npz — what a trained checkpoint's codebook looks like."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    C, P = cfg.dim, cfg.patch_dim
    inner = cfg.dim_head * cfg.heads
    ffi = cfg.ff_inner

    def lin(name, out_f, in_f, bias=True):
        b = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = _uniform(g, (out_f, in_f), b)
        if bias:
            sd[name + ".bias"] = _uniform(g, (out_f,), b)

    def ln(name, d, wkey="weight", bkey="bias", zero_bias=False):
        sd[f"{name}.{wkey}"] = 1.0 + 0.1 * _uniform(g, (d,), 1.0)
        sd[f"{name}.{bkey}"] = torch.zeros(d) if zero_bias else 0.05 * _uniform(g, (d,), 1.0)

    vt = "visual_transformer."
    ln(vt + "to_patch_emb.1", P)
    lin(vt + "to_patch_emb.2", C, P)
    ln(vt + "to_patch_emb.3", C)
    p2 = cfg.patch_size ** 2
    ln(vt + "to_patch_emb_first_frame.1", p2)
    lin(vt + "to_patch_emb_first_frame.2", C, p2)
    ln(vt + "to_patch_emb_first_frame.3", C)
    lin(vt + "spatial_rel_pos_bias.net.0.0", C, 2)
    lin(vt + "spatial_rel_pos_bias.net.1.0", C, C)
    lin(vt + "spatial_rel_pos_bias.net.2", cfg.heads, C)
    for tname, depth in (("enc_spatial_transformer", cfg.spatial_depth),
                         ("enc_temporal_transformer", cfg.temporal_depth)):
        for i in range(depth):
            p = f"{vt}{tname}.layers.{i}."
            sd[p + "0.dsconv.weight"] = _uniform(g, (C, 1, 3, 3, 3), 1.0 / math.sqrt(27))
            sd[p + "0.dsconv.bias"] = _uniform(g, (C,), 1.0 / math.sqrt(27))
            sd[p + "1.null_kv"] = torch.zeros(cfg.heads, 0, cfg.dim_head)
            sd[p + "1.q_scale"] = 1.0 + 0.1 * _uniform(g, (cfg.dim_head,), 1.0)
            sd[p + "1.k_scale"] = 1.0 + 0.1 * _uniform(g, (cfg.dim_head,), 1.0)
            ln(p + "1.norm", C, "gamma", "beta", zero_bias=True)
            ln(p + "1.context_norm", C, "gamma", "beta", zero_bias=True)
            lin(p + "1.to_q", inner, C, bias=False)
            lin(p + "1.to_kv", 2 * inner, C, bias=False)
            lin(p + "1.to_out", C, inner, bias=False)
            ln(p + "3.0", C)
            lin(p + "3.1", 2 * ffi, C, bias=False)
            lin(p + "3.4", C, ffi, bias=False)
        ln(f"{vt}{tname}.norm_out", C, "gamma", "beta", zero_bias=True)
    cb = _uniform(g, (1, cfg.codebook_size, C), 1.0)
    sd[vt + "vq._codebook.embed"] = F.normalize(cb, dim=-1)
    if codebook == "fitted":
        fc = fitted_codebook()
        assert tuple(fc.shape) == (1, cfg.codebook_size, C), "the fitted codebook exists for the FULL config only"
        sd[vt + "vq._codebook.embed"] = fc
    elif codebook != "random":
        raise ValueError(f"unknown codebook {codebook!r}")
    sd[vt + "vq._codebook.initted"] = torch.tensor([True])
    sd[vt + "vq._codebook.cluster_size"] = torch.zeros(1, cfg.codebook_size)
    lin("to_text_latent", cfg.dim_latent, cfg.dim_text, bias=False)
    lin("to_visual_latent", cfg.dim_latent, cfg.dim_image, bias=False)
    sd["temperature"] = torch.tensor(1.0)
    return sd


def synthetic_volume(cfg: CTConfig = FULL, index: int = 0, batch: int = 1) -> Tensor:
    """SURVEY §8d synthetic input: clamp(0.35*randn - 0.2, -1, 1) with a constant
    -1 border (mimics preprocess.py:135-147: HU/1000 clamped to [-1,1], padded
    with -1 so that constant patches exist).  Shape [B,1,D,H,W] fp32."""
    D, H = cfg.depth_voxels, cfg.image_size
    bd = max(1, (16 * D) // 240)
    bh = max(1, (40 * H) // 480)
    vols = []
    for b in range(batch):
        g = torch.Generator().manual_seed(1234 + index + b)
        x = (0.35 * torch.randn(D, H, H, generator=g) - 0.2).clamp_(-1, 1)
        x[:bd] = -1; x[-bd:] = -1
        x[:, :bh] = -1; x[:, -bh:] = -1
        x[:, :, :bh] = -1; x[:, :, -bh:] = -1
        vols.append(x)
    return torch.stack(vols)[:, None].contiguous()


def synthetic_text_embeds(cfg: CTConfig = FULL, seed: int = 7, batch: int = 1) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, cfg.dim_text, generator=g)


# --------------------------------------------------------------------------- #
# L1 building blocks  (src/utils/attention.py)                                #
# --------------------------------------------------------------------------- #
def l2norm(t: Tensor) -> Tensor:
    """attention.py:20-21  F.normalize(t, dim=-1) (eps 1e-12)."""
    return t / t.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def layer_norm(x: Tensor, weight: Tensor, bias: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    """attention.py:27-34 (gamma, zero beta buffer) and nn.LayerNorm (attention.py:46,
    ctvit.py:49,51): biased variance over the last dim, eps 1e-5."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    y = (x - mu) * torch.rsqrt(var + eps) * weight
    return y + bias if bias is not None else y


def gelu(x: Tensor) -> Tensor:
    """F.gelu default (erf form), used by GEGLU attention.py:38-41."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def cpb_rel_pos(h: int, w: int, device=None) -> Tensor:
    """attention.py:262-268: grid of (i - j) offsets for an h x w token grid with
    sign(d)*log(|d|+1).  Returns [hw, hw, 2] fp32."""
    ii, jj = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    grid = torch.stack([ii, jj]).reshape(2, -1).t()          # [(h w), 2]
    rel = grid[:, None, :] - grid[None, :, :]                # i - j
    rel = torch.sign(rel) * torch.log(rel.abs() + 1)
    return rel.to(torch.float32)


def cpb_bias(sd: Dict[str, Tensor], prefix: str, h: int, w: int) -> Tensor:
    """ContinuousPositionBias.forward attention.py:259-277:
    Linear(2,dim)+LeakyReLU(0.1) -> Linear(dim,dim)+LeakyReLU(0.1) -> Linear(dim,heads);
    'i j h -> h i j'."""
    x = cpb_rel_pos(h, w, device=sd[prefix + "net.0.0.weight"].device)
    x = F.leaky_relu(x @ sd[prefix + "net.0.0.weight"].t() + sd[prefix + "net.0.0.bias"], 0.1)
    x = F.leaky_relu(x @ sd[prefix + "net.1.0.weight"].t() + sd[prefix + "net.1.0.bias"], 0.1)
    x = x @ sd[prefix + "net.2.weight"].t() + sd[prefix + "net.2.bias"]
    return x.permute(2, 0, 1).contiguous()


def peg(x: Tensor, shape: Tuple[int, int, int, int], weight: Tensor, bias: Tensor) -> Tensor:
    """PEG.forward attention.py:61-83 with causal=True (ctvit.py:60-61).
    `x` is [B', n, C]; it is *flat-reinterpreted* as (b, t, h, w, C) (attention.py:69),
    zero-padded (1,1) on the last two grid axes and (2,0) on the first, passed through
    a depthwise 3x3x3 Conv3d (cross-correlation) and reshaped back.  Returns conv(x)
    only — the residual add lives in Transformer.forward (attention.py:325)."""
    orig = x.shape
    C = x.shape[-1]
    v = x.reshape(*shape, C).permute(0, 4, 1, 2, 3)          # b d t h w
    v = F.pad(v, (1, 1, 1, 1, 2, 0), value=0.0)
    v = F.conv3d(v, weight, bias, groups=C)
    v = v.permute(0, 2, 3, 4, 1)
    return v.reshape(orig)


def attention(x: Tensor, sd: Dict[str, Tensor], p: str, heads: int, scale: float,
              attn_bias: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """Attention.forward attention.py:126-182, self-attention instance
    (causal=False, num_null_kv=0, no mask, dropout 0).  Returns (out, attn[b,h,n,n]).
    QUIRK (attention.py:138-142): `kv_input = default(context, x)` is taken BEFORE
    `x = self.norm(x)`, so q is computed from LayerNorm(x) but k and v from the RAW,
    un-normalised x."""
    b, n, _ = x.shape
    xn = layer_norm(x, sd[p + "norm.gamma"], sd[p + "norm.beta"])
    q = xn @ sd[p + "to_q.weight"].t()
    kv = x @ sd[p + "to_kv.weight"].t()
    k, v = kv.chunk(2, dim=-1)
    split = lambda t_: t_.reshape(b, n, heads, -1).permute(0, 2, 1, 3)   # b h n d
    q, k, v = split(q), split(k), split(v)
    q = l2norm(q) * sd[p + "q_scale"]
    k = l2norm(k) * sd[p + "k_scale"]
    sim = torch.matmul(q, k.transpose(-1, -2)) * scale
    if attn_bias is not None:
        sim = sim + attn_bias
    attn = sim.softmax(dim=-1)
    out = torch.matmul(attn, v)
    out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
    return out @ sd[p + "to_out.weight"].t(), attn


def feedforward(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    """FeedForward attention.py:43-51: LayerNorm -> Linear(dim, 2*inner, no bias) ->
    GEGLU (x = first half, gate = second half; gelu(gate)*x) -> Linear(inner, dim)."""
    y = layer_norm(x, sd[p + "0.weight"], sd[p + "0.bias"])
    y = y @ sd[p + "1.weight"].t()
    a, gate = y.chunk(2, dim=-1)
    y = gelu(gate) * a
    return y @ sd[p + "4.weight"].t()


def transformer(x: Tensor, sd: Dict[str, Tensor], p: str, depth: int,
                video_shape: Tuple[int, int, int, int], heads: int, scale: float,
                attn_bias: Optional[Tensor], capture: Optional[dict] = None,
                kind: str = "spatial") -> Tensor:
    """Transformer.forward attention.py:322-336 with layers [PEG, Attention, None, FF].
    `capture` mirrors the forward-hook protocol of visualizations.py:153-263: the
    attention module output (feature map, probs) and the FF output of every layer are
    recorded (graph-attached so the caller can take gradients)."""
    for i in range(depth):
        lp = f"{p}layers.{i}."
        x = peg(x, video_shape, sd[lp + "0.dsconv.weight"], sd[lp + "0.dsconv.bias"]) + x
        a_out, probs = attention(x, sd, lp + "1.", heads, scale, attn_bias)
        x = a_out + x
        f_out = feedforward(x, sd, lp + "3.")
        x = f_out + x
        if capture is not None:
            capture.setdefault(kind + "_features", []).append(a_out)
            capture.setdefault(kind + "_attention_weights", []).append(probs)
            capture.setdefault(kind + "_ff_features", []).append(f_out)
    return layer_norm(x, sd[p + "norm_out.gamma"], sd[p + "norm_out.beta"])


# --------------------------------------------------------------------------- #
# L2 model  (src/utils/ctvit.py, src/models/ctclip.py)                        #
# --------------------------------------------------------------------------- #
def patchify(image: Tensor, cfg: CTConfig) -> Tensor:
    """ctvit.py:44-48 Rearrange 'b c (t pt) (h p1) (w p2) -> b t h w (c pt p1 p2)'."""
    b, c, D, H, W = image.shape
    pt, p = cfg.temporal_patch_size, cfg.patch_size
    x = image.reshape(b, c, D // pt, pt, H // p, p, W // p, p)
    x = x.permute(0, 2, 4, 6, 1, 3, 5, 7)
    return x.reshape(b, D // pt, H // p, W // p, c * pt * p * p)


def patch_embed(image: Tensor, sd: Dict[str, Tensor], cfg: CTConfig,
                p: str = "visual_transformer.to_patch_emb.") -> Tensor:
    """ctvit.py:44-52 / :112 : patchify -> LayerNorm(P) -> Linear(P, dim) -> LayerNorm(dim)."""
    x = patchify(image, cfg)
    x = layer_norm(x, sd[p + "1.weight"], sd[p + "1.bias"])
    x = x @ sd[p + "2.weight"].t() + sd[p + "2.bias"]
    return layer_norm(x, sd[p + "3.weight"], sd[p + "3.bias"])


def encode(tokens: Tensor, sd: Dict[str, Tensor], cfg: CTConfig, capture: Optional[dict] = None,
           vt: str = "visual_transformer.") -> Tensor:
    """CTViT.encode ctvit.py:88-103."""
    b, t, h, w, C = tokens.shape
    attn_bias = cpb_bias(sd, vt + "spatial_rel_pos_bias.", h, w)
    video_shape = (b, t, h, w)
    x = tokens.reshape(b * t, h * w, C)                                   # (b t) (h w) d
    x = transformer(x, sd, vt + "enc_spatial_transformer.", cfg.spatial_depth, video_shape,
                    cfg.heads, cfg.attn_scale, attn_bias, capture, "spatial")
    x = x.reshape(b, t, h, w, C).permute(0, 2, 3, 1, 4).reshape(b * h * w, t, C)   # (b h w) t d
    x = transformer(x, sd, vt + "enc_temporal_transformer.", cfg.temporal_depth, video_shape,
                    cfg.heads, cfg.attn_scale, None, capture, "temporal")
    return x.reshape(b, h, w, t, C).permute(0, 3, 1, 2, 4)               # b t h w d


def vq_cosine(x: Tensor, codebook: Tensor, grad_mode: str = "ste_l2norm",
              force_indices: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """`VectorQuantize(dim, codebook_size, use_cosine_sim=True)` in train() mode with
    freeze_codebook=True (ctvit.py:66,117-118).  Third-party, restated from the published
    algorithm (lucidrains/vector-quantize-pytorch, CosineSimCodebook): x̂ = l2norm(x.float());
    dist = x̂ · Eᵀ (E unit rows); ind = argmax; quantize = onehot(ind) @ E; straight-through.
    Forward value is E[ind].  Backward (PARITY UNPINNED, see module docstring):
      ste_l2norm : out = x̂ + (E[ind] − x̂).detach()   (releases that normalise first)
      ste_raw    : out = x  + (E[ind] − x ).detach()   (older releases)
    x: [b, n, d]; codebook: [1, K, d].  Returns (out [b,n,d], ind [b,n] int64).
    `force_indices` (test aid, not in the reference) overrides the arg-max so that a comparison
    can be conditioned on identical code assignments."""
    # the library's codebook forward is decorated `@autocast(enabled=False)` and casts to fp32: the distance GEMM and
    # the arg-max are fp32 even when the model runs under Accelerate's fp16 autocast (CTClipInference.py:56-63)
    with torch.autocast(device_type=x.device.type, enabled=False):
        x = x.float()
        E = codebook[0].float()
        xh = l2norm(x)
        dist = xh @ E.t()
        ind = dist.argmax(dim=-1) if force_indices is None else force_indices.reshape(x.shape[:-1]).long()
        q = E[ind]
        base = xh if grad_mode == "ste_l2norm" else x
        if grad_mode not in ("ste_l2norm", "ste_raw"):
            raise ValueError(f"unknown vq_grad_mode {grad_mode}")
        out = base + (q - base).detach()
    return out, ind


def ctvit_forward(image: Tensor, sd: Dict[str, Tensor], cfg: CTConfig,
                  capture: Optional[dict] = None, force_indices: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """CTViT.forward ctvit.py:105-125 (model_type='ctclip').  Returns
    (tokens [b,t,h,w,dim], codebook indices [b, t*h*w])."""
    vt = "visual_transformer."
    tokens = patch_embed(image, sd, cfg)
    if capture is not None:
        capture["patch_tokens"] = tokens
    tokens = encode(tokens, sd, cfg, capture)
    b, t, h, w, C = tokens.shape
    flat = tokens.reshape(b, t * h * w, C)
    if capture is not None:
        capture["pre_vq"] = flat
    q, ind = vq_cosine(flat, sd[vt + "vq._codebook.embed"], cfg.vq_grad_mode, force_indices)
    if capture is not None:
        capture["vq_features"] = q
    return q.reshape(b, t, h, w, C), ind


def ctclip_forward(image: Tensor, text_embeds: Tensor, sd: Dict[str, Tensor], cfg: CTConfig,
                   capture: Optional[dict] = None, force_indices: Optional[Tensor] = None):
    """CTCLIP.forward models/ctclip.py:99-129 on the `text_embeds` path (text tower
    bypassed, :107), world size 1 (gather is the identity, :94-97).
    Returns (sim [B,B], image_latents, text_latents, exp(T), image_tokens, indices)."""
    tokens, ind = ctvit_forward(image, sd, cfg, capture, force_indices)
    img = tokens.mean(dim=1)                                  # over t  (ctclip.py:111)
    img = img.reshape(img.shape[0], -1)                       # (h w c) (ctclip.py:112)
    text_latents = text_embeds @ sd["to_text_latent.weight"].t()
    image_latents = img @ sd["to_visual_latent.weight"].t()
    text_latents = text_latents / text_latents.norm(dim=-1, keepdim=True)
    image_latents = image_latents / image_latents.norm(dim=-1, keepdim=True)
    temp = sd["temperature"].exp()
    sim = image_latents @ text_latents.t() * temp
    return sim, image_latents, text_latents, temp, tokens, ind


# --------------------------------------------------------------------------- #
# L3 attribution numerics  (src/utils/visualizations.py)                      #
# --------------------------------------------------------------------------- #
def occlusion_windows(shape: Tuple[int, int, int], patch_size=(20, 40, 40), stride=(10, 20, 20)
                      ) -> List[Tuple[int, int, int]]:
    """visualizations.py:339-349: nested d -> h -> w enumeration (w fastest)."""
    D, H, W = shape
    return [(d, h, w)
            for d in range(0, D - patch_size[0] + 1, stride[0])
            for h in range(0, H - patch_size[1] + 1, stride[1])
            for w in range(0, W - patch_size[2] + 1, stride[2])]


def shard_windows(windows: Sequence, rank: int, world: int) -> List:
    """visualizations.py:351-361: per_rank = total // world; list truncated to
    per_rank*world (remainder dropped); contiguous slices."""
    per = len(windows) // world
    trimmed = list(windows)[: per * world]
    return trimmed[rank * per:(rank + 1) * per]


def occlusion_mask_apply(image: Tensor, window, patch_size, value: float = -1.0) -> Tensor:
    """visualizations.py:380-381."""
    d, h, w = window
    out = image.clone()
    out[:, :, d:d + patch_size[0], h:h + patch_size[1], w:w + patch_size[2]] = value
    return out


def occlusion_accumulate(shape, windows, patch_size, original_score: float, occluded_scores
                         ) -> Tuple[np.ndarray, np.ndarray]:
    """visualizations.py:366-367, 390-392: float64 numpy heat / count maps,
    importance = max(orig - occ, 0)."""
    heat = np.zeros(shape)
    count = np.zeros(shape)
    for (d, h, w), s in zip(windows, occluded_scores):
        imp = max(original_score - float(s), 0)
        heat[d:d + patch_size[0], h:h + patch_size[1], w:w + patch_size[2]] += imp
        count[d:d + patch_size[0], h:h + patch_size[1], w:w + patch_size[2]] += 1
    return heat, count


def occlusion_finalize(heat: np.ndarray, count: np.ndarray, threshold: float = 0.0,
                       rotate: bool = True) -> np.ndarray:
    """visualizations.py:404-424 (after the cross-rank SUM): fp32 tensors, count==0 -> 1,
    divide, (h-min)/(max-min+1e-8), identity-size trilinear interpolate, threshold, rot90."""
    ht = torch.tensor(heat, dtype=torch.float32)
    ct = torch.tensor(count, dtype=torch.float32)
    ct[ct == 0] = 1
    ht = ht / ct
    ht = (ht - ht.min()) / (ht.max() - ht.min() + 1e-8)
    full = F.interpolate(ht[None, None], size=tuple(ht.shape), mode="trilinear",
                         align_corners=False)[0, 0].numpy()
    full[full < threshold] = 0
    return np.rot90(full, k=-1, axes=(1, 2)) if rotate else full


def attention_rollout(attn_weights_list: Sequence[Tensor], head_fusion: str = "mean",
                      discard_ratio: float = 0.0, use_residual: bool = True) -> Tensor:
    """Visualizations.attention_rollout visualizations.py:707-743."""
    n = attn_weights_list[0].size(-1)
    result = torch.eye(n, device=attn_weights_list[0].device)
    for attn in attn_weights_list:
        if head_fusion == "mean":
            attn = attn.mean(dim=0)
        elif head_fusion == "max":
            attn = attn.max(dim=0)[0]
        else:
            raise ValueError(f"Unsupported head_fusion: {head_fusion}")
        if discard_ratio > 0:
            flat = attn.reshape(attn.shape[0], -1)
            num_discard = int(flat.shape[1] * discard_ratio)
            thr = flat.topk(flat.shape[1] - num_discard, dim=1)[0].min(dim=1, keepdim=True)[0]
            attn = torch.where(attn >= thr, attn, torch.zeros_like(attn))
        attn = attn / (attn.sum(dim=-1, keepdim=True) + 1e-8)
        if use_residual:
            attn = attn + torch.eye(attn.size(0), device=attn.device)
            attn = attn / attn.sum(dim=-1, keepdim=True)
        result = attn @ result
    return result


def norm_minmax_range(v: Tensor) -> Tensor:
    """(v-min)/(max-min+1e-8): rollout :812,:841 and occlusion :414."""
    return (v - v.min()) / (v.max() - v.min() + 1e-8)


def norm_minmax_max(v: Tensor) -> Tensor:
    """(v-min)/(max+1e-8): raw attention :674, Grad-CAM :946-947,:970-971,:991, IG first :882."""
    return (v - v.min()) / (v.max() + 1e-8)


def rollout_maps(spatial_attn: Sequence[Tensor], temporal_attn: Sequence[Tensor], grid=(24, 24, 24)
                 ) -> Tuple[Tensor, Tensor]:
    """visualize_attention_rollout visualizations.py:795-841, up to (excluding) the
    upsample.  spatial_attn: per layer [t, heads, hw, hw]; temporal_attn: per layer
    [hw, heads, t, t].  Spatial: one single-matrix 'rollout' per (layer, slice) ->
    column sums -> [layers*t, h, w] (layer-major; the reference comment says 24^3 but the
    stack is 96 deep).  Temporal: per (h,w) token chain the layers -> column sums ->
    [hw, t] -> view(h, w, t) -> permute(2,0,1)."""
    D, H, W = grid
    rows = []
    for blk in spatial_attn:
        for d in range(blk.shape[0]):
            r = attention_rollout([blk[d]])
            rows.append(r.sum(dim=0).reshape(H, W))
    vol = torch.stack(rows, dim=0)
    vol = norm_minmax_range(vol)
    trows = []
    for tok in range(temporal_attn[0].shape[0]):
        r = attention_rollout([layer[tok] for layer in temporal_attn])
        trows.append(r.sum(dim=0))
    tvol = torch.stack(trows).reshape(H, W, D).permute(2, 0, 1)
    tvol = norm_minmax_range(tvol)
    return vol, tvol


def raw_attention_maps(attn_list: Sequence[Tensor], mode: str, grid=(24, 24, 24)) -> Tensor:
    """visualize_attention_grid_gif visualizations.py:659-676 reductions: per layer, per
    head: mean over the query axis; spatial -> view(D,H,W); temporal -> view(H,W,D) ->
    permute(2,0,1); (v-min)/(max+1e-8); np.rot90(k=-1, axes=(0,1)).
    Returns [heads, layers, D, H, W]-shaped (post-rot90) fp32 tensor."""
    D, H, W = grid
    out = []
    for head in range(attn_list[0].shape[1]):
        per_layer = []
        for attn in attn_list:
            rec = attn[:, head].mean(dim=1)
            vol = rec.reshape(D, H, W) if mode == "spatial" else rec.reshape(H, W, D).permute(2, 0, 1)
            vol = norm_minmax_max(vol)
            per_layer.append(torch.from_numpy(np.rot90(vol.cpu().numpy(), k=-1, axes=(0, 1)).copy()))
        out.append(torch.stack(per_layer))
    return torch.stack(out)


def _cam(features: Tensor, grads: Tensor) -> Tensor:
    w = grads.mean(dim=(0, 1))
    return (features * w.view(1, 1, -1)).sum(dim=-1).relu()


def grad_cam_maps(sim_scalar: Tensor, capture: dict, grid=(24, 24, 24)) -> Dict[str, Tensor]:
    """visualize_grad_cam visualizations.py:918-991 up to (excluding) upsample.
    Reproduces the hook-order quirk: forward hooks append features in
    layer order, tensor grad hooks append gradients in *backward* order, and the code
    takes `[-1]` of both lists — i.e. LAST-layer features with FIRST-layer gradients."""
    D, H, W = grid
    names = ["spatial_features", "temporal_features", "spatial_ff_features", "temporal_ff_features"]
    tensors = [capture[n] for n in names]
    flat = [t_ for lst in tensors for t_ in lst] + [capture["vq_features"]]
    grads = torch.autograd.grad(sim_scalar, flat, allow_unused=False)
    gi = 0
    g: Dict[str, List[Tensor]] = {}
    for n, lst in zip(names, tensors):
        g[n] = list(grads[gi:gi + len(lst)])
        gi += len(lst)
    vq_grad = grads[gi]
    # backward order == reversed forward order; [-1] of that is layer 0
    sel = lambda n: (capture[n][-1].detach(), g[n][0])
    sff = _cam(*sel("spatial_ff_features")).reshape(D, H, W)
    tff = _cam(*sel("temporal_ff_features")).reshape(H, W, D).permute(2, 0, 1)
    sff, tff = norm_minmax_max(sff), norm_minmax_max(tff)
    sp = _cam(*sel("spatial_features")).reshape(D, H, W)
    tp = _cam(*sel("temporal_features")).reshape(H, W, D).permute(2, 0, 1)
    sp, tp = norm_minmax_max(sp), norm_minmax_max(tp)
    combined = torch.sqrt(sp * tp + 1e-8)
    vq_f = capture["vq_features"].detach().squeeze(0)
    vq_g = vq_grad.squeeze(0)
    wv = vq_g.mean(dim=0)
    vq_cam = (vq_f * wv).sum(dim=-1).relu().reshape(D, H, W)
    vq_cam = norm_minmax_max(vq_cam)
    return {"spatial_ff": sff, "temporal_ff": tff, "spatial": sp, "temporal": tp,
            "combined": combined, "vq": vq_cam}


def integrated_gradients_raw(image: Tensor, score_fn, steps: int = 50) -> Tuple[Tensor, List[float]]:
    """visualize_integrated_gradients visualizations.py:853-879: baseline = ones,
    alpha in linspace(0,1,steps), x_a = 1 + a*(x-1), grad of the score w.r.t. x_a,
    avg over steps, ig = relu((x-1) * avg).  `score_fn(x)` returns the scalar
    sim[rank,rank].  Returns (ig [D,H,W], list of scores)."""
    baseline = torch.ones_like(image)
    diff = image - baseline
    grads, scores = [], []
    for alpha in torch.linspace(0, 1, steps):
        xa = (baseline + alpha.to(image.device) * diff).detach().requires_grad_()
        s = score_fn(xa)
        (gr,) = torch.autograd.grad(s, xa)
        grads.append(gr.detach())
        scores.append(float(s))
    avg = torch.stack(grads).mean(dim=0)
    ig = (diff * avg).squeeze().relu()
    return ig, scores


def integrated_gradients_post(ig: Tensor, rotate: bool = True) -> np.ndarray:
    """visualizations.py:882-901: (ig-min)/(max+1e-8); host np.quantile(.,0.90), zero
    below; ** 0.05; /(max+1e-8); rot90."""
    ig = (ig - ig.min()) / (ig.max() + 1e-8)
    a = ig.cpu().numpy()
    q = np.quantile(a, 0.90)
    a = np.where(a >= q, a, 0.0)
    a = a ** 0.05
    a = a / (a.max() + 1e-8)
    return np.rot90(a, k=-1, axes=(1, 2)) if rotate else a


def upsample(x: Tensor, target_shape) -> np.ndarray:
    """Visualizations._upsample visualizations.py:289-293."""
    return F.interpolate(x[None, None].float(), size=tuple(target_shape), mode="trilinear",
                         align_corners=False).squeeze().detach().cpu().to(torch.float32).numpy()


def rot90(a: np.ndarray) -> np.ndarray:
    """np.rot90(k=-1, axes=(1,2)) used by every method (e.g. visualizations.py:816)."""
    return np.rot90(a, k=-1, axes=(1, 2))


# --------------------------------------------------------------------------- #
# convenience drivers used by tests / bench                                   #
# --------------------------------------------------------------------------- #
def to_device(sd: Dict[str, Tensor], device) -> Dict[str, Tensor]:
    return {k: v.to(device) for k, v in sd.items()}


def occlusion_scores(image: Tensor, text_embeds: Tensor, sd, cfg: CTConfig, windows, patch_size
                     ) -> Tuple[float, List[float]]:
    """_compute_occlusion hot loop visualizations.py:370-388 (no_grad, text_embeds path)."""
    with torch.no_grad():
        orig = float(ctclip_forward(image, text_embeds, sd, cfg)[0][0, 0])
        scores = []
        for win in windows:
            occ = occlusion_mask_apply(image, win, patch_size)
            scores.append(float(ctclip_forward(occ, text_embeds, sd, cfg)[0][0, 0]))
    return orig, scores


# --------------------------------------------------------------------------- #
# zero-shot scoring  (src/utils/CTClipInference.py)                           #
# --------------------------------------------------------------------------- #
def zero_shot_predictions(image_latents: Tensor, pair_latents: Tensor, temp: Tensor, rank: int = 0) -> Tensor:
    """CTClipInference.validate_prompts + the scoring lines of zeroshot (CTClipInference.py:133-145, 171-180)
    for ONE sample and P pathologies.  `pair_latents` [P,2,d]: row 0 = "There is X.", row 1 = "There is no X."
    (the tokenizer call at :159-165 puts them at even / odd rows, split at :137-138).  Per pathology:
    present = (il @ tl_present.T * temp).diag()[rank], absent likewise, p = softmax([present, absent])[0],
    stored into a float64 vector (:156, 180).  Returns float64 [P]."""
    out = torch.zeros(pair_latents.shape[0], dtype=torch.double)
    for j in range(pair_latents.shape[0]):
        tl = pair_latents[j]
        present = torch.diag(image_latents @ tl[0::2].t() * temp)[rank]
        absent = torch.diag(image_latents @ tl[1::2].t() * temp)[rank]
        out[j] = torch.softmax(torch.stack([present, absent]), dim=0)[0]
    return out
