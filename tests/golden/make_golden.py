"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/src) on CPU in the build container.

    python tests/golden/make_golden.py [--only tiny|full|attrib] [--ref /root/reference]

The reference cannot be imported whole here: `accelerate`, `nibabel`,
`matplotlib`, `monai`, `vector_quantize_pytorch` are absent and
`ContinuousPositionBias.forward` hard-codes `torch.device('cuda')`.  This script
installs *import stubs* for the I/O packages (never executed on the numeric path), a
`vector_quantize_pytorch.VectorQuantize` stand-in restating the published algorithm
(the one piece that therefore stays "parity unpinned"), and a `torch` proxy inside
`utils.attention` whose only change is `torch.device(...) -> cpu`.  Everything numeric
is executed by the reference's own classes: `Transformer`, `ContinuousPositionBias`,
`CTViT`, `CTCLIP`, `Visualizations`.

Outputs (small, committed):
  tiny_model.npz      real CTViT/CTCLIP at the TINY config (forward, grads)
  full_forward.npz    real CTCLIP forward at the benchmark config (1 volume)
  full_attrib.npz     real Visualizations.{visualize_grad_cam, visualize_attention_rollout,
                      visualize_attention_grid_gif reductions, visualize_integrated_gradients
                      (3 steps), _compute_occlusion (8 coarse windows)} at the benchmark config
  occ_row.npz         real CTCLIP.forward scores of 69 reference-size (20,40,40) occlusion windows and of the
                      64-window (60,120,120) grid, plus the real _compute_occlusion heat map of the latter
  *_fitted.npz        the same three full-size fixtures on the fitted-codebook checkpoint
                      (--codebook fitted; see make_fitted_codebook.py)
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch
from torch import nn

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ctclip_oracle as O  # noqa: E402


# ----------------------------------------------------------------------------- stubs
class _Anything:
    """Importable placeholder: any attribute / call returns another placeholder."""

    def __init__(self, name="stub"):
        self._n = name

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(self._n + "." + k)

    def __call__(self, *a, **k):
        return _Anything(self._n + "()")

    def __mro_entries__(self, bases):
        return (object,)


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    def _ga(k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(name + "." + k)
    m.__getattr__ = _ga  # type: ignore
    sys.modules[name] = m
    return m


class _VQCodebook(nn.Module):
    def __init__(self, dim, codebook_size):
        super().__init__()
        self.register_buffer("initted", torch.tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed", torch.zeros(1, codebook_size, dim))


class VectorQuantizeStandIn(nn.Module):
    """Stand-in for vector_quantize_pytorch.VectorQuantize (absent, unpinned): same ctor
    kwargs and call signature as used at ctvit.py:66,118; arithmetic = oracle.vq_cosine."""
    grad_mode = "ste_l2norm"

    def __init__(self, dim, codebook_size, use_cosine_sim=True, freeze_codebook=False, **kw):
        super().__init__()
        assert use_cosine_sim
        self._codebook = _VQCodebook(dim, codebook_size)

    def forward(self, x, freeze_codebook=False, **kw):
        out, ind = O.vq_cosine(x, self._codebook.embed, self.grad_mode)
        return out, ind, torch.zeros((), device=x.device)


class _TorchCPUProxy:
    """`torch` as seen from utils/attention.py, with torch.device(...) pinned to cpu
    (attention.py:135,169,199,218,261 hard-code 'cuda')."""

    def __getattr__(self, k):
        if k == "device":
            return lambda *a, **kw: torch.device("cpu")
        return getattr(torch, k)


class FakeAccelerator:
    is_main_process = True
    process_index = 0
    num_processes = 1
    device = torch.device("cpu")


class FakeText(nn.Module):
    """text tower stand-in: `text_transformer(**text_inputs).last_hidden_state[:, 0, :]`
    (ctclip.py:107) returns the supplied embedding."""

    def forward(self, embeds=None):
        return types.SimpleNamespace(last_hidden_state=embeds[:, None, :])


def import_reference(ref_root: str):
    src = str(Path(ref_root) / "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    from transformers import BertTokenizer  # noqa: F401  (real; must precede the accelerate stub)
    _stub_module("vector_quantize_pytorch", VectorQuantize=VectorQuantizeStandIn)
    for name in ["nibabel", "matplotlib", "matplotlib.pyplot", "matplotlib.animation",
                 "matplotlib.colors", "matplotlib.lines", "accelerate", "accelerate.utils"]:
        if name not in sys.modules:
            _stub_module(name)
    sys.modules["accelerate"].Accelerator = FakeAccelerator
    import utils.attention as ref_attention
    ref_attention.torch = _TorchCPUProxy()
    import utils.ctvit as ref_ctvit
    import models.ctclip as ref_ctclip
    import utils.visualizations as ref_vis
    # the reference switches on deterministic algorithms at import (visualizations.py:36);
    # keep CPU behaviour unchanged but avoid warnings for ops without deterministic impls
    torch.use_deterministic_algorithms(False)
    return ref_attention, ref_ctvit, ref_ctclip, ref_vis


def build_reference_model(ref_ctvit, ref_ctclip, cfg: O.CTConfig, sd):
    vit = ref_ctvit.CTViT(dim=cfg.dim, codebook_size=cfg.codebook_size, image_size=cfg.image_size,
                          patch_size=cfg.patch_size, temporal_patch_size=cfg.temporal_patch_size,
                          spatial_depth=cfg.spatial_depth, temporal_depth=cfg.temporal_depth,
                          dim_head=cfg.dim_head, heads=cfg.heads)
    clip = ref_ctclip.CTCLIP(text_encoder=FakeText(), image_encoder=vit, dim_text=cfg.dim_text,
                             dim_image=cfg.dim_image, dim_latent=cfg.dim_latent)
    missing, unexpected = clip.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("text_transformer" in m for m in missing), missing
    clip.eval()
    return clip


def sd_checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.dtype.is_floating_point))


# ----------------------------------------------------------------------------- tiny
def make_tiny(refs, out: Path):
    ref_attention, ref_ctvit, ref_ctclip, ref_vis = refs
    cfg = O.TINY
    sd = O.init_state_dict(cfg, seed=42)
    clip = build_reference_model(ref_ctvit, ref_ctclip, cfg, sd)
    img = O.synthetic_volume(cfg, 0, batch=2).requires_grad_()
    txt = O.synthetic_text_embeds(cfg, 7, batch=2)
    sim, il, tl, temp, tokens = clip({"embeds": txt}, img)
    (g_img,) = torch.autograd.grad(sim[0, 0] + 0.5 * sim[1, 1], img)
    # pieces
    vit = clip.visual_transformer
    bias = vit.spatial_rel_pos_bias(cfg.h, cfg.w)
    pe = vit.to_patch_emb(img)
    enc = vit.encode(pe)
    ind = vit(img, return_only_codebook_ids=True)
    np.savez_compressed(
        out, sd_checksum=sd_checksum(sd), sim=sim.detach().numpy(), image_latents=il.detach().numpy(),
        text_latents=tl.detach().numpy(), temp=float(temp), tokens=tokens.detach().numpy(),
        grad_image=g_img.numpy(), cpb_bias=bias.detach().numpy(), patch_emb=pe.detach().numpy(),
        encoded=enc.detach().numpy(), indices=ind.numpy())
    print("tiny ->", out, "sim", sim.detach().numpy().ravel())


# ----------------------------------------------------------------------------- full forward
def make_full_forward(refs, out: Path, codebook: str = "random"):
    ref_attention, ref_ctvit, ref_ctclip, ref_vis = refs
    cfg = O.FULL
    sd = O.init_state_dict(cfg, seed=42, codebook=codebook)
    clip = build_reference_model(ref_ctvit, ref_ctclip, cfg, sd)
    img = O.synthetic_volume(cfg, 0)
    txt = O.synthetic_text_embeds(cfg, 7)
    t0 = time.time()
    with torch.no_grad():
        vit = clip.visual_transformer
        pe = vit.to_patch_emb(img)
        enc = vit.encode(pe)
        sim, il, tl, temp, tokens = clip(None, img, txt)
        ind = vit(img, return_only_codebook_ids=True)
    print(f"full forward (x2) {time.time() - t0:.1f}s  sim={float(sim):.6f}")
    rows = np.arange(0, cfg.n_tokens, 97)
    encf = enc.reshape(-1, cfg.dim)
    # top-2 cosine margin of every token (parity hazard bookkeeping)
    xh = O.l2norm(encf)
    top2 = (xh @ sd["visual_transformer.vq._codebook.embed"][0].t()).topk(2, dim=-1).values
    np.savez_compressed(
        out, sd_checksum=sd_checksum(sd), sim=sim.numpy(), image_latents=il.numpy(), text_latents=tl.numpy(),
        temp=float(temp), indices=ind.numpy().astype(np.int16), rows=rows,
        patch_emb_rows=pe.reshape(-1, cfg.dim)[rows].numpy(), encoded_rows=encf[rows].numpy(),
        encoded_norms=encf.norm(dim=-1).numpy(), vq_margin=(top2[:, 0] - top2[:, 1]).numpy())
    print("full_forward ->", out)


# ----------------------------------------------------------------------------- full attribution
class _NumpyCapture:
    """`np` as seen from utils/visualizations.py: np.save and np.quantile are recorded."""

    def __init__(self):
        self.saved = {}
        self.quantile_inputs = []

    def __getattr__(self, k):
        return getattr(np, k)

    def save(self, path, arr, *a, **k):
        self.saved[Path(str(path)).name] = arr

    def quantile(self, a, q, *args, **kw):
        self.quantile_inputs.append(np.array(a, copy=True))
        return np.quantile(a, q, *args, **kw)


# Occlusion windows scored for the heat-map parity tests (tests/test_gpu_parity.py):
#   row: reference-size (20,40,40) windows — one full row of the sweep (all 23 w positions at h = 200) at three
#        depths, the last one clipped by the end of the volume (frames 22-23)
#   med: the complete non-overlapping (60,120,120) grid, 4 x 4 x 4 = 64 windows (a sweep the reference runs with
#        patch_size = stride = (60,120,120))
OCC_ROW_PATCH = (20, 40, 40)
OCC_ROW_WINDOWS = [(d, 200, w) for d in (50, 100, 220) for w in range(0, 441, 20)]
OCC_MED_PATCH = (60, 120, 120)


def make_occ_row(refs, out: Path, codebook: str = "random"):
    """Window scores through the real CTCLIP.forward, exactly as the hot loop of _compute_occlusion runs them
    (visualizations.py:376-388: clone, fill -1, forward, sim[0, 0]); the `med` heat map is produced by the real
    `_compute_occlusion` itself."""
    ref_attention, ref_ctvit, ref_ctclip, ref_vis = refs
    cfg = O.FULL
    sd = O.init_state_dict(cfg, seed=42, codebook=codebook)
    clip = build_reference_model(ref_ctvit, ref_ctclip, cfg, sd)
    img = O.synthetic_volume(cfg, 0)
    txt = O.synthetic_text_embeds(cfg, 7)
    t0 = time.time()
    res = {"sd_checksum": sd_checksum(sd), "codebook": np.array(codebook)}

    def score(windows, ps, tag):
        scores = []
        with torch.no_grad():
            for (d, h, w) in windows:
                occ = img.clone()
                occ[:, :, d:d + ps[0], h:h + ps[1], w:w + ps[2]] = -1
                scores.append(float(clip(None, occ, txt)[0][0, 0]))
                print(f"  {tag} window {(d, h, w)}: {scores[-1]:+.6f}  ({time.time() - t0:.0f} s)", flush=True)
        return np.array(scores)

    with torch.no_grad():
        res["orig"] = float(clip(None, img, txt)[0][0, 0])
        res["base_indices"] = clip.visual_transformer(img, return_only_codebook_ids=True).reshape(-1).numpy().astype(np.int16)
    res["row_patch"] = np.array(OCC_ROW_PATCH)
    res["row_windows"] = np.array(OCC_ROW_WINDOWS)
    res["row_scores"] = score(OCC_ROW_WINDOWS, OCC_ROW_PATCH, "row")
    med = O.occlusion_windows((cfg.depth_voxels, cfg.image_size, cfg.image_size), OCC_MED_PATCH, OCC_MED_PATCH)
    res["med_patch"] = np.array(OCC_MED_PATCH)
    res["med_windows"] = np.array(med)
    res["med_scores"] = score(med, OCC_MED_PATCH, "med")
    # the med heat map from the reference's own _compute_occlusion (another 65 forwards): sub-sampled
    import tempfile
    vis = ref_vis.Visualizations(clip, FakeAccelerator(), None, None, 1, Path(tempfile.mkdtemp()), "", None)
    heat = vis._compute_occlusion(img, None, txt, OCC_MED_PATCH, OCC_MED_PATCH, 0.0)
    res["med_heat_sub"] = np.rot90(heat, k=1, axes=(1, 2))[5::10, 10::20, 10::20].astype(np.float32)
    np.savez_compressed(out, **res)
    print("occ_row ->", out, "orig", res["orig"])


def make_full_attrib(refs, out: Path, ig_steps: int = 3, codebook: str = "random"):
    ref_attention, ref_ctvit, ref_ctclip, ref_vis = refs
    cfg = O.FULL
    sd = O.init_state_dict(cfg, seed=42, codebook=codebook)
    clip = build_reference_model(ref_ctvit, ref_ctclip, cfg, sd)
    img = O.synthetic_volume(cfg, 0)
    txt = O.synthetic_text_embeds(cfg, 7)
    tokens_in = {"embeds": txt}

    cap = _NumpyCapture()
    ref_vis.np = cap
    ups_inputs = []

    import tempfile
    tmp = Path(tempfile.mkdtemp())
    vis = ref_vis.Visualizations(clip, FakeAccelerator(), None, None, 1, tmp, "", None)
    vis.visualize_overlay = lambda *a, **k: None
    vis._upsample = lambda x, shape: (ups_inputs.append(x.detach().clone().numpy()
                                                        if isinstance(x, torch.Tensor) else np.array(x)),
                                      np.zeros((2, 2, 2), np.float32))[1]
    res = {"sd_checksum": sd_checksum(sd)}

    # ---- Grad-CAM (visualizations.py:913-1026) --------------------------------------
    t0 = time.time()
    clip.zero_grad()
    vis.visualize_grad_cam(img, tokens_in, None, "scan", "")
    # order of _upsample calls :995-1000
    for name, arr in zip(["spatial", "spatial_ff", "temporal", "temporal_ff", "combined", "vq"], ups_inputs):
        res["gradcam_" + name] = arr.astype(np.float32)
    # raw captured tensors for finer checks
    so = vis.saved_outputs
    res["gc_w_spatial_ff"] = so["spatial_ff_gradients"][-1].mean(dim=(0, 1)).numpy()
    res["gc_w_temporal_ff"] = so["temporal_ff_gradients"][-1].mean(dim=(0, 1)).numpy()
    res["gc_w_spatial"] = so["spatial_gradients"][-1].mean(dim=(0, 1)).numpy()
    res["gc_w_temporal"] = so["temporal_gradients"][-1].mean(dim=(0, 1)).numpy()
    res["gc_w_vq"] = so["vq_gradients"].squeeze(0).mean(dim=0).numpy()
    print(f"grad-cam {time.time() - t0:.1f}s")

    # ---- raw attention reductions (visualizations.py:659-676) on the same captured probs
    sp_attn = so["spatial_attention_weights"]
    tp_attn = so["temporal_attention_weights"]
    res["rawattn_spatial"] = np.stack([
        np.stack([a[:, hd].mean(dim=1).numpy() for a in sp_attn]) for hd in range(cfg.heads)])   # [H, L, 24, 576]
    res["rawattn_temporal"] = np.stack([
        np.stack([a[:, hd].mean(dim=1).numpy() for a in tp_attn]) for hd in range(cfg.heads)])   # [H, L, 576, 24]

    # ---- attention rollout (visualizations.py:779-849) --------------------------------
    t0 = time.time()
    ups_inputs.clear()
    clip.zero_grad()
    vis.visualize_attention_rollout(img, tokens_in, None, "scan", "")
    res["rollout_spatial"] = ups_inputs[0].astype(np.float32)     # [96,24,24]
    res["rollout_temporal"] = ups_inputs[1].astype(np.float32)    # [24,24,24]
    print(f"rollout {time.time() - t0:.1f}s")
    vis.saved_outputs.clear()

    # ---- integrated gradients, few steps (visualizations.py:851-910) -----------------
    t0 = time.time()
    cap.saved.clear(); cap.quantile_inputs.clear()
    clip.zero_grad()
    vis.visualize_integrated_gradients(img, tokens_in, None, "scan", "", steps=ig_steps)
    ig_final = cap.saved["scan.npy"]                 # rot90'd final map
    ig_pre = cap.quantile_inputs[0]                  # normalised, pre-threshold [D,H,W]
    res["ig_steps"] = ig_steps
    res["ig_pre_sub"] = ig_pre[::4, ::8, ::8].astype(np.float32)
    res["ig_pre_sum"] = float(ig_pre.astype(np.float64).sum())
    res["ig_pre_tokensum"] = ig_pre.reshape(24, 10, 24, 20, 24, 20).astype(np.float64).sum(axis=(1, 3, 5)).astype(np.float32)
    res["ig_q90"] = float(np.quantile(ig_pre, 0.90))
    res["ig_final_sub"] = np.rot90(ig_final, k=1, axes=(1, 2))[::4, ::8, ::8].astype(np.float32)
    res["ig_final_nonzero"] = int((ig_final > 0).sum())
    print(f"IG({ig_steps}) {time.time() - t0:.1f}s")

    # ---- occlusion, coarse windows (visualizations.py:335-424) ------------------------
    t0 = time.time()
    ps, st = (120, 240, 240), (120, 240, 240)
    heat = vis._compute_occlusion(img, None, txt, ps, st, 0.0)
    res["occ_patch"] = np.array(ps); res["occ_stride"] = np.array(st)
    res["occ_heat_sub"] = np.rot90(heat, k=1, axes=(1, 2))[5::10, 10::20, 10::20].astype(np.float32)
    # raw per-window scores, recomputed through the real model for the fixture
    with torch.no_grad():
        orig = float(clip(None, img, txt)[0][0, 0])
        scores = []
        for (d, h, w) in O.occlusion_windows((240, 480, 480), ps, st):
            occ = img.clone()
            occ[:, :, d:d + ps[0], h:h + ps[1], w:w + ps[2]] = -1
            scores.append(float(clip(None, occ, txt)[0][0, 0]))
    res["occ_orig"] = orig
    res["occ_scores"] = np.array(scores)
    print(f"occlusion {time.time() - t0:.1f}s  orig={orig:.6f} scores={scores}")

    np.savez_compressed(out, **res)
    print("full_attrib ->", out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default="all")
    ap.add_argument("--codebook", default="random", choices=["random", "fitted"],
                    help="fitted: the checkpoint whose codebook is tests/golden/fitted_codebook.npz "
                         "(make_fitted_codebook.py); writes *_fitted.npz")
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    refs = import_reference(args.ref)
    sfx = "" if args.codebook == "random" else "_fitted"
    if args.only in ("all", "tiny") and args.codebook == "random":
        make_tiny(refs, HERE / "tiny_model.npz")
    if args.only in ("all", "full"):
        make_full_forward(refs, HERE / f"full_forward{sfx}.npz", codebook=args.codebook)
    if args.only in ("all", "attrib"):
        make_full_attrib(refs, HERE / f"full_attrib{sfx}.npz", codebook=args.codebook)
    if args.only in ("all", "occrow"):
        make_occ_row(refs, HERE / f"occ_row{sfx}.npz", codebook=args.codebook)


if __name__ == "__main__":
    main()
