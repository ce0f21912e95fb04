"""The five attribution methods of src/utils/visualizations.py, re-implemented on explicit engine
outputs (no forward/tensor hooks) and sharded over GPUs.

Numerics follow the reference line by line, including its quirks:
  * occlusion: window enumeration d->h->w, contiguous `total // world` shards with the remainder
    dropped (visualizations.py:339-361), importance = max(orig - occ, 0) accumulated in float64;
  * rollout: spatial "rollout" never chains layers and yields a [layers*t, h, w] stack (:800-812);
  * Grad-CAM: LAST-layer features paired with FIRST-layer gradients (:929-934, hook order);
  * IG: baseline of ones, alpha = linspace(0, 1, steps), five-stage post-processing (:878-901).
What changes is the execution: perturbed volumes are never materialised (the occlusion cube and the
IG interpolation are fused into the patch-embedding load), windows / alpha steps are batched, the
text latent and the position-bias table are computed once, and the two 221 MB cross-rank reductions
of the occlusion path become an all-gather of per-window scores.
"""
from __future__ import annotations

import math
import time
from datetime import timedelta
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, stream_ptr
from .engine import Ctx, Engine

PATHOLOGIES = [
    "Medical material", "Arterial wall calcification", "Cardiomegaly", "Pericardial effusion",
    "Coronary artery wall calcification", "Hiatal hernia", "Lymphadenopathy", "Emphysema", "Atelectasis",
    "Lung nodule", "Lung opacity", "Pulmonary fibrotic sequela", "Pleural effusion", "Mosaic attenuation pattern",
    "Peribronchial thickening", "Consolidation", "Bronchiectasis", "Interlobular septal thickening"]


# ------------------------------------------------------------------------------------------- sharding
def occlusion_windows(shape, patch_size=(20, 40, 40), stride=(10, 20, 20)) -> List[Tuple[int, int, int]]:
    """visualizations.py:339-349 — nested d -> h -> w (w fastest)."""
    D, H, W = shape
    return [(d, h, w)
            for d in range(0, D - patch_size[0] + 1, stride[0])
            for h in range(0, H - patch_size[1] + 1, stride[1])
            for w in range(0, W - patch_size[2] + 1, stride[2])]


def shard_range(total: int, rank: int, world: int, parity: bool = True) -> Tuple[int, int]:
    """Index range [start, end) of this rank.  parity=True reproduces visualizations.py:351-361
    (`total // world` per rank, remainder dropped); parity=False is a balanced split covering all units."""
    if parity:
        per = total // world
        return rank * per, (rank + 1) * per
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_indices(total: int, rank: int, world: int, parity: bool = True) -> np.ndarray:
    """Window indices this rank evaluates.  The SET of evaluated windows is the reference's: with parity=True the
    first (total // world) * world windows (visualizations.py:351-361 truncates the list, then hands out contiguous
    chunks), else all of them.  Which rank evaluates which window does not enter any result (scores are independent
    and combined by index), so the windows are dealt round-robin: contiguous chunks give the ranks that own the
    first / last depth slabs only padding windows (no-ops) and clipped changed-frame sets, and the others wait."""
    n = (total // world) * world if parity else total
    return np.arange(rank, n, world, dtype=np.int64)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ------------------------------------------------------------------------------------------- small helpers
_PINNED: Dict[Tuple[int, torch.dtype, int], torch.Tensor] = {}


def to_host(t: torch.Tensor, slot: int = 0, view: bool = False) -> torch.Tensor:
    """Device -> host copy of a result map through a cached PINNED staging buffer (a pageable `.cpu()` of a
    221 MB map runs at a fraction of the PCIe rate).  Returns a tensor the caller OWNS.  view=True skips the
    host-side copy and returns a view of the staging buffer itself, valid only until the next to_host() call with
    the same element count, dtype and `slot` — for callers that consume the map at once (np.save)."""
    t = t.detach().contiguous()
    key = (t.numel(), t.dtype, slot)
    buf = _PINNED.get(key)
    if buf is None:
        if len(_PINNED) >= 4:
            _PINNED.clear()
        buf = _PINNED[key] = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
    out = buf.view(t.shape)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out if view else out.clone()


def _minmax(x: torch.Tensor) -> torch.Tensor:
    mm = torch.tensor([float("inf"), float("-inf")], device=x.device)
    call("ctc_minmax", x, x.numel(), mm, stream_ptr())
    return mm


def normalize(x: torch.Tensor, mode: int, rot90: bool = False) -> torch.Tensor:
    """mode 0: (v-min)/(max+1e-8); 1: (v-min)/(max-min+1e-8); 2: v/(max+1e-8)."""
    assert x.dim() == 3 and x.is_contiguous()
    D, H, W = x.shape
    out = torch.empty((D, W, H) if rot90 else (D, H, W), device=x.device)
    call("ctc_normalize", x, D, H, W, _minmax(x), mode, int(rot90), out, stream_ptr())
    return out


def upsample(x: torch.Tensor, shape, rot90: bool = True) -> torch.Tensor:
    """Visualizations._upsample (:289-293) + np.rot90(k=-1, axes=(1,2)) fused; returns a device tensor."""
    x = x.contiguous().float()
    d, h, w = x.shape
    D, H, W = shape
    out = torch.empty((D, W, H) if rot90 else (D, H, W), device=x.device)
    call("ctc_upsample_trilinear", x, d, h, w, out, D, H, W, int(rot90), stream_ptr())
    return out


def kth_value(x: torch.Tensor, k: int) -> float:
    """k-th smallest (0-based) element of a non-negative fp32 tensor, exact, by two 16-bit radix passes."""
    n = x.numel()
    hist = torch.empty(65536, dtype=torch.int32, device=x.device)
    call("ctc_hist16", x, n, 16, 0, hist, stream_ptr())
    cum = torch.cumsum(hist.to(torch.int64), 0)
    hi = int(torch.searchsorted(cum, torch.tensor([k + 1], device=x.device)))
    before = int(cum[hi - 1]) if hi > 0 else 0
    call("ctc_hist16", x, n, 0, hi, hist, stream_ptr())
    cum = torch.cumsum(hist.to(torch.int64), 0)
    lo = int(torch.searchsorted(cum, torch.tensor([k + 1 - before], device=x.device)))
    bits = np.array([(hi << 16) | lo], dtype=np.uint32)
    return float(bits.view(np.float32)[0])


def quantile_linear(x: torch.Tensor, q: float) -> float:
    """np.quantile(x, q) for an fp32 tensor, evaluated on the device and bit-identical to NumPy 2.x:
    NumPy casts a Python-float q to the array dtype, so the virtual index (n-1)*q, the interpolation
    weight and the lerp are all float32 (at n = 55 296 000 the fp32 index is already an integer)."""
    n = x.numel()
    qf = np.asanyarray(q, dtype=np.float32)
    vi = (n - 1) * qf
    prev = int(np.floor(vi))
    nxt = min(prev + 1, n - 1)
    gamma = np.asanyarray(vi - np.float32(prev), dtype=np.float32)
    a = np.float32(kth_value(x, prev))
    b = np.float32(kth_value(x, nxt)) if gamma > 0 else a
    d = np.subtract(b, a)
    r = np.add(a, d * gamma)
    if gamma >= 0.5:
        r = np.subtract(b, d * (1 - gamma))
    return float(np.float32(r))


def kth_value_dev(x: torch.Tensor, k: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """kth_value with the result left in device memory (fp32 [1]) and no host synchronisation."""
    ws = torch.empty(_lib.load().ctc_kth_value_ws_bytes() // 8 + 1, dtype=torch.int64, device=x.device)
    out = out if out is not None else torch.empty(1, device=x.device)
    call("ctc_kth_value", x, x.numel(), int(k), ws, out, stream_ptr())
    return out


def quantile_linear_dev(x: torch.Tensor, q: float) -> torch.Tensor:
    """quantile_linear evaluated entirely on the device: fp32 [1], bit-identical to NumPy 2.x (see quantile_linear),
    nothing read back by the host (the IG post-processing chain stays asynchronous)."""
    n = x.numel()
    qf = np.asanyarray(q, dtype=np.float32)
    vi = (n - 1) * qf
    prev = int(np.floor(vi))
    nxt = min(prev + 1, n - 1)
    gamma = float(np.asanyarray(vi - np.float32(prev), dtype=np.float32))
    a = kth_value_dev(x, prev)
    b = kth_value_dev(x, nxt) if gamma > 0 else a
    out = torch.empty(1, device=x.device)
    call("ctc_quantile_lerp", a, b, gamma, out, stream_ptr())
    return out


# ------------------------------------------------------------------------------------------- occlusion
def occlusion_scores(engine: Engine, volume: torch.Tensor, text_latents: torch.Tensor, windows: Sequence,
                     patch_size, batch: int = 8, fill: float = -1.0, reuse: Optional[bool] = None,
                     reuse_batch: int = 32, all_prompts: bool = False, skip_noop: bool = True,
                     stats: Optional[dict] = None):
    """Baseline score and one score per window (visualizations.py:370-388).  Returns (orig, scores fp32
    [len(windows)] on device).  Perturbed volumes are never materialised.

    all_prompts=True scores every row of text_latents [P, NL] from the SAME image forward (the image tower does
    not depend on the prompt): returns (orig fp32 [P], scores fp32 [len(windows), P]) — the reference runs one
    complete sweep per pathology prompt (visualizations.py:1037-1044).

    reuse=True (default whenever every window is aligned to the token grid, as the reference's
    (20,40,40)/(10,20,20) sweep is): Engine.forward_occluded — the patch embedding and every spatial-transformer
    frame the cube cannot reach come from the cached baseline; same arithmetic, ~40 % fewer executed FLOPs.
    reuse=False: dense path, the cube is applied inside the patch-embedding load of a full forward.
    skip_noop (reuse path only): a window whose patches are ALL already filled with `fill` (air / padding, exactly
    -1 after preprocess.py:135-147) leaves the volume unchanged, so its score is the baseline score bit for bit;
    such windows are detected on the device (ctc_patch_is_constant) and not evaluated.  `stats` receives
    {"evaluated", "noop"} counts."""
    dev = engine.dev
    cfg = engine.cfg
    tl = text_latents if all_prompts else text_latents[:1]
    scores = torch.empty(len(windows), tl.shape[0], device=dev)

    def result(orig):
        return (orig[0].clone(), scores) if all_prompts else (float(orig[0, 0]), scores[:, 0])
    tp, ps = cfg.temporal_patch_size, cfg.patch_size
    aligned = (patch_size[0] % tp == 0 and patch_size[1] % ps == 0 and patch_size[2] % ps == 0 and
               all(d % tp == 0 and h % ps == 0 and w % ps == 0 for (d, h, w) in windows))
    if reuse is None:
        reuse = aligned
    if reuse:
        if not aligned:
            raise ValueError("occlusion reuse needs windows aligned to the token grid")
        cache = engine.occlusion_baseline(volume, tl, fill)
        cubes = np.array([[d // tp, h // ps, w // ps] for (d, h, w) in windows], dtype=np.int64).reshape(-1, 3)
        shape = (patch_size[0] // tp, patch_size[1] // ps, patch_size[2] // ps)
        todo = np.arange(len(windows))
        if skip_noop and len(windows):
            D, H, W = volume.shape[-3:]
            flags = torch.empty(D // tp, H // ps, W // ps, dtype=torch.uint8, device=dev)
            call("ctc_patch_is_constant", volume, D, H, W, tp, ps, float(fill), flags, stream_ptr())
            f = flags.cpu().numpy().astype(bool)
            noop = np.ones(len(windows), dtype=bool)
            for a in range(shape[0]):
                for b in range(shape[1]):
                    for c in range(shape[2]):
                        noop &= f[cubes[:, 0] + a, cubes[:, 1] + b, cubes[:, 2] + c]
            todo = np.nonzero(~noop)[0]
            if noop.any():
                scores[torch.from_numpy(np.nonzero(noop)[0]).to(dev)] = cache.sim[0]
        if stats is not None:
            stats.update(evaluated=int(len(todo)), noop=int(len(windows) - len(todo)))
        for s in range(0, len(todo), reuse_batch):
            sel = todo[s:s + reuse_batch]
            sim = engine.forward_occluded(cache, cubes[sel], shape, tl).sim
            if len(sel) == sel[-1] - sel[0] + 1:
                scores[sel[0]:sel[-1] + 1] = sim
            else:
                scores[torch.from_numpy(sel).to(dev)] = sim
        return result(cache.sim)
    orig = engine.forward(volume, tl).sim
    wins = torch.tensor([[d, h, w, patch_size[0], patch_size[1], patch_size[2]] for (d, h, w) in windows],
                        dtype=torch.int32, device=dev).reshape(-1, 6)
    for s in range(0, len(windows), batch):
        e = min(s + batch, len(windows))
        ctx = engine.forward(volume, tl, batch=e - s, occl=wins[s:e].contiguous(), occl_value=fill)
        scores[s:e] = ctx.sim
    return result(orig)


def occlusion_heatmap(orig: float, scores: torch.Tensor, included: torch.Tensor, shape, patch_size, stride,
                      threshold: float = 0.0, rot90: bool = True) -> torch.Tensor:
    """visualizations.py:390-424 on the regular window grid.  scores/included are over ALL windows of the
    grid (enumeration order); `included` marks the ones that were evaluated."""
    D, H, W = shape
    nd = len(range(0, D - patch_size[0] + 1, stride[0]))
    nh = len(range(0, H - patch_size[1] + 1, stride[1]))
    nw = len(range(0, W - patch_size[2] + 1, stride[2]))
    imp = torch.clamp(orig - scores.float(), min=0).contiguous()           # importance = max(orig - occ, 0)
    inc = included.to(torch.uint8).contiguous()
    heat = torch.empty(D, H, W, device=scores.device)
    call("ctc_occlusion_heatmap", imp, inc, nd, nh, nw, *patch_size, *stride, D, H, W, heat, stream_ptr())
    out = normalize(heat, 1, rot90)          # (h-min)/(max-min+1e-8); the identity-size interpolate is a no-op
    if threshold > 0:
        out = torch.where(out < threshold, torch.zeros_like(out), out)
    return out


def combine_sharded(local: torch.Tensor, start: int, end: int, total: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Assemble per-unit values computed on disjoint index ranges across the ranks of the default group
    into one [total] vector + an `included` mask (units no rank evaluated stay 0 / excluded).  The supports
    are disjoint, so a SUM all-reduce is a gather; this is the only exchange of the occlusion path."""
    scores = torch.zeros(total, device=local.device, dtype=torch.float32)
    included = torch.zeros(total, dtype=torch.uint8, device=local.device)
    scores[start:end] = local
    included[start:end] = 1
    if _world()[1] > 1:
        dist.all_reduce(scores)
        inc32 = included.to(torch.int32)
        dist.all_reduce(inc32)
        included = inc32.to(torch.uint8)
    return scores, included


def combine_indexed(local: torch.Tensor, idx: np.ndarray, total: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """combine_sharded for an arbitrary (disjoint across ranks) index set per rank: local [len(idx)] or
    [len(idx), P] -> scores [total(, P)] + included mask [total]."""
    scores = torch.zeros((total, *local.shape[1:]), device=local.device, dtype=torch.float32)
    included = torch.zeros(total, dtype=torch.int32, device=local.device)
    if len(idx):
        ix = torch.from_numpy(np.asarray(idx, dtype=np.int64)).to(local.device)
        scores[ix] = local.float()
        included[ix] = 1
    if _world()[1] > 1:
        dist.all_reduce(scores)
        dist.all_reduce(included)
    return scores, included.to(torch.uint8)


def occlusion_sensitivity(engine: Engine, volume: torch.Tensor, text_latents: torch.Tensor,
                          patch_size=(20, 40, 40), stride=(10, 20, 20), batch: int = 8, parity_sharding: bool = True,
                          threshold: float = 0.0, rot90: bool = True, reuse: Optional[bool] = None,
                          skip_noop: bool = True):
    """_compute_occlusion (visualizations.py:335-424), sharded over the ranks of the default process group.
    Cross-rank exchange: ONE all-gather of per-window scores (<= 49 KB) instead of two 221 MB reduces."""
    rank, world = _world()
    D, H, W = volume.shape[-3:]
    windows = occlusion_windows((D, H, W), patch_size, stride)
    idx = shard_indices(len(windows), rank, world, parity_sharding)
    stats: dict = {}
    orig, local = occlusion_scores(engine, volume, text_latents, [windows[i] for i in idx], patch_size, batch,
                                   reuse=reuse, skip_noop=skip_noop, stats=stats)
    scores, included = combine_indexed(local, idx, len(windows))
    heat = occlusion_heatmap(orig, scores, included, (D, H, W), patch_size, stride, threshold, rot90)
    return heat, {"orig": orig, "scores": scores, "included": included, "windows": windows, "stats": stats}


def occlusion_sensitivity_multi(engine: Engine, volume: torch.Tensor, text_latents: torch.Tensor,
                                patch_size=(20, 40, 40), stride=(10, 20, 20), batch: int = 8,
                                parity_sharding: bool = True, threshold: float = 0.0, rot90: bool = True,
                                reuse: Optional[bool] = None):
    """One occlusion sweep shared by all P prompts of text_latents [P, NL]: the reference re-runs the whole sweep
    for every positive pathology (visualizations.py:1037-1044) although only the final dot product depends on the
    prompt.  Returns ([P heat maps], aux); each heat map equals occlusion_sensitivity() with that prompt alone."""
    rank, world = _world()
    D, H, W = volume.shape[-3:]
    windows = occlusion_windows((D, H, W), patch_size, stride)
    idx = shard_indices(len(windows), rank, world, parity_sharding)
    orig, local = occlusion_scores(engine, volume, text_latents, [windows[i] for i in idx], patch_size, batch,
                                   reuse=reuse, all_prompts=True)
    P = text_latents.shape[0]
    scores, included = combine_indexed(local, idx, len(windows))            # one exchange for all prompts
    heats = [occlusion_heatmap(float(orig[j]), scores[:, j].contiguous(), included, (D, H, W), patch_size, stride,
                               threshold, rot90) for j in range(P)]
    return heats, {"orig": orig, "scores": scores if P else None, "included": included, "windows": windows}


# ------------------------------------------------------------------------------------------- integrated gradients
def integrated_gradients(engine: Engine, volume: torch.Tensor, text_latents: torch.Tensor, steps: int = 50,
                         batch: int = 5, shard_steps: bool = True, rot90: bool = True):
    """visualize_integrated_gradients (visualizations.py:851-901).  The alpha steps are batched, the
    interpolation 1 + alpha (x - 1) is applied inside the patch-embedding load and the per-step gradients
    are summed on the fly into one [D,H,W] buffer (the reference keeps 50 x 221 MB).
    shard_steps=True: EVERY rank of the default process group must hold the SAME volume; the alpha steps are dealt
    over the ranks and the partial sums combined with one NCCL reduce to rank 0, which alone runs the
    post-processing (the other ranks return (None, aux)).  shard_steps=False: this rank attributes its own volume
    (what the reference does, and what `Visualizations.visualize` needs: its loader hands every rank a different scan).
    The post-processing chain (relu, min/max, exact 0.90 quantile, ** 0.05, rescale, rot90) runs on the device
    without a host synchronisation."""
    rank, world = _world() if shard_steps else (0, 1)
    dev = engine.dev
    D, H, W = volume.shape[-3:]
    alphas = torch.linspace(0, 1, steps, device=dev)
    mine = np.arange(rank, steps, world)             # round-robin: every rank gets floor or ceil(steps / world)
    gsum = torch.zeros(D, H, W, device=dev)
    scores = torch.zeros(steps, device=dev)
    tl = text_latents[:1]
    for s in range(0, len(mine), batch):
        sel = torch.from_numpy(mine[s:s + batch]).to(dev)
        ctx = engine.forward(volume, tl, batch=len(sel), alpha=alphas[sel].contiguous(), save=True)
        scores[sel] = ctx.sim[:, 0]
        engine.backward(ctx, grad_out=gsum, sum_over_batch=True)
        del ctx
    if world > 1:
        dist.reduce(gsum, dst=0)
        dist.reduce(scores, dst=0)
        if rank != 0:
            return None, {"scores": scores, "gsum": gsum}
    n = D * H * W
    ig = torch.empty(D, H, W, device=dev)
    mm = torch.tensor([float("inf"), float("-inf")], device=dev)
    call("ctc_ig_combine", volume, gsum, n, 1.0 / steps, ig, mm, stream_ptr())     # relu((x-1) * mean grad)
    pre = torch.empty_like(ig)                                                     # (ig-min)/(max+1e-8)
    call("ctc_normalize", ig, D, H, W, mm, 0, 0, pre, stream_ptr())
    q = quantile_linear_dev(pre, 0.90)
    out = torch.empty((D, W, H) if rot90 else (D, H, W), device=dev)
    call("ctc_ig_finalize_dev", ig, D, H, W, mm, q, int(rot90), out, stream_ptr())
    return out, {"pre_threshold": pre, "q90": q, "scores": scores, "gsum": gsum}


# ------------------------------------------------------------------------------------------- Grad-CAM
def grad_cam(engine: Engine, volume: torch.Tensor, text_latents: torch.Tensor) -> Dict[str, torch.Tensor]:
    """visualize_grad_cam (visualizations.py:913-991): six 24^3 CAMs (pre-upsample, un-rotated).
    Features = LAST layer's attention / FF module output (difference of saved residual streams);
    gradients = FIRST layer's (the reference's hook-order quirk)."""
    cfg = engine.cfg
    ctx = engine.forward(volume, text_latents[:1], save=True, want_tokens=True)
    engine.backward(ctx, capture_grads=True, to_input=False)
    T, H, C = ctx.T, cfg.hw, cfg.dim
    R = T * H * H

    ws = torch.empty(_lib.load().ctc_colmean_ws_floats(R, C), device=engine.dev)

    def cam(fa, fb, grad):
        w = torch.empty(C, device=engine.dev)
        call("ctc_colmean", grad, R, C, w, ws, stream_ptr())           # two-stage, fixed order: bit-reproducible
        out = torch.empty(R, device=engine.dev)
        call("ctc_gradcam", fa, fb, w, R, C, out, stream_ptr())
        return out.view(T, H, H)

    ls, lt = ctx.spatial[-1], ctx.temporal[-1]
    g = ctx.grads
    maps = {
        "spatial_ff": cam(ctx.x_s_last, ls.x2, g["spatial0_ff"]),
        "temporal_ff": cam(ctx.x_t_last, lt.x2, g["temporal0_ff"]),
        "spatial": cam(ls.x2, ls.x1, g["spatial0_attn"]),
        "temporal": cam(lt.x2, lt.x1, g["temporal0_attn"]),
        "vq": cam(ctx.tokens, None, g["vq"].contiguous()),
    }
    # rows are already in canonical (t, h, w) order, which is what the reference obtains for the temporal
    # kinds after view(h, w, t).permute(2, 0, 1) (:944, :968)
    maps = {k: normalize(v.contiguous(), 0) for k, v in maps.items()}
    maps["combined"] = torch.sqrt(maps["spatial"] * maps["temporal"] + 1e-8)
    maps["_sim"] = ctx.sim
    return maps


# ------------------------------------------------------------------------------------------- attention maps
def attention_rollout_maps(engine: Engine, volume: torch.Tensor, text_latents: torch.Tensor):
    """visualize_attention_rollout (visualizations.py:779-841), pre-upsample: spatial [layers*t, h, w]
    (layer-major stack of single-matrix rollouts), temporal [t, h, w].  The attention kernels emit the head-mean
    probabilities directly (Engine.attention_fused: 32 MB per spatial layer); the [b, heads, n, n] tensors the
    reference's hooks retain are never materialised."""
    cfg = engine.cfg
    ctx = engine.forward(volume, text_latents[:1], keep_attn=True)
    T, H = ctx.T, cfg.hw
    n = H * H
    rows = []
    for layer in range(cfg.spatial_depth):
        fused, _ = engine.attention_fused(ctx, "spatial", layer, fused=True)       # [T, n, n] head mean
        out = torch.empty(T, n, device=engine.dev)
        call("ctc_rollout_spatial", fused, T, 1, n, out, stream_ptr())
        rows.append(out.view(T, H, H))
        del fused
    spatial = normalize(torch.cat(rows, 0).contiguous(), 1)
    tp = torch.stack([engine.attention_fused(ctx, "temporal", l, fused=True)[0]
                      for l in range(cfg.temporal_depth)]).contiguous()            # [L, n, T, T]
    out = torch.empty(n, T, device=engine.dev)
    call("ctc_rollout_temporal", tp, cfg.temporal_depth, n, 1, T, out, stream_ptr())
    temporal = normalize(out.view(H, H, T).permute(2, 0, 1).contiguous(), 1)
    return spatial, temporal


def raw_attention_maps(engine: Engine, volume: torch.Tensor, text_latents: torch.Tensor):
    """visualize_attention_grid_gif reductions (visualizations.py:659-676): per head, per layer the mean
    over the query axis, as 24^3 volumes.  Returns (spatial, temporal), each [heads, layers, D, H, W]
    normalised with (v-min)/(max+1e-8) (the final np.rot90(axes=(0,1)) is left to the renderer)."""
    cfg = engine.cfg
    ctx = engine.forward(volume, text_latents[:1], keep_attn=True)
    T, H = ctx.T, cfg.hw
    n = H * H
    sp, tp = [], []
    for layer in range(cfg.spatial_depth):
        _, cm = engine.attention_fused(ctx, "spatial", layer, colmean=True)        # [T, heads, n]
        sp.append(cm.permute(1, 0, 2).reshape(cfg.heads, T, H, H))
    for layer in range(cfg.temporal_depth):
        _, cm = engine.attention_fused(ctx, "temporal", layer, colmean=True)       # [n, heads, T]
        tp.append(cm.permute(1, 0, 2).reshape(cfg.heads, H, H, T).permute(0, 3, 1, 2))
    norm = lambda v: torch.stack([torch.stack([normalize(v[l][h].contiguous(), 0) for l in range(len(v))])
                                  for h in range(cfg.heads)])
    return norm(sp), norm(tp)


def attention_rollout(attn_weights_list: Sequence[torch.Tensor], head_fusion="mean", discard_ratio=0.0,
                      use_residual=True) -> torch.Tensor:
    """Visualizations.attention_rollout (visualizations.py:707-743) for any list of [heads, n, n] CUDA tensors,
    every option included (head_fusion 'mean' / 'max', discard_ratio, use_residual), on two kernels per layer:
    ctc_rollout_fuse (fusion, exact per-row top-k threshold, the two row normalisations) and ctc_matmul_f32
    (result = attn @ result).  Returns the [n, n] rollout matrix."""
    if head_fusion not in ("mean", "max"):
        raise ValueError(f"Unsupported head_fusion: {head_fusion}")
    first = attn_weights_list[0]
    if not first.is_cuda:
        raise RuntimeError("ctclip_b200: attention_rollout runs on CUDA tensors only (there is no CPU path)")
    n = first.size(-1)
    result = torch.eye(n, device=first.device)
    k_keep = n - int(n * discard_ratio) if discard_ratio > 0 else n
    for attn in attn_weights_list:
        attn = attn.detach().float().contiguous()
        if attn.dim() != 3 or attn.shape[-2:] != (n, n):
            raise ValueError(f"attention_rollout: expected [heads, {n}, {n}] matrices, got {tuple(attn.shape)}")
        a = torch.empty(n, n, device=attn.device)
        call("ctc_rollout_fuse", attn, attn.shape[0], n, int(head_fusion == "max"), max(k_keep, 1), int(use_residual),
             a, stream_ptr())
        nxt = torch.empty(n, n, device=attn.device)
        call("ctc_matmul_f32", a, result, n, nxt, stream_ptr())
        result = nxt
    return result


# ------------------------------------------------------------------------------------------- Visualizations
class Visualizations:
    """Drop-in for utils.visualizations.Visualizations (visualizations.py:73-1195): same constructor,
    `visualize(**flags)` dispatcher, per-method entry points and output `.npy` names.  GIF rendering
    (matplotlib, :427-567) is out of scope: overlays are skipped, arrays are always saved."""

    def __init__(self, model, accelerator, dataset, dist_dataloader, batch_size, results_folder, diff_embeds_folder,
                 tokenizer, window_batch: int = 8, ig_batch: int = 10, parity_sharding: bool = True):
        self.model = model.module if hasattr(model, "module") else model
        self.accelerator = accelerator
        self.dataset, self.dist_dataloader, self.batch_size = dataset, dist_dataloader, batch_size
        self.tokenizer = tokenizer
        self.results_folder = Path(results_folder)
        self.diff_embeds_folder = diff_embeds_folder
        self.rank = accelerator.process_index
        self.world_size = accelerator.num_processes
        self.maybe_print = print if accelerator.is_main_process else (lambda *a, **k: None)
        self.window_batch, self.ig_batch, self.parity_sharding = window_batch, ig_batch, parity_sharding
        self.saved_outputs: Dict[str, object] = {}

    # -- helpers -----------------------------------------------------------------------------
    def _results_subdirectory(self, name):
        sub = self.results_folder / name
        sub.mkdir(parents=True, exist_ok=True)
        idx = len([d for d in sub.iterdir() if d.is_dir()]) + 1
        sub = sub / str(idx)
        sub.mkdir(parents=True, exist_ok=True)
        return sub

    def _engine(self) -> Engine:
        return self.model.engine(self.accelerator.device)

    def _text_latents(self, text_tokens, text_embeds=None) -> torch.Tensor:
        """Text tower output -> latent, ONCE per (volume, prompt); the reference re-runs BERT on every
        window / step (ctclip.py:107)."""
        eng = self._engine()
        if isinstance(text_embeds, torch.Tensor) and text_embeds.ndim > 1:
            e = text_embeds
        else:
            with torch.no_grad():
                e = self.model.text_transformer(**text_tokens).last_hidden_state[:, 0, :]
        return eng.text_latents(e.detach())

    def _upsample(self, x, target_shape):
        return upsample(x, target_shape, rot90=False).cpu().numpy()

    def _save(self, path, device_array):
        np.save(path, to_host(device_array, view=True).numpy())     # consumed at once: no second host copy

    def visualize_overlay(self, *a, **k):
        return None  # GIF rendering is out of scope

    attention_rollout = staticmethod(attention_rollout)

    # -- methods -----------------------------------------------------------------------------
    def visualize_raw_attention_maps(self, image, text_tokens, labels, scan_name, original_scan_path):
        sp, tp = raw_attention_maps(self._engine(), image.float().contiguous(), self._text_latents(text_tokens))
        self.saved_outputs["raw_attention"] = (sp, tp)
        if self.accelerator.is_main_process:
            d = self._results_subdirectory("raw_attention_grids")
            self._save(d / f"{scan_name}_spatial_grid.npy", sp)
            self._save(d / f"{scan_name}_temporal_grid.npy", tp)

    def visualize_attention_rollout(self, image, text_tokens, labels, scan_name, original_scan_path):
        image = image.float().contiguous()
        shape = tuple(image.shape[-3:])
        sp, tp = attention_rollout_maps(self._engine(), image, self._text_latents(text_tokens))
        volume, temporal_vol = upsample(sp, shape), upsample(tp, shape)
        self.saved_outputs["rollout"] = (sp, tp)
        if self.accelerator.is_main_process:
            d = self._results_subdirectory("attention_rollout")
            self._save(d / f"{scan_name}_spatial.npy", volume)
            self._save(d / f"{scan_name}_temporal.npy", temporal_vol)

    def visualize_integrated_gradients(self, image, text_tokens, labels, scan_name, original_scan_path, steps=50):
        # shard_steps=False: `visualize` feeds this method from the DistributedSampler loader, i.e. every rank holds a
        # DIFFERENT scan and attributes it alone, as in the reference (sharding the alpha steps here would sum the
        # gradients of different volumes).  Latency mode for ONE volume: integrated_gradients(shard_steps=True).
        ig, aux = integrated_gradients(self._engine(), image.float().contiguous(), self._text_latents(text_tokens),
                                       steps=steps, batch=self.ig_batch, shard_steps=False)
        self.saved_outputs["integrated_gradients"] = aux
        if self.accelerator.is_main_process:
            d = self._results_subdirectory("integrated_gradients")
            self._save(d / f"{scan_name}.npy", ig)

    def visualize_grad_cam(self, image, text_tokens, labels, scan_name, original_scan_path):
        image = image.float().contiguous()
        shape = tuple(image.shape[-3:])
        maps = grad_cam(self._engine(), image, self._text_latents(text_tokens))
        self.saved_outputs["grad_cam"] = maps
        if self.accelerator.is_main_process:
            d = self._results_subdirectory("grad_cam")
            for k in ("spatial_ff", "temporal_ff", "spatial", "temporal", "combined", "vq"):
                self._save(d / f"{scan_name}_{k}.npy", upsample(maps[k], shape))

    def _compute_occlusion(self, image, text_tokens, text_embeds, patch_size, stride, threshold):
        heat, aux = occlusion_sensitivity(self._engine(), image.float().contiguous(),
                                          self._text_latents(text_tokens, text_embeds), patch_size, stride,
                                          self.window_batch, self.parity_sharding, threshold)
        self.saved_outputs["occlusion"] = aux
        return to_host(heat).numpy() if self.accelerator.is_main_process else None   # an owned host copy

    def visualize_occlusion_sensitivity(self, image, text_tokens, labels, scan_name, original_scan_path,
                                        patch_size=(20, 40, 40), stride=(10, 20, 20), use_text_embeds=False, prompt=""):
        threshold = 0.0
        heatmaps = {}
        if use_text_embeds:
            emb = np.load(self.diff_embeds_folder, allow_pickle=True).item()
            tens = {k: torch.tensor(v, dtype=torch.float32, device=self.accelerator.device).unsqueeze(0)
                    for k, v in emb.items()}
            pos = (labels == 1).nonzero(as_tuple=True)[0]
            names = [PATHOLOGIES[i] for i in pos.tolist()]
            if names:
                # one sweep for all positive pathologies (the reference runs one sweep each, :1037-1044)
                self.maybe_print("Processing pathologies:", ", ".join(names))
                tl = self._text_latents(text_tokens, torch.cat([tens[n] for n in names], dim=0))
                heats, aux = occlusion_sensitivity_multi(self._engine(), image.float().contiguous(), tl, patch_size,
                                                         stride, self.window_batch, self.parity_sharding, threshold)
                self.saved_outputs["occlusion"] = aux
                for n, hm in zip(names, heats):
                    heatmaps[n] = hm.cpu().numpy() if self.accelerator.is_main_process else None
        else:
            heatmap = self._compute_occlusion(image, text_tokens, None, patch_size, stride, threshold)
        if self.accelerator.is_main_process:
            d = self._results_subdirectory("occlusion")
            if use_text_embeds:
                np.save(d / f"{scan_name}_{str(patch_size)}_{str(stride)}_{prompt}_heatmaps.npy", heatmaps)
            else:
                np.save(d / f"{scan_name}_{prompt}_heatmap.npy", heatmap)

    # -- dispatcher (visualizations.py:1085-1195) ------------------------------------------------
    def visualize(self, **kwargs):
        for name, val in kwargs.items():
            if not val:
                continue
            if self.world_size > 1:
                dist.barrier()
            start = time.time()
            funcs = {"raw_attention_maps": self.visualize_raw_attention_maps,
                     "attention_rollout": self.visualize_attention_rollout,
                     "integrated_gradients": self.visualize_integrated_gradients,
                     "grad_cam": self.visualize_grad_cam}
            if name in funcs:
                self.maybe_print(f"{name} visualization started.")
                for batch in self.dist_dataloader:
                    image, texts, labels, scan_names, paths = [
                        b.to(self.accelerator.device) if isinstance(b, torch.Tensor) else b for b in batch]
                    tokens = self.tokenizer(texts, return_tensors="pt", padding="max_length", truncation=True,
                                            max_length=512).to(self.accelerator.device)
                    funcs[name](image, tokens, labels[0], scan_names[0], paths[0])
            elif name == "occlusion":
                self.maybe_print("Occlusion visualization started.")
                for idx in range(len(self.dataset)):
                    sample = self.dataset[idx] if self.rank == 0 else None
                    sample = self._broadcast_sample(sample)
                    image, texts, labels, scan_names, paths = sample
                    tokens = self.tokenizer(texts, return_tensors="pt", padding="max_length", truncation=True,
                                            max_length=512).to(self.accelerator.device)
                    self.visualize_occlusion_sensitivity(image, tokens, labels[0], scan_names[0], paths[0],
                                                         use_text_embeds=False)
            else:
                self.maybe_print(f"{name} is not a valid visualization argument.")
                return
            self.maybe_print(f"{name} visualization completed. Time: {timedelta(seconds=time.time() - start)}")

    def _broadcast_sample(self, sample, src=0):
        """visualizations.py:296-318: rank 0 loads, everyone receives (one 221 MB NCCL broadcast)."""
        dev = self.accelerator.device
        if self.world_size == 1:
            return [x.unsqueeze(0).to(dev) if isinstance(x, torch.Tensor) else [x] for x in sample]
        obj = [None]
        if self.rank == src:
            obj[0] = [(tuple(x.shape), x.dtype) if isinstance(x, torch.Tensor) else x for x in sample]
        dist.broadcast_object_list(obj, src=src)
        out = []
        for i, meta in enumerate(obj[0]):
            if isinstance(meta, tuple) and len(meta) == 2 and isinstance(meta[1], torch.dtype):
                t = (sample[i].unsqueeze(0).to(dev).contiguous() if self.rank == src
                     else torch.empty((1, *meta[0]), dtype=meta[1], device=dev))
                dist.broadcast(t, src=src)
                out.append(t)
            else:
                out.append([meta])
        return out
