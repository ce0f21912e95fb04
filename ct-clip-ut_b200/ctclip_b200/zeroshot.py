"""Zero-shot pathology scoring on the CTViT engine — the caller next to the attribution path
(reference `src/utils/CTClipInference.py:133-201`).

The reference runs, per volume, one full CTCLIP forward (image tower + BERT) for EACH of the 18 pathologies with the
prompt pair ("There is X.", "There is no X."), then softmaxes the two logits (:158-180).  The image latent does not
depend on the prompt and the text latents do not depend on the volume, so here the 36 text latents are computed once
per model (`pair_text_latents`) and a volume costs ONE image forward with Bt = 36 (`zero_shot_probabilities`); the
per-pair softmax is a kernel (`ctc_pair_softmax`).  Results equal the reference loop's (tests/golden/zero_shot.npz,
tests/test_zeroshot.py).

Note: as committed the reference loop unpacks six values from CTCLIP.forward's five (:169) and raises before it
scores anything; the behaviour mirrored here is the one its own lines after the unpack define.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ._lib import call, stream_ptr
from .attribution import PATHOLOGIES
from .engine import Engine


def zero_shot_prompts(pathologies: Sequence[str] = PATHOLOGIES) -> List[str]:
    """Prompt list in the engine's pair order: row 2j = present, row 2j+1 = absent (CTClipInference.py:159-160)."""
    out: List[str] = []
    for p in pathologies:
        out += [f"There is {p}.", f"There is no {p}."]
    return out


def pair_text_latents(model, tokenizer, pathologies: Sequence[str] = PATHOLOGIES, device=None,
                      prompt_batch: int = 12) -> torch.Tensor:
    """l2norm(to_text_latent([CLS])) of the 2P prompts, fp32 [2P, dim_latent].  The text tower is the user's
    torch module; tokenisation arguments follow CTClipInference.py:159-165."""
    eng: Engine = model.engine(device)
    prompts = zero_shot_prompts(pathologies)
    cls = []
    with torch.no_grad():
        for s in range(0, len(prompts), prompt_batch):
            tok = tokenizer(prompts[s:s + prompt_batch], return_tensors="pt", padding="max_length", truncation=True,
                            max_length=512).to(eng.dev)
            cls.append(model.text_transformer(**tok).last_hidden_state[:, 0, :].float())
    return eng.text_latents(torch.cat(cls, dim=0))


def zero_shot_probabilities(engine: Engine, volume: torch.Tensor, pair_latents: torch.Tensor) -> torch.Tensor:
    """volume fp32 [B,1,D,H,W] (device), pair_latents fp32 [2P, d] from `pair_text_latents` ->
    float64 [B, P]: probability of "present" per pathology (softmax over each logit pair, :171-180)."""
    if pair_latents.dim() != 2 or pair_latents.shape[0] % 2:
        raise ValueError(f"pair_latents must be [2P, d] (present/absent rows interleaved), got {tuple(pair_latents.shape)}")
    B, P = volume.shape[0], pair_latents.shape[0] // 2
    out = torch.empty(B, P, dtype=torch.float64, device=engine.dev)
    if P == 0:
        return out
    ctx = engine.forward(volume, pair_latents, batch=B, save=False)
    call("ctc_pair_softmax", ctx.sim, B, P, out, stream_ptr())
    return out


def zero_shot(model, dataloader, tokenizer, pathologies: Sequence[str] = PATHOLOGIES, device=None,
              pair_latents: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """CTClipInference.zeroshot (:147-190) up to the gather: iterates `dataloader` (batches
    `(images, _, labels, _, _)`, labels wrapped in a batch list as the reference's dataset yields them),
    returns (predictions float64 [n, P], targets [n, P]) gathered over the process group in rank order
    (accelerator.gather_for_metrics, :187).  Metrics / plots (:193-201) are the caller's."""
    eng: Engine = model.engine(device)
    if pair_latents is None:
        pair_latents = pair_text_latents(model, tokenizer, pathologies, eng.dev)
    preds, targets = [], []
    for batch in dataloader:
        images, labels = batch[0], batch[2]
        images = images.to(eng.dev, torch.float32, non_blocking=True).contiguous()
        preds.append(zero_shot_probabilities(eng, images, pair_latents))
        if isinstance(labels, (list, tuple)):                       # the reference unwraps `labels[0]` at batch size 1
            labels = torch.stack([torch.as_tensor(l) for l in labels])
        targets.append(torch.as_tensor(labels).to(eng.dev).reshape(images.shape[0], -1))
    P = pair_latents.shape[0] // 2
    pred = torch.cat(preds) if preds else torch.empty(0, P, dtype=torch.float64, device=eng.dev)
    targ = torch.cat(targets) if targets else torch.empty(0, P, device=eng.dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        # ranks may hold different numbers of rows (an uneven loader split): exchange the counts first, pad to the
        # longest, gather, trim — what accelerator.gather_for_metrics does for the ragged tail
        w = dist.get_world_size()
        counts = torch.zeros(w, dtype=torch.int64, device=eng.dev)
        counts[dist.get_rank()] = pred.shape[0]
        dist.all_reduce(counts)
        counts = [int(c) for c in counts.tolist()]
        pad = max(counts) if counts else 0

        def gather(t):
            buf = t.new_zeros(pad, *t.shape[1:])
            buf[:t.shape[0]] = t
            parts = [torch.empty_like(buf) for _ in range(w)]
            dist.all_gather(parts, buf)
            return torch.cat([p[:c] for p, c in zip(parts, counts)])
        pred, targ = gather(pred), gather(targ)
    return pred, targ
