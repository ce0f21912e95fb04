"""Fit a VQ codebook to the encoder's own outputs and store it as a small fixture.

    python tests/golden/make_fitted_codebook.py          (CPU, ~1 min)

Why: the benchmark checkpoint is random-init and its codebook is 8192 random unit vectors that have
nothing to do with what the encoder emits: the top-1 cosine is 0.16 and 38 % of the tokens have a top-2 margin below
5e-3, so the hard arg-max (ctvit.py:117-118) turns last-bit rounding differences into different codes — under ANY
change of arithmetic, including the reference's own fp16 autocast.  A trained `ctclip_v2.pt` is not like that: its
codebook is the EMA / k-means fit of encoder outputs and tokens sit next to their code.  This script produces the
second seeded checkpoint the parity tests run on: the same seed-42 weights with the codebook replaced by a spherical
k-means fit (K = 8192, 10 Lloyd iterations, seeded) of the l2-normalised pre-VQ activations of synthetic volume 0
computed by the fp32 oracle.  Top-1 cosine 0.97, 2.7 % of margins below 5e-3.

The fit itself is not reproducible bit for bit across hosts (BLAS summation order), so the RESULT is the fixture:
rows are quantised to int8 with a per-row scale and the codebook is DEFINED as l2norm(int8 * scale) — every consumer
(make_golden.py here, the tests on the GPU box) decodes the same file to the same fp32 bits
(oracle.ctclip_oracle.fitted_codebook).
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import ctclip_oracle as O  # noqa: E402


def spherical_kmeans(x: torch.Tensor, K: int, iters: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    cb = x[torch.randperm(x.shape[0], generator=g)[:K]].clone()
    for _ in range(iters):
        a = (x @ cb.t()).argmax(1)
        new = torch.zeros_like(cb).index_add_(0, a, x)
        cnt = torch.bincount(a, minlength=K)
        new[cnt == 0] = cb[cnt == 0]
        cb = O.l2norm(new)
    return cb


def main():
    torch.set_num_threads(os.cpu_count())
    cfg = O.FULL
    sd = O.init_state_dict(cfg, 42)
    with torch.no_grad():
        tok = O.encode(O.patch_embed(O.synthetic_volume(cfg, 0), sd, cfg), sd, cfg)
    x = O.l2norm(tok.reshape(-1, cfg.dim))
    cb = spherical_kmeans(x, cfg.codebook_size, iters=10, seed=5)
    scale = cb.abs().amax(dim=1, keepdim=True) / 127.0
    q = torch.round(cb / scale).clamp_(-127, 127).to(torch.int8)
    out = HERE / "fitted_codebook.npz"
    np.savez_compressed(out, q=q.numpy(), scale=scale.squeeze(1).numpy().astype(np.float32),
                        fitted_on=np.array("oracle fp32 pre-VQ activations of synthetic volume 0, weights seed 42; "
                                           "spherical k-means K=8192, 10 iterations, seed 5"))
    dec = O.fitted_codebook(out)[0]
    top2 = (x @ dec.t()).topk(2, dim=-1).values
    m = top2[:, 0] - top2[:, 1]
    print(f"{out}: {out.stat().st_size / 1e6:.2f} MB; top-1 cosine median {float(top2[:, 0].median()):.3f}; "
          f"margin < 5e-3: {float((m < 5e-3).float().mean()):.4f}; quantisation cos "
          f"{float((dec * cb).sum(1).min()):.6f}")


if __name__ == "__main__":
    main()
