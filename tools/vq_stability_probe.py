"""How many VQ code assignments of an occluded forward are PROVABLY those of the un-occluded volume?
For unit-norm codes e, |x'.e - x.e| <= |x' - x|, so if the baseline top-2 score margin of a token exceeds
2 |x' - x| its arg-max cannot change.  Development probe for an incremental VQ search."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import numpy as np
import torch
from oracle import ctclip_oracle as O
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan

dev = torch.device("cuda")
eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), dev))
vol = O.synthetic_volume(O.FULL, 0).to(dev)
tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(dev))
base = eng.forward(vol, tl)
xb = base.x_pre_vq
cb = eng.plan.codebook.float().view(-1, 512)
sc = xb @ cb.t()
top2 = sc.topk(2, dim=-1).values
margin = (top2[:, 0] - top2[:, 1])
print("baseline margin quantiles", [float(margin.quantile(q)) for q in (0.01, 0.1, 0.5, 0.9)], "|x| mean", float(xb.norm(dim=-1).mean()))
cache = eng.occlusion_baseline(vol, tl)
cubes = [(0, 0, 0), (11, 11, 11), (22, 22, 22), (5, 12, 3), (18, 4, 20), (11, 0, 0), (2, 20, 11), (15, 15, 15)]
ctx = eng.forward_occluded(cache, cubes, (2, 2, 2), tl)
xo = ctx.x_pre_vq.view(len(cubes), -1, 512)
for i, c in enumerate(cubes):
    d = (xo[i] - xb).norm(dim=-1)
    safe = margin > 2 * d + 1e-4
    changed = ctx.indices.view(len(cubes), -1)[i] != base.indices
    print(f"cube {c}: |dx| quantiles {[round(float(d.quantile(q)), 4) for q in (0.1, 0.5, 0.9, 0.99)]}  provably unchanged {float(safe.float().mean()):.3f}"
          f"  actually changed {float(changed.float().mean()):.4f}  changed among 'safe' {int((changed & safe).sum())}")
