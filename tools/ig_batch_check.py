"""The IG partial sum must not depend on how the alpha steps are batched (development aid; the N-rank version of this
check is ctclip_b200.selfcheck.sharding_parity).  python tools/ig_batch_check.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
from oracle import ctclip_oracle as O
from ctclip_b200 import attribution as A
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan

dev = torch.device("cuda")
eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), dev))
vol = O.synthetic_volume(O.FULL, 0).to(dev)
tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(dev))
ref = None
for batch in (1, 3, 2, 6, 3):
    _, aux = A.integrated_gradients(eng, vol, tl, steps=6, batch=batch, shard_steps=False)
    g = aux["gsum"].clone()
    if ref is None:
        ref = g
    print(f"batch {batch}: |gsum| max {float(g.abs().max()):.4e}  rel diff to batch 1: {float((g - ref).abs().max() / ref.abs().max()):.3e}  scores {aux['scores'].flatten().tolist()[:6]}")
