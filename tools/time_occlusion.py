"""Device timing + kernel table of the occlusion fast path (Engine.forward_occluded); development aid."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import numpy as np
import torch
from oracle import ctclip_oracle as O
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan

dev = torch.device("cuda")
Wn = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), dev))
vol = O.synthetic_volume(O.FULL, 0).to(dev)
tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(dev))
cache = eng.occlusion_baseline(vol, tl)
rng = np.random.default_rng(0)
cubes = np.stack([rng.integers(0, 23, Wn), rng.integers(0, 23, Wn), rng.integers(0, 23, Wn)], axis=1)

def run():
    return eng.forward_occluded(cache, cubes, (2, 2, 2), tl).sim

for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"forward_occluded({Wn} windows): {ms:.3f} ms -> {ms / Wn:.4f} ms/window")
wins = torch.tensor([[10 * int(c[0]), 20 * int(c[1]), 20 * int(c[2]), 20, 40, 40] for c in cubes[:8]], dtype=torch.int32, device=dev)
def dense():
    return eng.forward(vol, tl, batch=8, occl=wins).sim
for _ in range(2): dense()
torch.cuda.synchronize()
e0.record()
for _ in range(5): dense()
e1.record(); torch.cuda.synchronize()
print(f"dense forward(8 windows): {e0.elapsed_time(e1) / 5:.3f} ms -> {e0.elapsed_time(e1) / 40:.4f} ms/window")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
