"""One fwd+bwd step inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
from oracle import ctclip_oracle as O
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda")
eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), dev))
vol = O.synthetic_volume(O.FULL, 0, batch=B).to(dev)
tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(dev))
for _ in range(2):
    ctx = eng.forward(vol, tl, save=True); eng.backward(ctx)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ctx = eng.forward(vol, tl, save=True); eng.backward(ctx)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(ctx.sim[0, 0]))
