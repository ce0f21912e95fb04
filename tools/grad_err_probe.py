"""Input-gradient max relative error vs the oracle (same VQ codes) with individual kernel choices toggled."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
from oracle import ctclip_oracle as O
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
cfg = O.FULL
sd = O.init_state_dict(cfg, 42)
eng = Engine(Plan(sd, Config(), dev))
sdd = O.to_device(sd, dev)
vol = O.synthetic_volume(cfg, 0).to(dev)
txt = O.synthetic_text_embeds(cfg, 7).to(dev)
tl = eng.text_latents(txt)
alpha = torch.tensor([0.5], device=dev)
for tc in (True, False):
    eng.attn_tc = tc
    ctx = eng.forward(vol, tl, alpha=alpha, save=True)
    grad = eng.backward(ctx)
    torch.cuda.synchronize()
    xa = (1 + 0.5 * (vol - 1)).detach().requires_grad_()
    sim = O.ctclip_forward(xa, txt, sdd, cfg, None, force_indices=ctx.indices)[0]
    (g_ref,) = torch.autograd.grad(sim[0, 0], xa)
    d = (grad - g_ref).abs()
    rel = float(d.max() / g_ref.abs().max())
    rms = float(d.square().mean().sqrt() / g_ref.square().mean().sqrt())
    q = torch.quantile(d.flatten()[::64].float(), torch.tensor([0.5, 0.99, 0.9999], device=dev)) / g_ref.abs().max()
    print(f"attn_tc={tc}: sim {float(ctx.sim):.6f} vs {float(sim):.6f}  max rel.err {rel:.3e}  rms rel.err {rms:.3e}  "
          f"|err|/max|g| quantiles 50/99/99.99%: {[f'{float(v):.2e}' for v in q]}")
