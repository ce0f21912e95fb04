"""Run the same forward(+backward) several times and report the first tensor that differs (debug aid)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
from oracle import ctclip_oracle as O
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan

dev = torch.device("cuda")
eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), dev))
vol = O.synthetic_volume(O.FULL, 0).to(dev)
tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(dev))

def snapshot(save):
    ctx = eng.forward(vol, tl, save=save, keep_attn=True)
    out = {"x_pre_vq": ctx.x_pre_vq, "indices": ctx.indices, "pooled": ctx.pooled, "latent": ctx.latent, "sim": ctx.sim}
    if save:
        out["x_lin"] = ctx.x_lin; out["x_in"] = ctx.x_in
    for kind, layers in (("s", ctx.spatial), ("t", ctx.temporal)):
        for i, lc in enumerate(layers):
            for f in ("x1", "q", "kv", "o", "lse", "x2", "u"):
                v = getattr(lc, f)
                if v is not None:
                    out[f"{kind}{i}.{f}"] = v
    if save:
        out["grad"] = eng.backward(ctx)
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in out.items()}, ctx

for save in (False, True):
    ref, _ = snapshot(save)
    for rep in range(3):
        cur, _ = snapshot(save)
        bad = [k for k in ref if not torch.equal(ref[k], cur[k])]
        order = [k for k in ref]
        print(f"save={save} rep={rep} sim={float(cur['sim']):.6f} mismatching: {bad[:12]}")
        if bad:
            first = [k for k in order if k in bad][0]
            d = (ref[first].float() - cur[first].float()).abs()
            print("   first:", first, "max abs diff", float(d.max()), "n diff", int((d > 0).sum()), "of", d.numel())
