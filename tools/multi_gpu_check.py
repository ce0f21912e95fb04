"""torchrun --nproc-per-node N tools/multi_gpu_check.py : N-rank sharded occlusion + IG must reproduce the
single-rank result (NCCL all-reduce of window scores / IG partial sums)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
import torch.distributed as dist
from oracle import ctclip_oracle as O
from ctclip_b200 import attribution as A
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), dev))
vol = O.synthetic_volume(O.FULL, 0).to(dev)
tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(dev))
ps, st = (80, 160, 160), (80, 160, 160)          # 27 windows: 27 % 2 = 1 window dropped at world 2 (parity mode)

heat, aux = A.occlusion_sensitivity(eng, vol, tl, ps, st, batch=4, parity_sharding=True)
ig, iaux = A.integrated_gradients(eng, vol, tl, steps=6, batch=3)
# single-rank reference on every rank (same deterministic kernels)
windows = A.occlusion_windows(vol.shape[-3:], ps, st)
orig, scores = A.occlusion_scores(eng, vol, tl, windows, ps, batch=4)
kept = len(windows) // world * world
ok = abs(orig - aux["orig"]) == 0
ok &= bool(torch.equal(aux["scores"][:kept], scores[:kept])) and int(aux["included"].sum()) == kept
inc = torch.zeros(len(windows), dtype=torch.uint8, device=dev); inc[:kept] = 1
heat1 = A.occlusion_heatmap(orig, scores * inc, inc, vol.shape[-3:], ps, st)
ok &= bool(torch.equal(heat, heat1))
# multi-prompt sweep: one exchange carries the scores of all prompts
tl3 = eng.text_latents(torch.randn(3, 768, generator=torch.Generator().manual_seed(11)).to(dev))
heats, maux = A.occlusion_sensitivity_multi(eng, vol, tl3, ps, st, batch=4)
o3, s3 = A.occlusion_scores(eng, vol, tl3, windows, ps, batch=4, all_prompts=True)
ok &= bool(torch.equal(maux["scores"][:kept], s3[:kept])) and bool(torch.equal(maux["orig"], o3))
ok &= int(maux["included"].sum()) == kept and len(heats) == 3
ig1, iaux1 = A.integrated_gradients(eng, vol, tl, steps=6, batch=3, shard_steps=False)
rel = float((iaux["gsum"] - iaux1["gsum"]).abs().max() / iaux1["gsum"].abs().max())
ok &= rel < 1e-5 and bool(torch.equal(iaux["scores"], iaux1["scores"]))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} occlusion kept {kept}/{len(windows)} windows, IG partial-sum rel.diff {rel:.2e} ->",
          "PASS" if int(flag) else "FAIL")
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
