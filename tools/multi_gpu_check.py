"""torchrun --nproc-per-node N tools/multi_gpu_check.py : N-rank sharded occlusion + IG must reproduce the
single-rank result (NCCL exchange of window scores / IG partial sums): ctclip_b200.selfcheck.sharding_parity on the
benchmark configuration.  `bench.py --gpus N` runs the same check and reports it as `parity`."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
import torch.distributed as dist
from oracle import ctclip_oracle as O
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan
from ctclip_b200.selfcheck import sharding_parity

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), dev))
vol = O.synthetic_volume(O.FULL, 0).to(dev)
tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(dev))
tl3 = eng.text_latents(torch.randn(3, 768, generator=torch.Generator().manual_seed(11)).to(dev))
res = sharding_parity(eng, vol, tl, tl3)
if dist.get_rank() == 0:
    print(json.dumps(res), "->", "PASS" if res["pass"] else "FAIL")
dist.destroy_process_group()
sys.exit(0 if res["pass"] else 1)
