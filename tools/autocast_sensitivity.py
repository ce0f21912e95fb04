"""How far does the reference's OWN reduced-precision path move the logit?  Runs the oracle (the reference arithmetic)
at the benchmark configuration on the host in fp32 and under torch.autocast(bf16) — the CPU analogue of the fp16
autocast the reference runs under Accelerate (CTClipInference.py:56-63) — and reports VQ code agreement and the
unconditional / code-conditioned logit differences.  CPU only (about two minutes).
    python tools/autocast_sensitivity.py"""
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ctclip_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count())
cfg = O.FULL
sd = O.init_state_dict(cfg, 42)
txt = O.synthetic_text_embeds(cfg, 7)
DT = torch.float16 if "--fp16" in sys.argv else torch.bfloat16
for vol_seed in ((0,) if "--fp16" in sys.argv else (0, 1)):
    vol = O.synthetic_volume(cfg, vol_seed)
    with torch.no_grad():
        t0 = time.time()
        cap = {}
        sim32, _, _, _, _, ind32 = O.ctclip_forward(vol, txt, sd, cfg, cap)
        pre32 = cap["pre_vq"]
        with torch.autocast("cpu", dtype=DT):
            cap = {}
            sim16, _, _, _, _, ind16 = O.ctclip_forward(vol, txt, sd, cfg, cap)
            pre16 = cap["pre_vq"].float()
        simc = O.ctclip_forward(vol, txt, sd, cfg, None, force_indices=ind16)[0]
    agree = float((ind32 == ind16).float().mean())
    rel = float((pre16 - pre32).abs().max() / pre32.abs().max())
    print(f"volume {vol_seed}: fp32 logit {float(sim32):+.6f}  {str(DT)[6:]}-autocast logit {float(sim16):+.6f}  "
          f"|diff| {abs(float(sim32) - float(sim16)):.2e};  VQ code agreement {agree:.4f};  pre-VQ rel.err {rel:.2e};  "
          f"fp32 logit with the autocast run's codes {float(simc):+.6f} (|diff to autocast| "
          f"{abs(float(simc) - float(sim16)):.2e})   [{time.time() - t0:.0f} s]")
