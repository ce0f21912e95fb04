"""Quick device timing of the engine (CUDA events); development aid, not the bench."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
from oracle import ctclip_oracle as O
from ctclip_b200.engine import Engine
from ctclip_b200.plan import Config, Plan
from ctclip_b200 import _lib

dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), dev))
vol = O.synthetic_volume(O.FULL, 0, batch=B).to(dev)
tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(dev))

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

ms, wall = timeit(lambda: eng.forward(vol, tl))
print(f"B={B} forward: {ms:.3f} ms device ({wall:.3f} ms wall)  -> {B/ms*1e3:.1f} vol/s, {0.7896*B/ms:.1f} TFLOP/s")
def fb():
    ctx = eng.forward(vol, tl, save=True); eng.backward(ctx)
ms, wall = timeit(fb, n=3, warm=1)
print(f"B={B} fwd+bwd: {ms:.3f} ms device ({wall:.3f} ms wall) -> {B/ms*1e3:.2f} vol/s, {1.497*B/ms:.1f} TFLOP/s")
print("peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
