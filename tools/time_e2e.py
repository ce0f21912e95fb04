"""Break the bench's end-to-end step into its parts on one GPU: pinned H2D alone, eager CTCLIP.forward + backward
alone (inputs resident), and both overlapped as bench.py runs them.  python tools/time_e2e.py"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
from ctclip_b200.modules import CTCLIP, CTViT  # noqa: E402

BATCH = 8
dev = torch.device("cuda", 0)
torch.manual_seed(42)
vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
            spatial_depth=4, temporal_depth=4, dim_head=32, heads=8)
clip = CTCLIP(text_encoder=torch.nn.Identity(), image_encoder=vit, dim_text=768, dim_image=294912, dim_latent=512)
clip.return_image_tokens = False
eng = clip.engine(dev)
host = (0.35 * torch.randn(BATCH, 1, 240, 480, 480) - 0.2).clamp_(-1, 1).pin_memory()
text = torch.randn(1, 768).to(dev)
bufs = [torch.empty(host.shape, device=dev) for _ in range(2)]


def wall(fn, n=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print(f"H2D alone (default stream)      {wall(lambda: bufs[0].copy_(host, non_blocking=True)):8.2f} ms")
cs = torch.cuda.Stream()


def h2d_side():
    with torch.cuda.stream(cs):
        bufs[1].copy_(host, non_blocking=True)


print(f"H2D alone (copy stream)         {wall(h2d_side):8.2f} ms")


def compute(parts=None):
    t = [time.perf_counter()]
    x = bufs[0].requires_grad_()
    sim, *_ = clip(None, x, text)
    if parts is not None:
        torch.cuda.synchronize(); t.append(time.perf_counter())
    sim[:, 0].sum().backward()
    if parts is not None:
        torch.cuda.synchronize(); t.append(time.perf_counter())
    out = torch.cat([sim.detach().flatten(), x.grad.square().sum(dim=(1, 2, 3, 4))])
    x.grad = None
    bufs[0].requires_grad_(False)
    r = out.cpu()
    if parts is not None:
        t.append(time.perf_counter())
        parts.append([(b - a) * 1e3 for a, b in zip(t, t[1:])])
    return r


print(f"eager fwd+bwd+readback alone    {wall(compute):8.2f} ms")
import gc


def stats():
    m = torch.cuda.memory_stats()
    return (m["num_device_alloc"], m["num_device_free"], m["num_alloc_retries"], sum(s["collections"] for s in gc.get_stats()),
            round(torch.cuda.memory_allocated() / 2**30, 1), round(torch.cuda.memory_reserved() / 2**30, 1))


for label in ("default", "gc disabled", "last_ctx dropped"):
    if label == "gc disabled":
        gc.collect(); gc.disable()
    parts = []
    print(f"-- {label}: [fwd, bwd, reduce+D2H] ms | cudaMalloc, cudaFree, retries, gc runs, allocated GiB, reserved GiB")
    for _ in range(6):
        s0 = stats()
        compute(parts)
        if label == "last_ctx dropped":
            clip.last_ctx = None
        s1 = stats()
        print("  ", [round(v, 2) for v in parts[-1]], "|", [b - a for a, b in zip(s0[:4], s1[:4])], s1[4:])
gc.enable()


def both():
    h2d_side()
    compute()


print(f"H2D (side stream) + compute     {wall(both):8.2f} ms")
# engine-level, no autograd
vol = bufs[0]
tl = eng.text_latents(text)


def eng_only():
    ctx = eng.forward(vol, tl, save=True)
    eng.backward(ctx)


print(f"engine.forward+backward eager   {wall(eng_only):8.2f} ms")
print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB")
