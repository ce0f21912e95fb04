"""Device time of ctc_preprocess_ct on a scan that is already at the target spacing / shape (the bench's int16 ingest) and on a
typical CT-RATE geometry (512 x 512 x 300 at 0.7 mm / 1.0 mm); development aid."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
from ctclip_b200.preprocess import process_volume
dev = torch.device("cuda")
for name, shape, xy, z in (("identity 480x480x240", (240, 480, 480), 0.75, 1.5), ("ct-rate-like 512x512x300", (300, 512, 512), 0.7, 1.0)):
    D0, H0, W0 = shape
    raw = torch.randint(-1000, 1000, (D0, W0, H0), dtype=torch.int16, device=dev)      # NIfTI file order: H fastest, then W, then D
    out = torch.empty(1, 240, 480, 480, device=dev)
    f = lambda: process_volume(raw.permute(2, 1, 0), 1.0, 0.0, xy, z, device=dev, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us device, {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms wall per scan")
