"""BASELINE.json configs[4]: the full attribution suite (raw attention, rollout, Grad-CAM, IG-50, occlusion sweep)
over N synthetic 480x480x240 volumes through the drop-in entry point `CTClipInference.infer()` ->
`Visualizations.visualize(...)`, `.npy` outputs included.  Volumes are sharded over ranks for the four single-volume
methods (the reference's DistributedSampler); the occlusion sweep shards windows and broadcasts each sample
(visualizations.py:1141-1178).

    python tools/suite.py --volumes 2 [--out /tmp/ctclip_suite]
    torchrun --nproc-per-node 8 tools/suite.py --volumes 64
"""
import argparse
import json
import os
import shutil
import sys
import time
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
from models.ctclip import CTCLIP  # noqa: E402
from utils.ctvit import CTViT  # noqa: E402
from utils.CTClipInference import CTClipInference  # noqa: E402


class Tokens(dict):
    def to(self, device):
        return Tokens({k: v.to(device) for k, v in self.items()})


class Tokenizer:
    def __call__(self, texts, **kw):
        n = len(texts) if isinstance(texts, (list, tuple)) else 1
        return Tokens(input_ids=torch.arange(n * 4).view(n, 4))


class TextTower(torch.nn.Module):
    """Stand-in for CXR-BERT (weights are not available offline): [CLS] = an embedding row."""

    def __init__(self, dim_text=768):
        super().__init__()
        self.emb = torch.nn.Embedding(64, dim_text)

    def forward(self, input_ids):
        return SimpleNamespace(last_hidden_state=self.emb(input_ids % 64))


def volume(i):
    g = torch.Generator().manual_seed(1234 + i)
    v = (0.35 * torch.randn(1, 240, 480, 480, generator=g) - 0.2).clamp_(-1, 1)
    v[:, :16] = -1; v[:, -16:] = -1
    v[:, :, :40] = -1; v[:, :, -40:] = -1
    v[..., :40] = -1; v[..., -40:] = -1
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--volumes", type=int, default=2)
    ap.add_argument("--out", default="/tmp/ctclip_suite")
    ap.add_argument("--keep", action="store_true", help="keep the .npy outputs (2.4 GB per volume) instead of "
                    "unlinking each file right after it has been written")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
                spatial_depth=4, temporal_depth=4, dim_head=32, heads=8)
    clip = CTCLIP(text_encoder=TextTower(), image_encoder=vit, dim_text=768, dim_image=294912, dim_latent=512)
    labels = torch.zeros(18)
    # Only rank 0 ever indexes the dataset (the occlusion branch loads on rank 0 and broadcasts, visualizations.py:296-318);
    # the other ranks hold just their own shard of volumes, so that 64 volumes x 221 MB are not replicated in every process.
    sample = lambda i: (volume(i), "no acute findings", labels, f"scan{i}", f"scan{i}.nii.gz")
    mine = {i: sample(i) for i in range(args.volumes) if rank == 0 or i % world == rank}

    class Dataset:
        def __len__(self):
            return args.volumes

        def __getitem__(self, i):
            return mine[i]
    dataset = Dataset()
    loader = [(mine[i][0].unsqueeze(0).pin_memory(), [mine[i][1]], labels.unsqueeze(0), [mine[i][3]], [mine[i][4]])
              for i in range(args.volumes) if i % world == rank]
    inf = CTClipInference(clip, batch_size=1, dataset=dataset, dataloader=loader, tokenizer=Tokenizer(),
                          results_folder=args.out)
    written = {"files": 0, "bytes": 0}
    if not args.keep:
        # 64 volumes x 11 maps x 221 MB = 155 GB: every file is written in full (the write is part of the measured
        # suite) and unlinked at once, so the disk never holds more than one map
        import numpy as np
        real_save = np.save

        def save_and_unlink(path, arr, *a, **k):
            real_save(path, arr, *a, **k)
            f = Path(str(path) if str(path).endswith(".npy") else str(path) + ".npy")
            written["files"] += 1
            written["bytes"] += f.stat().st_size
            f.unlink()
        np.save = save_and_unlink
    times = {}
    for method in ("raw_attention_maps", "attention_rollout", "grad_cam", "integrated_gradients", "occlusion"):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        inf.vis.visualize(**{method: True})
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times[method] = time.perf_counter() - t0
    total = sum(times.values())
    if rank == 0:
        files = list(Path(inf.results_folder).rglob("*.npy"))
        nbytes = sum(f.stat().st_size for f in files) + written["bytes"]
        files = files + [None] * written["files"]
        print(json.dumps({"suite": "raw attention + rollout + Grad-CAM + IG-50 + occlusion(12167 windows), .npy outputs on disk",
                          "volumes": args.volumes, "n_gpus": world, "seconds": {k: round(v, 3) for k, v in times.items()},
                          "total_s": round(total, 3), "volumes_per_s": args.volumes / total,
                          "npy_files": len(files), "npy_gb": round(nbytes / 1e9, 2)}))
        if not args.keep:
            shutil.rmtree(args.out, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
