"""Drop-in for the reference's `src/utils/CTClipInference.py:35-223` — the caller of the hot path.

Only the construction contract and `infer()` -> `Visualizations.visualize(...)` are mirrored;
dataset I/O, Accelerate and the zero-shot evaluation are out of scope, so the dataset / dataloader and the
process-group description are injected instead of being built from file paths.
"""
from __future__ import annotations

import time
from datetime import datetime, timedelta
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.distributed as dist

from ctclip_b200.attribution import Visualizations


def default_accelerator(device=None):
    """Stand-in for accelerate.Accelerator exposing the four attributes the hot path reads
    (visualizations.py:100-103): is_main_process, process_index, num_processes, device."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    device = device or torch.device("cuda", torch.cuda.current_device())
    return SimpleNamespace(is_main_process=rank == 0, process_index=rank, num_processes=world, device=device)


class CTClipInference(torch.nn.Module):
    def __init__(self, model, batch_size=1, dataset=None, dataloader=None, tokenizer=None, results_folder="./results",
                 diff_embeds_folder="./resources", accelerator=None, zero_shot=False, visualize=True, **unused):
        super().__init__()
        if zero_shot:
            raise RuntimeError("zero-shot evaluation (CTClipInference.py:146-201) is outside the attribution hot path")
        self.accelerator = accelerator or default_accelerator()
        self.model = model.to(self.accelerator.device).eval()
        self.model.accelerator = self.accelerator
        self.visualize = visualize
        self.results_folder = Path(results_folder) / datetime.now().strftime("%d-%m-%Y")
        if self.accelerator.is_main_process:
            self.results_folder.mkdir(parents=True, exist_ok=True)
        self.vis = Visualizations(self.model, self.accelerator, dataset, dataloader, batch_size, self.results_folder,
                                  diff_embeds_folder, tokenizer)

    def infer(self, raw_attention_maps=False, attention_rollout=False, integrated_gradients=False, grad_cam=False,
              occlusion=True):
        """CTClipInference.infer (CTClipInference.py:203-223); the committed reference enables occlusion only."""
        start = time.time()
        if self.visualize:
            self.vis.visualize(raw_attention_maps=raw_attention_maps, attention_rollout=attention_rollout,
                               integrated_gradients=integrated_gradients, grad_cam=grad_cam, occlusion=occlusion)
        if self.accelerator.is_main_process:
            print(f"Evaluation completed. Total Evaluation Time: {timedelta(seconds=time.time() - start)}")
