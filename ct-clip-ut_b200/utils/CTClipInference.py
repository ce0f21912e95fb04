"""Drop-in for the reference's `src/utils/CTClipInference.py:35-223` — the caller of the hot path.

The construction contract, `infer()` -> `Visualizations.visualize(...)` and the zero-shot scoring
loop `zeroshot()` (:147-190) are mirrored; dataset I/O, Accelerate and the metric / plotting
helpers are out of scope, so the dataset / dataloader and the process-group description are injected instead of
being built from file paths, and `zeroshot()` stores the gathered predictions and targets as `.npy`.
"""
from __future__ import annotations

import time
from datetime import datetime, timedelta
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.distributed as dist

import numpy as np

from ctclip_b200.attribution import PATHOLOGIES, Visualizations  # noqa: F401
from ctclip_b200.zeroshot import zero_shot


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
        self.accelerator = accelerator or default_accelerator()
        self.model = model.to(self.accelerator.device).eval()
        self.model.accelerator = self.accelerator
        self.zero_shot, self.visualize = zero_shot, visualize
        self.dl, self.tokenizer = dataloader, tokenizer
        self.results_folder = Path(results_folder) / datetime.now().strftime("%d-%m-%Y")
        if self.accelerator.is_main_process:
            self.results_folder.mkdir(parents=True, exist_ok=True)
        self.vis = Visualizations(self.model, self.accelerator, dataset, dataloader, batch_size, self.results_folder,
                                  diff_embeds_folder, tokenizer)

    def zeroshot(self):
        """CTClipInference.zeroshot (:147-190): positive-prompt probabilities [n, 18] (float64) and targets, gathered
        over ranks; rank 0 writes zero_shot_predictions.npy / zero_shot_targets.npy (the reference hands the same two
        arrays to its metric and plot helpers, :193-201).  One image forward per volume instead of 18."""
        pred, targ = zero_shot(self.model, self.dl, self.tokenizer, PATHOLOGIES, self.accelerator.device)
        pred, targ = pred.cpu().numpy(), targ.cpu().numpy()
        if self.accelerator.is_main_process:
            np.save(self.results_folder / "zero_shot_predictions.npy", pred)
            np.save(self.results_folder / "zero_shot_targets.npy", targ)
        return pred, targ

    def infer(self, raw_attention_maps=False, attention_rollout=False, integrated_gradients=False, grad_cam=False,
              occlusion=True):
        """CTClipInference.infer (CTClipInference.py:203-223); the committed reference enables occlusion only."""
        start = time.time()
        if self.zero_shot:
            self.zeroshot()
        if self.visualize:
            self.vis.visualize(raw_attention_maps=raw_attention_maps, attention_rollout=attention_rollout,
                               integrated_gradients=integrated_gradients, grad_cam=grad_cam, occlusion=occlusion)
        if self.accelerator.is_main_process:
            print(f"Evaluation completed. Total Evaluation Time: {timedelta(seconds=time.time() - start)}")
