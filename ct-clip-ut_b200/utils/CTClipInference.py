"""Drop-in for the reference's `src/utils/CTClipInference.py:35-223` — the caller of the hot path.

The constructor takes the reference's arguments in the reference's order (`CTClipInference.py:39-53`), so the call in
`src/inference_ctclip.py:43-59` works unchanged: it builds the `InferenceDataset` (on the fused GPU `process_file`), the
sampler (`DistributedSampler(shuffle=False, drop_last=True)` across ranks, `RandomSampler` alone, :77-88), the loader and
the `Visualizations` object.  `infer()` -> `Visualizations.visualize(...)` and the zero-shot scoring loop
`zeroshot()` (:147-190, §8f rank 4) are mirrored.  Accelerate is replaced by a four-attribute description of the process
group (one process per GPU, torch.distributed over NCCL); the metric / plot helpers behind the zero-shot loop
(`utils/metrics.py`) are out of scope, so `zeroshot()` stores the gathered predictions and targets as `.npy`.
A ready dataset / loader / accelerator can be injected by keyword instead (tests, tools).
"""
from __future__ import annotations

import os
import time
from datetime import datetime, timedelta
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.distributed as dist
from torch.utils.data import RandomSampler
from torch.utils.data.distributed import DistributedSampler

import numpy as np

from ctclip_b200.attribution import PATHOLOGIES, Visualizations  # noqa: F401
from ctclip_b200.dataset import DeviceLoader, InferenceDataset
from ctclip_b200.zeroshot import zero_shot


def default_accelerator(device=None):
    """Stand-in for accelerate.Accelerator (CTClipInference.py:56-63) exposing the four attributes the hot path reads
    (visualizations.py:100-103): is_main_process, process_index, num_processes, device.  Under torchrun
    (WORLD_SIZE > 1 in the environment) it joins the NCCL process group the way Accelerate would."""
    if not torch.cuda.is_available():
        raise RuntimeError("ctclip_b200: CTClipInference needs an sm_100a CUDA device (there is no CPU path)")
    if dist.is_available() and not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", timeout=timedelta(seconds=36000))
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    device = device or torch.device("cuda", torch.cuda.current_device())
    return SimpleNamespace(is_main_process=rank == 0, process_index=rank, num_processes=world, device=device)


class CTClipInference(torch.nn.Module):
    def __init__(self, model, batch_size=1, data_valid=None, valid_reports=None, valid_labels=None, valid_metadata=None,
                 tokenizer=None, results_folder="./results", diff_embeds_folder="./resources", num_workers=8,
                 num_valid_samples=0, zero_shot=False, visualize=False, *, dataset=None, dataloader=None,
                 accelerator=None):
        super().__init__()
        self.accelerator = accelerator or default_accelerator()
        self.rank, self.world_size = self.accelerator.process_index, self.accelerator.num_processes
        self.maybe_print = print if self.accelerator.is_main_process else (lambda *a, **k: None)
        self.model = model.to(self.accelerator.device).eval()
        self.model.accelerator = self.accelerator
        if tokenizer is None:                                    # CTClipInference.py:70-73
            from transformers import BertTokenizer
            tokenizer = BertTokenizer.from_pretrained("microsoft/BiomedVLP-CXR-BERT-specialized", do_lower_case=True)
        self.tokenizer = tokenizer
        self.batch_size = batch_size
        self.num_valid_samples = num_valid_samples if num_valid_samples else self.world_size      # :75
        if dataset is None and dataloader is None:
            missing = [k for k, v in dict(data_valid=data_valid, valid_reports=valid_reports, valid_labels=valid_labels,
                                          valid_metadata=valid_metadata).items() if v is None]
            if missing:
                raise TypeError(f"CTClipInference: missing {', '.join(missing)} (or inject dataset= / dataloader=)")
            dataset = InferenceDataset(data_folder=data_valid, reports=valid_reports, metadata=valid_metadata,
                                       labels=valid_labels, num_samples=self.num_valid_samples,
                                       device=self.accelerator.device)
        self.ds = dataset
        if dataloader is None:
            if self.world_size > 1:                              # :79-88
                self.sampler = DistributedSampler(self.ds, num_replicas=self.world_size, rank=self.rank,
                                                  shuffle=False, drop_last=True)
            else:
                self.sampler = RandomSampler(self.ds)
            dataloader = DeviceLoader(self.ds, batch_size=batch_size, sampler=self.sampler, num_workers=num_workers)
        self.dl = dataloader
        self.metrics = []
        self.zero_shot, self.visualize = zero_shot, visualize
        self.diff_embeds_folder = diff_embeds_folder
        self.results_folder = Path(results_folder) / datetime.now().strftime("%d-%m-%Y")     # :103-106
        if self.accelerator.is_main_process:
            self.results_folder.mkdir(parents=True, exist_ok=True)
        self.vis = Visualizations(self.model, self.accelerator, self.ds, self.dl, batch_size, self.results_folder,
                                  diff_embeds_folder, tokenizer)
        if self.ds is not None and hasattr(self.ds, "__len__"):
            self.maybe_print(f"Validation size: {len(self.ds)}")

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
        self.maybe_print("Evaluation started")
        if self.zero_shot:
            self.zeroshot()
        if self.visualize:
            self.vis.visualize(raw_attention_maps=raw_attention_maps, attention_rollout=attention_rollout,
                               integrated_gradients=integrated_gradients, grad_cam=grad_cam, occlusion=occlusion)
        self.maybe_print(f"Evaluation completed. Total Evaluation Time: {timedelta(seconds=time.time() - start)}")
