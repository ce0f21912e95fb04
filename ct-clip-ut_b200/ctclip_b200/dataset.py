"""The dataset in front of the hot path — the reference's src/utils/InferenceDataset.py:8-76 behind the same
constructor and sample tuple `(image [1, D, H, W], report text, labels fp32 [18], scan name, path)`.

What differs is where the work runs.  The reference does the whole of `process_file` (gunzip, HU rescale, trilinear
resample, clamp, crop / pad; preprocess.py:84-151) on the CPU inside DataLoader workers.  Here a sample has two halves:
`load_raw` (file read + gunzip + metadata lookup: CPU only, safe in forked workers) and `finish` (one H2D copy of the raw
voxels + the fused `ctc_preprocess_ct` kernel, in the process that owns the GPU).  `DeviceLoader` runs the first half in
`num_workers` DataLoader workers and the second in the consumer, and yields the batches a default-collate DataLoader over
the reference dataset would yield.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader, Dataset

from .preprocess import process_volume, read_nii_raw


def _clean_report(text: str) -> str:
    """InferenceDataset.py:70-74: quotes and parentheses removed, stripped."""
    for ch in ('"', "'", "(", ")"):
        text = text.replace(ch, "")
    return text.strip()


class InferenceDataset(Dataset):
    def __init__(self, data_folder, reports, metadata, labels, num_samples=500, model_type="ctclip", device=None):
        if model_type != "ctclip":
            raise NotImplementedError("ctclip_b200 serves the CT-CLIP path only (model_type='ctclip')")
        self.data_folder = data_folder
        self.metadata_df = pd.read_csv(metadata)
        self.labels = labels
        self.model_type = model_type
        self.device = device
        self.observations = self._load_observations(reports)
        self.samples = self._prepare_samples()
        if num_samples < len(self.samples):                     # InferenceDataset.py:28-29
            self.samples = self.samples[:num_samples]

    def _load_observations(self, reports):
        """VolumeName -> (findings, impressions) (InferenceDataset.py:31-37; NaN cells become the string 'nan')."""
        df = pd.read_csv(reports)
        return {name: (str(f) or "", str(i) or "")
                for name, f, i in zip(df["VolumeName"], df["Findings_EN"], df["Impressions_EN"])}

    def _prepare_samples(self):
        """os.walk over the data folder in the reference's order; a scan needs a report row and a label row
        (InferenceDataset.py:39-62).  Labels = the columns after the first of the label CSV."""
        labels_df = pd.read_csv(self.labels)
        onehot = {name: row for name, row in zip(labels_df["VolumeName"], labels_df[list(labels_df.columns[1:])].values)}
        samples = []
        for root, _, files in os.walk(self.data_folder):
            for file in files:
                if not file.endswith(".nii.gz") or file not in self.observations or file not in onehot:
                    continue
                findings, impressions = self.observations[file]
                samples.append((os.path.join(root, file), findings + impressions, onehot[file], file))
        return samples

    def __len__(self):
        return len(self.samples)

    # -- first half: CPU only -----------------------------------------------------------------------------------
    def load_raw(self, index) -> dict:
        path, observations, labels, name = self.samples[index]
        raw, s, i = read_nii_raw(path)
        row = self.metadata_df[self.metadata_df["VolumeName"] == name]
        if row.empty:
            raise KeyError(f"No metadata found for {name}.")
        if s != 1.0 or i != 0.0:                       # nibabel's get_fdata() scaling, in float64 (preprocess.py:8-18)
            raw = (raw.astype(np.float64) * s + i).astype(np.float32)
        return {"raw": raw,
                "slope": float(row["RescaleSlope"].iloc[0]), "intercept": float(row["RescaleIntercept"].iloc[0]),
                "xy": float(row["XYSpacing"].iloc[0][1:][:-2].split(",")[0]), "z": float(row["ZSpacing"].iloc[0]),
                "text": _clean_report(observations), "labels": np.asarray(labels, dtype=np.float32),
                "name": name.replace(".nii.gz", ""), "path": path}

    # -- second half: on the GPU --------------------------------------------------------------------------------
    def finish(self, r: dict):
        image = process_volume(r["raw"], r["slope"], r["intercept"], r["xy"], r["z"], self.device)
        return image, r["text"], torch.from_numpy(r["labels"]), r["name"], r["path"]

    def __getitem__(self, index):
        return self.finish(self.load_raw(index))


class _RawView(Dataset):
    def __init__(self, ds: InferenceDataset):
        self.ds = ds

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, index):
        return self.ds.load_raw(index)


class DeviceLoader:
    """DataLoader(ds, batch_size, sampler, num_workers) of CTClipInference.py:90 with the split described in the module
    docstring.  Iterating yields (images [B,1,D,H,W] on the device, [texts], labels [B,18], [names], [paths])."""

    def __init__(self, dataset: InferenceDataset, batch_size: int = 1, sampler=None, num_workers: int = 0):
        self.dataset, self.batch_size, self.sampler = dataset, batch_size, sampler
        self._dl = DataLoader(_RawView(dataset), batch_size=batch_size, sampler=sampler, num_workers=num_workers,
                              collate_fn=list)

    def __len__(self):
        return len(self._dl)

    def __iter__(self):
        for raws in self._dl:
            done = [self.dataset.finish(r) for r in raws]
            images = torch.stack([d[0] for d in done])
            labels = torch.stack([d[2] for d in done])
            yield images, [d[1] for d in done], labels, [d[3] for d in done], [d[4] for d in done]
