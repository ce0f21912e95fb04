"""TEST INFRASTRUCTURE — CPU restatement of the reference's CT preprocessing (src/utils/preprocess.py) for the
"ctclip" model type.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.

Pinned: tests/golden/preprocess.npz holds samples of the UNMODIFIED reference `process_file` (imported from
/root/reference with stubbed nibabel / matplotlib and a patched `read_nii_data`) on seeded synthetic scans;
tests/test_oracle_golden.py checks this restatement against them bit for bit."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def resize_array(array: torch.Tensor, current_spacing, target_spacing) -> torch.Tensor:
    """preprocess.py:20-37."""
    shape = array.shape[2:]
    factors = [current_spacing[i] / target_spacing[i] for i in range(3)]
    new_shape = [int(shape[i] * factors[i]) for i in range(3)]
    return F.interpolate(array, size=new_shape, mode="trilinear", align_corners=False)


def crop_and_pad(array: torch.Tensor, target_shape, pad_value=-1) -> torch.Tensor:
    """preprocess.py:39-82 on an [H, W, D] tensor: centre crop / symmetric pad (pad_before = total // 2)."""
    out = array
    for i in range(3):
        n, t = array.shape[i], target_shape[i]
        if n > t:
            start = (n - t) // 2
            out = out.narrow(i, start, t)
        elif n < t:
            before = (t - n) // 2
            pad = [0, 0, 0, 0, 0, 0]
            pad[2 * (2 - i)] = before
            pad[2 * (2 - i) + 1] = t - n - before
            out = F.pad(out, pad, mode="constant", value=pad_value)
    return out


def process_volume(raw_hwd: np.ndarray, slope: float, intercept: float, xy_spacing: float, z_spacing: float,
                   target_spacing=(1.5, 0.75, 0.75), target_shape_hwd=(480, 480, 240)) -> torch.Tensor:
    """preprocess.py:118-151 (model_type == "ctclip") after the NIfTI read: returns [1, D, H, W] fp32."""
    img = torch.from_numpy(np.ascontiguousarray(raw_hwd)).float()
    img = slope * img + intercept
    img = img.permute(2, 0, 1).unsqueeze(0).unsqueeze(0)
    img = resize_array(img, (z_spacing, xy_spacing, xy_spacing), target_spacing)
    img = torch.clamp(img, -1000, 1000) / 1000.0
    img = img[0, 0].permute(1, 2, 0)
    img = crop_and_pad(img, target_shape_hwd, pad_value=-1)
    return img.permute(2, 0, 1).unsqueeze(0).unsqueeze(0).squeeze(0)


def synthetic_scan(case: int):
    """Seeded raw scans + metadata exercising pad-only, crop + exact-fit rounding, and mixed crop/pad."""
    cfg = [dict(shape=(160, 160, 60), slope=1.0, intercept=-1024.0, xy=1.2, z=3.0),
           dict(shape=(200, 180, 100), slope=1.0, intercept=-1024.0, xy=2.0, z=4.0),
           dict(shape=(96, 256, 33), slope=0.5, intercept=-200.0, xy=0.9, z=5.0)][case]
    g = np.random.default_rng(100 + case)
    H, W, D = cfg["shape"]
    raw = (g.normal(1000.0, 600.0, size=(H, W, D))).astype(np.float32)
    raw[: H // 8] = 0.0                                     # air
    return raw, cfg


def sample_indices(n: int = 8192, numel: int = 240 * 480 * 480, seed: int = 5):
    return np.random.default_rng(seed).integers(0, numel, n)
