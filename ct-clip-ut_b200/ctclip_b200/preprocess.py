"""CT scan preprocessing in front of the patch embedding — the reference's src/utils/preprocess.py
(`read_nii_data`, `process_file`) with the numeric part (HU rescale, trilinear resample to (1.5, 0.75, 0.75) mm,
clamp, / 1000, centre crop / symmetric -1 pad to 480 x 480 x 240) fused into ONE sm_100a kernel
(`ctc_preprocess_ct`) that reads the raw voxels in NIfTI file order and writes the fp32 volume the model consumes.
The reference does this with five full-size intermediate tensors on the CPU inside DataLoader workers.

Only the "ctclip" model type is on this path; "ctgenerate" is out of scope."""
from __future__ import annotations

import ctypes
import gzip
import struct
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import call, stream_ptr

TARGET_SPACING = (1.5, 0.75, 0.75)          # (z, xy, xy)  preprocess.py:128
TARGET_SHAPE_HWD = (480, 480, 240)          # preprocess.py:140
_NIFTI_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4"}
_KERNEL_DTYPES = {torch.float32: 0, torch.int16: 1, torch.float64: 2}


def read_nii_raw(file_path) -> Tuple[np.ndarray, float, float]:
    """Minimal NIfTI-1 reader (.nii / .nii.gz, single-file): returns (stored voxels as an (i, j, k)-shaped array in
    FILE order, i.e. Fortran-contiguous, scl_slope, scl_inter).  nibabel's `get_fdata()` (preprocess.py:8-18) is
    `stored * scl_slope + scl_inter` as float64 when scl_slope is non-zero and finite, else the stored values."""
    p = Path(file_path)
    opener = gzip.open if p.suffix == ".gz" else open
    with opener(p, "rb") as f:
        buf = bytearray(f.read())                     # writable, so torch.from_numpy can wrap it without a copy
    if len(buf) < 352:
        raise ValueError(f"{p}: not a NIfTI-1 file (too short)")
    end = "<" if struct.unpack("<i", buf[:4])[0] == 348 else ">"
    if struct.unpack(end + "i", buf[:4])[0] != 348:
        raise ValueError(f"{p}: not a NIfTI-1 file (sizeof_hdr != 348)")
    dim = struct.unpack(end + "8h", buf[40:56])
    datatype, = struct.unpack(end + "h", buf[70:72])
    vox_offset, scl_slope, scl_inter = struct.unpack(end + "3f", buf[108:120])
    if dim[0] < 3 or any(d != 1 for d in dim[4:1 + dim[0]]):
        raise ValueError(f"{p}: expected a 3-D volume, header dim = {dim}")
    if datatype not in _NIFTI_DTYPES:
        raise ValueError(f"{p}: unsupported NIfTI datatype code {datatype}")
    shape = tuple(int(d) for d in dim[1:4])
    dt = np.dtype(end + _NIFTI_DTYPES[datatype])
    off = int(vox_offset) if vox_offset >= 352 else 352
    n = shape[0] * shape[1] * shape[2]
    data = np.frombuffer(buf, dtype=dt, count=n, offset=off).reshape(shape, order="F")
    if dt.byteorder == ">" or dt.name not in ("int16", "float32", "float64"):
        # the kernel reads native int16 / float32 / float64; everything else is widened once on the host
        data = data.astype(np.float32 if dt.name not in ("int16", "float64") else dt.newbyteorder("="), order="F")
    if not np.isfinite(scl_slope) or scl_slope == 0:
        scl_slope, scl_inter = 1.0, 0.0
    if not np.isfinite(scl_inter):
        scl_inter = 0.0
    return data, float(scl_slope), float(scl_inter)


def read_nii_data(file_path) -> Optional[np.ndarray]:
    """Same contract as preprocess.py:8-18: float64 voxel array or None (with a message) on failure."""
    try:
        data, s, i = read_nii_raw(file_path)
        return data.astype(np.float64) * s + i if (s != 1.0 or i != 0.0) else data.astype(np.float64)
    except Exception as e:          # the reference prints and returns None
        print(f"Error reading file {file_path}: {e}")
        return None


def process_volume(raw, slope: float, intercept: float, xy_spacing: float, z_spacing: float,
                   device: Optional[torch.device] = None, target_spacing=TARGET_SPACING,
                   target_shape_hwd=TARGET_SHAPE_HWD, pad_value: float = -1.0,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`process_file` (preprocess.py:84-151, model_type "ctclip") after the file read: raw voxels of logical shape
    [H, W, D] (numpy array or tensor, host or device; any strides, float32 / float64 / int16) -> fp32 tensor [1, D, H, W]
    on `device`.  One H2D copy of the raw voxels (none for a device tensor), one kernel.  `out`: write into this
    contiguous fp32 [1, D, H, W] device tensor (e.g. one row of a batch buffer) instead of allocating the result."""
    dev = device or torch.device("cuda", torch.cuda.current_device())
    t = torch.from_numpy(raw) if isinstance(raw, np.ndarray) else raw
    if t.dtype not in _KERNEL_DTYPES:
        t = t.to(torch.float32)                       # the reference casts to fp32 before the HU transform anyway
    if t.dim() != 3:
        raise ValueError(f"expected a 3-D [H, W, D] scan, got shape {tuple(t.shape)}")
    # keep the memory order of the file (first axis fastest): move the bytes, not a transposed copy
    perm = sorted(range(3), key=lambda a: -t.stride(a))
    td = t.permute(*perm).contiguous().to(dev, non_blocking=True)
    inv = [perm.index(a) for a in range(3)]
    td = td.permute(*inv)                             # logical [H, W, D] view of the device copy
    H0, W0, D0 = td.shape
    Ht, Wt, Dt = target_shape_hwd
    if out is None:
        out = torch.empty(1, Dt, Ht, Wt, dtype=torch.float32, device=dev)
    elif (tuple(out.shape) != (1, Dt, Ht, Wt) or out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous()):
        raise ValueError(f"out must be a contiguous fp32 CUDA tensor of shape {(1, Dt, Ht, Wt)}, got {tuple(out.shape)} {out.dtype}")
    res = (ctypes.c_int * 3)()
    call("ctc_preprocess_ct", td, _KERNEL_DTYPES[td.dtype], H0, W0, D0, td.stride(0), td.stride(1), td.stride(2),
         float(np.float32(slope)), float(np.float32(intercept)), float(z_spacing), float(xy_spacing),
         float(target_spacing[0]), float(target_spacing[1]), Dt, Ht, Wt, float(pad_value), out, res, stream_ptr())
    out.resampled_shape = (res[0], res[1], res[2])
    return out


def process_file(file_path, file_name, metadata_df, model_type, device: Optional[torch.device] = None):
    """Drop-in for preprocess.py:84-151 (same arguments, same prints / None on failure, same [1, D, H, W] result)."""
    if model_type != "ctclip":
        raise NotImplementedError("ctclip_b200 preprocesses for the CT-CLIP path only (model_type='ctclip')")
    try:
        raw, s, i = read_nii_raw(file_path)
    except Exception as e:
        print(f"Error reading file {file_path}: {e}")
        print(f"Read failure for {file_path}.")
        return None
    row = metadata_df[metadata_df["VolumeName"] == file_name]
    if row.empty:
        print(f"No metadata found for {file_name}.")
        return None
    try:
        slope = float(row["RescaleSlope"].iloc[0])
        intercept = float(row["RescaleIntercept"].iloc[0])
        xy_spacing = float(row["XYSpacing"].iloc[0][1:][:-2].split(",")[0])
        z_spacing = float(row["ZSpacing"].iloc[0])
    except Exception as e:
        print(f"Error processing metadata for {file_name}: {e}")
        return None
    if s != 1.0 or i != 0.0:                          # get_fdata() scaling happens in float64 before the fp32 cast
        raw = (raw.astype(np.float64) * s + i).astype(np.float32)
    return process_volume(raw, slope, intercept, xy_spacing, z_spacing, device)
