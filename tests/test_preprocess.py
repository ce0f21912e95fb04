"""CT preprocessing (reference src/utils/preprocess.py).
CPU: the oracle restatement against golden samples of the UNMODIFIED reference `process_file`
(tests/golden/make_golden_preprocess.py), and the minimal NIfTI reader.
GPU: the fused kernel (`ctc_preprocess_ct`) against the oracle on the same scans."""
import gzip
import struct

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as PO


def write_nifti(path, data, scl_slope=0.0, scl_inter=0.0, endian="<"):
    """Write a single-file NIfTI-1 (.nii or .nii.gz) with `data` (i, j, k) stored in file (Fortran) order."""
    code = {np.dtype("int16"): 4, np.dtype("float32"): 16, np.dtype("uint8"): 2, np.dtype("float64"): 64}[data.dtype]
    hdr = bytearray(352)
    struct.pack_into(endian + "i", hdr, 0, 348)
    struct.pack_into(endian + "8h", hdr, 40, 3, data.shape[0], data.shape[1], data.shape[2], 1, 1, 1, 1)
    struct.pack_into(endian + "h", hdr, 70, code)
    struct.pack_into(endian + "h", hdr, 72, data.dtype.itemsize * 8)
    struct.pack_into(endian + "3f", hdr, 108, 352.0, scl_slope, scl_inter)
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + data.astype(data.dtype.newbyteorder(endian)).tobytes(order="F")
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(payload)


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("case", [0, 1, 2])
def test_oracle_matches_reference_process_file(golden_dir, case):
    gold = np.load(golden_dir / "preprocess.npz")
    raw, cfg = PO.synthetic_scan(case)
    vol = PO.process_volume(raw, cfg["slope"], cfg["intercept"], cfg["xy"], cfg["z"])
    assert tuple(vol.shape) == (1, 240, 480, 480)
    flat = vol.reshape(-1).numpy()
    np.testing.assert_array_equal(flat[PO.sample_indices()], gold[f"samples{case}"])        # bit-exact
    assert int((flat == -1.0).sum()) == int(gold[f"npad{case}"])
    assert abs(float(flat.astype(np.float64).sum()) - float(gold[f"sum{case}"])) < 1e-6 * abs(float(gold[f"sum{case}"]))


@pytest.mark.parametrize("suffix,dtype,endian,scl", [(".nii", "int16", "<", (0.0, 0.0)), (".nii.gz", "int16", "<", (2.0, -1024.0)),
                                                     (".nii.gz", "float32", ">", (0.0, 0.0)), (".nii", "uint8", "<", (1.0, 0.0))])
def test_nifti_reader(tmp_path, suffix, dtype, endian, scl):
    from ctclip_b200.preprocess import read_nii_data, read_nii_raw
    g = np.random.default_rng(3)
    data = g.integers(0, 200, size=(7, 5, 4)).astype(dtype)
    p = tmp_path / ("scan" + suffix)
    write_nifti(p, data, scl[0], scl[1], endian)
    raw, s, i = read_nii_raw(p)
    assert raw.shape == (7, 5, 4) and raw.flags["F_CONTIGUOUS"]
    np.testing.assert_array_equal(raw.astype(np.float64), data.astype(np.float64))
    expect = data.astype(np.float64) * (scl[0] if scl[0] != 0 else 1.0) + (scl[1] if scl[0] != 0 else 0.0)
    np.testing.assert_array_equal(read_nii_data(p), expect)
    assert read_nii_data(tmp_path / "missing.nii") is None               # reference prints and returns None


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1, 2])
@pytest.mark.parametrize("layout", ["file_order", "c_order_int16"])
def test_kernel_matches_oracle(case, layout):
    from ctclip_b200.preprocess import process_volume
    raw, cfg = PO.synthetic_scan(case)
    if layout == "c_order_int16":
        raw = np.round(raw).astype(np.int16)                              # what a CT NIfTI stores
        arr = np.ascontiguousarray(raw)
    else:
        arr = np.asfortranarray(raw)                                      # NIfTI file order: first axis fastest
    ref = PO.process_volume(raw.astype(np.float32), cfg["slope"], cfg["intercept"], cfg["xy"], cfg["z"])
    out = process_volume(arr, cfg["slope"], cfg["intercept"], cfg["xy"], cfg["z"])
    torch.cuda.synchronize()
    assert tuple(out.shape) == (1, 240, 480, 480) and out.dtype == torch.float32
    o = out.cpu()
    assert torch.equal(o == -1.0, ref == -1.0) or int(((o == -1.0) != (ref == -1.0)).sum()) < 50   # same crop / pad box
    err = float((o - ref).abs().max())
    print(f"[preprocess case {case} {layout}] resampled {out.resampled_shape} max abs err {err:.2e}")
    assert err < 2e-6                                                     # fp32 evaluation-order differences only


@pytest.mark.gpu
def test_process_file_end_to_end(tmp_path):
    import pandas as pd
    from ctclip_b200.preprocess import process_file
    raw, cfg = PO.synthetic_scan(2)
    stored = np.round(raw).astype(np.int16)
    write_nifti(tmp_path / "scan.nii.gz", stored)
    meta = pd.DataFrame([{"VolumeName": "scan.nii.gz", "RescaleSlope": cfg["slope"], "RescaleIntercept": cfg["intercept"],
                          "XYSpacing": f"[{cfg['xy']}, {cfg['xy']}]", "ZSpacing": cfg["z"]}])
    out = process_file(tmp_path / "scan.nii.gz", "scan.nii.gz", meta, "ctclip")
    ref = PO.process_volume(stored.astype(np.float32), cfg["slope"], cfg["intercept"], cfg["xy"], cfg["z"])
    assert float((out.cpu() - ref).abs().max()) < 2e-6
    assert process_file(tmp_path / "scan.nii.gz", "other.nii.gz", meta, "ctclip") is None   # no metadata row
    assert process_file(tmp_path / "nope.nii.gz", "scan.nii.gz", meta, "ctclip") is None    # unreadable file
