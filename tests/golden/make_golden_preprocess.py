"""Generate tests/golden/preprocess.npz from the UNMODIFIED reference `process_file`
(/root/reference/src/utils/preprocess.py:84-151) — run in the build container only:
    python tests/golden/make_golden_preprocess.py
nibabel / matplotlib are absent here, so they are stubbed for the import and `read_nii_data` is patched to hand the
synthetic scan to the reference function; everything numeric is the reference's own code on the CPU."""
import sys
import types
from pathlib import Path

import numpy as np
import pandas as pd

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
for name in ("nibabel", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, "/root/reference/src")
from utils import preprocess as ref            # noqa: E402  (the reference module)
from oracle import preprocess_oracle as PO     # noqa: E402

out = {}
idx = PO.sample_indices()
for case in range(3):
    raw, cfg = PO.synthetic_scan(case)
    ref.read_nii_data = lambda _p, raw=raw: raw.astype(np.float64)          # what nibabel.get_fdata() returns
    meta = pd.DataFrame([{"VolumeName": "scan.nii.gz", "RescaleSlope": cfg["slope"], "RescaleIntercept": cfg["intercept"],
                          "XYSpacing": f"[{cfg['xy']}, {cfg['xy']}]", "ZSpacing": cfg["z"]}])
    vol = ref.process_file("unused", "scan.nii.gz", meta, "ctclip")
    assert tuple(vol.shape) == (1, 240, 480, 480), vol.shape
    flat = vol.reshape(-1).numpy()
    out[f"samples{case}"] = flat[idx].copy()
    out[f"sum{case}"] = np.float64(flat.astype(np.float64).sum())
    out[f"npad{case}"] = np.int64((flat == -1.0).sum())
np.savez_compressed(Path(__file__).resolve().parent / "preprocess.npz", **out)
print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})
