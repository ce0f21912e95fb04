"""The drop-in boundary of SURVEY §8 a21 / §8b: the reference entry script `src/inference_ctclip.py` and the
reference's `CTClipInference(...)` constructor call must work against this package unchanged.

CPU (build container only — needs /root/reference for the script text): the UNMODIFIED script is exec'd with
`monai`, the HuggingFace downloads and the absolute `/mnt/...` paths stubbed; the CUDA device is replaced by a
recording stand-in, so what is checked is the construction contract: every keyword of the reference call is accepted,
an `InferenceDataset` over the two synthetic `.nii.gz` scans, the sampler and the loader are built, and `infer()`
reaches `Visualizations.visualize(occlusion=True)`.
GPU: the same construction through the reference's keywords, then the real thing: `infer()` reads the NIfTI files,
preprocesses them on the device and writes the occlusion heat maps."""
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import preprocess_oracle as PO
from test_preprocess import write_nifti

REF_SCRIPT = Path("/root/reference/src/inference_ctclip.py")
LABEL_COLS = ["Medical material", "Arterial wall calcification", "Cardiomegaly", "Pericardial effusion",
              "Coronary artery wall calcification", "Hiatal hernia", "Lymphadenopathy", "Emphysema", "Atelectasis",
              "Lung nodule", "Lung opacity", "Pulmonary fibrotic sequela", "Pleural effusion", "Mosaic attenuation pattern",
              "Peribronchial thickening", "Consolidation", "Bronchiectasis", "Interlobular septal thickening"]


def make_dataset_tree(root: Path, n: int = 2):
    """Two synthetic scans in the CT-RATE layout the reference walks (InferenceDataset.py:47-60) + the three CSVs."""
    data = root / "data_volumes" / "dataset" / "valid" / "valid_1" / "valid_1_a"
    data.mkdir(parents=True)
    rows_meta, rows_rep, rows_lab = [], [], []
    for i in range(n):
        raw, cfg = PO.synthetic_scan(i)
        name = f"valid_{i + 1}_a_1.nii.gz"
        write_nifti(data / name, np.round(raw).astype(np.int16))
        rows_meta.append({"VolumeName": name, "RescaleSlope": cfg["slope"], "RescaleIntercept": cfg["intercept"],
                          "XYSpacing": f"[{cfg['xy']}, {cfg['xy']}]", "ZSpacing": cfg["z"]})
        rows_rep.append({"VolumeName": name, "Findings_EN": f'Findings "{i}" (none).', "Impressions_EN": " No acute disease."})
        rows_lab.append({"VolumeName": name, **{c: int((i + j) % 5 == 0) for j, c in enumerate(LABEL_COLS)}})
    paths = {"data": root / "data_volumes" / "dataset" / "valid", "meta": root / "valid_metadata.csv",
             "reports": root / "valid_reports.csv", "labels": root / "valid_labels.csv"}
    pd.DataFrame(rows_meta).to_csv(paths["meta"], index=False)
    pd.DataFrame(rows_rep).to_csv(paths["reports"], index=False)
    pd.DataFrame(rows_lab).to_csv(paths["labels"], index=False)
    return paths


class FakeTokens(dict):
    def to(self, device):
        return FakeTokens({k: v.to(device) for k, v in self.items()})


class FakeTokenizer:
    def __len__(self):
        return 64

    def __call__(self, texts, **kw):
        n = len(texts) if isinstance(texts, (list, tuple)) else 1
        return FakeTokens(input_ids=torch.arange(n * 4).view(n, 4))


class FakeBert(torch.nn.Module):
    def __init__(self, dim_text=768):
        super().__init__()
        self.emb = torch.nn.Embedding(64, dim_text)

    def resize_token_embeddings(self, n):
        return None

    def forward(self, input_ids):
        return SimpleNamespace(last_hidden_state=self.emb(input_ids % 64))


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.skipif(not REF_SCRIPT.exists(), reason="needs the reference checkout (build container only)")
def test_unmodified_reference_entry_script_constructs_and_dispatches(tmp_path, monkeypatch):
    paths = make_dataset_tree(tmp_path / "ct_clip_data")
    src = REF_SCRIPT.read_text()
    assert "CTClipInference(" in src and "inference.infer()" in src
    # --- stubs: the two imports the script never uses, the HuggingFace downloads, the checkpoint load
    for name in ("monai", "monai.networks", "monai.networks.nets", "monai.networks.nets.swin_unetr", "monai.utils"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    sys.modules["monai.networks.nets.swin_unetr"].SwinTransformer = object
    sys.modules["monai.utils"].ensure_tuple_rep = lambda *a, **k: None
    import transformers
    monkeypatch.setattr(transformers.BertTokenizer, "from_pretrained", classmethod(lambda cls, *a, **k: FakeTokenizer()))
    monkeypatch.setattr(transformers.BertModel, "from_pretrained", classmethod(lambda cls, *a, **k: FakeBert()))
    import models.ctclip as mc
    import utils.CTClipInference as uci
    loaded = []
    monkeypatch.setattr(mc.CTCLIP, "load", lambda self, path, strict=False: loaded.append(str(path)))
    # --- no GPU in the build container: a CPU description of the process group and a recording visualize()
    monkeypatch.setattr(uci, "default_accelerator", lambda device=None: SimpleNamespace(
        is_main_process=True, process_index=0, num_processes=1, device=torch.device("cpu")))
    calls = []
    monkeypatch.setattr(uci.Visualizations, "visualize", lambda self, **kw: calls.append(kw))
    # --- the absolute paths of the author's machine -> the synthetic tree (the only edit: path literals)
    text = (src.replace("/mnt/ct_clip_data/data_volumes/dataset/valid", str(paths["data"]))
               .replace("/mnt/ct_clip/CT-CLIP-UT/reports/valid_reports.csv", str(paths["reports"]))
               .replace("/mnt/ct_clip/CT-CLIP-UT/labels/valid_labels.csv", str(paths["labels"]))
               .replace("/mnt/ct_clip/CT-CLIP-UT/metadata/valid_metadata.csv", str(paths["meta"]))
               .replace("/mnt/ct_clip/CT-CLIP-UT/src/results/valid/ctclip", str(tmp_path / "results"))
               .replace("/mnt/ct_clip/CT-CLIP-UT/src/resources/pathology_diff_embeddings.npy", str(tmp_path / "emb.npy")))
    assert "/mnt/" not in text.replace("/mnt/ct_clip/pretrained_models/ctclip_v2.pt", "")
    scope = {"__name__": "__main__"}
    exec(compile(text, str(REF_SCRIPT), "exec"), scope)
    inf = scope["inference"]
    assert loaded == ["/mnt/ct_clip/pretrained_models/ctclip_v2.pt"]
    assert type(inf).__name__ == "CTClipInference" and inf.batch_size == 1 and inf.num_valid_samples == 10
    assert len(inf.ds) == 2 and len(inf.dl) == 2 and inf.dl.dataset is inf.ds
    assert type(inf.sampler).__name__ == "RandomSampler"
    assert inf.visualize is True and inf.zero_shot is False
    assert calls == [dict(raw_attention_maps=False, attention_rollout=False, integrated_gradients=False, grad_cam=False,
                          occlusion=True)]
    assert inf.vis.dataset is inf.ds and inf.vis.dist_dataloader is inf.dl
    # the CPU half of a sample (file read + gunzip + metadata): what the DataLoader workers run
    raw = inf.ds.load_raw(0)
    assert raw["raw"].ndim == 3 and raw["labels"].shape == (18,) and raw["name"].startswith("valid_")
    assert '"' not in raw["text"] and "(" not in raw["text"] and raw["text"].endswith("No acute disease.")


def test_reference_constructor_signature_is_accepted_positionally():
    """CTClipInference.py:39-53: (model, batch_size, data_valid, valid_reports, valid_labels, valid_metadata, tokenizer,
    results_folder, diff_embeds_folder, num_workers, num_valid_samples, zero_shot, visualize)."""
    import inspect
    from utils.CTClipInference import CTClipInference
    names = list(inspect.signature(CTClipInference.__init__).parameters)[1:14]
    assert names == ["model", "batch_size", "data_valid", "valid_reports", "valid_labels", "valid_metadata", "tokenizer",
                     "results_folder", "diff_embeds_folder", "num_workers", "num_valid_samples", "zero_shot", "visualize"]
    d = inspect.signature(CTClipInference.__init__).parameters
    assert d["num_workers"].default == 8 and d["num_valid_samples"].default == 0
    assert d["zero_shot"].default is False and d["visualize"].default is False


def test_visualize_ig_does_not_shard_steps_across_ranks_with_different_volumes(monkeypatch, tmp_path):
    """`visualize()` feeds the gradient methods from the DistributedSampler loader — every rank holds
    a DIFFERENT scan — so integrated gradients must not shard its alpha steps over the process group there."""
    from ctclip_b200 import attribution as A
    seen = {}

    def fake_ig(engine, image, tl, steps=50, batch=5, shard_steps=True, rot90=True):
        seen.update(shard_steps=shard_steps, steps=steps)
        return torch.zeros(2, 2, 2), {}
    monkeypatch.setattr(A, "integrated_gradients", fake_ig)
    acc = SimpleNamespace(is_main_process=False, process_index=1, num_processes=2, device=torch.device("cpu"))
    model = SimpleNamespace(engine=lambda device: SimpleNamespace(text_latents=lambda e: e),
                            text_transformer=lambda **kw: SimpleNamespace(last_hidden_state=torch.zeros(1, 1, 4)))
    vis = A.Visualizations(model, acc, None, None, 1, tmp_path, "", None)
    vis.visualize_integrated_gradients(torch.zeros(1, 1, 2, 2, 2), {}, None, "scan", "", steps=7)
    assert seen == {"shard_steps": False, "steps": 7}


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_reference_constructor_end_to_end_on_nifti_files(tmp_path):
    from models.ctclip import CTCLIP
    from utils.CTClipInference import CTClipInference
    from utils.ctvit import CTViT
    from utils.InferenceDataset import InferenceDataset
    paths = make_dataset_tree(tmp_path / "ct_clip_data")
    torch.manual_seed(0)
    vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10, spatial_depth=4,
                temporal_depth=4, dim_head=32, heads=8)
    clip = CTCLIP(text_encoder=FakeBert(), image_encoder=vit, dim_text=768, dim_image=294912, dim_latent=512)
    inference = CTClipInference(                                   # the call of src/inference_ctclip.py:43-59
        clip,
        valid_reports=str(paths["reports"]),
        data_valid=str(paths["data"]),
        valid_labels=str(paths["labels"]),
        valid_metadata=str(paths["meta"]),
        results_folder=str(tmp_path / "results"),
        diff_embeds_folder=str(tmp_path / "emb.npy"),
        tokenizer=FakeTokenizer(),
        batch_size=1,
        num_workers=2,
        num_valid_samples=10,
        zero_shot=False,
        visualize=True)
    assert isinstance(inference.ds, InferenceDataset) and len(inference.ds) == 2
    # the loader: raw halves from two worker processes, device halves here
    batches = list(inference.dl)
    assert len(batches) == 2
    img, texts, labels, names, files = batches[0]
    assert img.is_cuda and tuple(img.shape) == (1, 1, 240, 480, 480) and labels.shape == (1, 18)
    idx = [s[3] for s in inference.ds.samples].index(names[0] + ".nii.gz")
    raw, cfg = PO.synthetic_scan(int(names[0].split("_")[1]) - 1)
    ref = PO.process_volume(np.round(raw).astype(np.int16).astype(np.float32), cfg["slope"], cfg["intercept"], cfg["xy"], cfg["z"])
    assert float((img[0].cpu() - ref).abs().max()) < 2e-6              # == the reference's process_file
    assert torch.equal(inference.ds[idx][0], img[0])
    inference.infer()                                                   # occlusion over both scans
    out = sorted((inference.results_folder / "occlusion").rglob("*_heatmap.npy"))
    assert len(out) == 2, [str(p) for p in inference.results_folder.rglob("*")]
    for f in out:
        a = np.load(f)
        assert a.shape == (240, 480, 480) and a.dtype == np.float32 and np.isfinite(a).all()
        assert 0.0 <= float(a.min()) and float(a.max()) <= 1.0
