"""CPU: pins the oracle's restatement of the ATTRIBUTION numerics (oracle.grad_cam_maps / rollout_maps / the raw
attention reductions / occlusion accumulate + finalize) against tests/golden/full_attrib.npz, i.e. against the outputs
of the unmodified reference `Visualizations` class at the benchmark configuration (tests/golden/make_golden.py).
One fp32 forward + backward of the full model on the host (~1 min on 8 cores).  Integrated gradients is not repeated
here (3 more forward/backward passes): its pre-threshold map is compared with the same fixture on the GPU
(tests/test_gpu_attribution.py) and its post-processing chain is `integrated_gradients_post`, checked there too."""
import os

import numpy as np
import pytest
import torch

from oracle import ctclip_oracle as O


def pearson(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    a, b = a - a.mean(), b - b.mean()
    return float((a * b).sum() / np.sqrt((a * a).sum() * (b * b).sum()))


@pytest.fixture(scope="module")
def run(golden_dir):
    torch.set_num_threads(os.cpu_count())
    g = np.load(golden_dir / "full_attrib.npz")
    cfg = O.FULL
    sd = O.init_state_dict(cfg, 42)
    assert abs(float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.dtype.is_floating_point))
               - float(g["sd_checksum"])) < 1e-6 * float(g["sd_checksum"])
    vol = O.synthetic_volume(cfg, 0).requires_grad_()          # a leaf with grad: the captured features join a graph
    cap = {}
    sim = O.ctclip_forward(vol, O.synthetic_text_embeds(cfg, 7), sd, cfg, cap)[0]
    maps = {k: v.detach().numpy() for k, v in O.grad_cam_maps(sim[0, 0], cap).items()}
    sp = [a.detach() for a in cap["spatial_attention_weights"]]
    tp = [a.detach() for a in cap["temporal_attention_weights"]]
    return g, float(sim.detach()), maps, sp, tp


def test_logit_matches_reference_forward(run):
    g, sim, *_ = run
    assert abs(sim - float(g["occ_orig"])) < 2e-6


def test_grad_cam_maps_match_reference(run):
    """visualize_grad_cam (visualizations.py:913-991), incl. the last-layer-features x first-layer-gradients quirk."""
    g, _, maps, _, _ = run
    for k in ("spatial", "spatial_ff", "temporal", "temporal_ff", "combined", "vq"):
        ref = g["gradcam_" + k]
        assert maps[k].shape == ref.shape == (24, 24, 24)
        assert float(np.abs(maps[k] - ref).max()) < 5e-5, k      # maps are normalised to [0, 1]
        assert pearson(maps[k], ref) > 0.999999, k


def test_rollout_and_raw_attention_match_reference(run):
    """visualize_attention_rollout (:779-849) and the reductions of visualize_attention_grid_gif (:659-676)."""
    g, _, _, sp, tp = run
    vol, tvol = O.rollout_maps(sp, tp)
    assert tuple(vol.shape) == (96, 24, 24) and tuple(tvol.shape) == (24, 24, 24)
    assert float(np.abs(vol.numpy() - g["rollout_spatial"]).max()) < 2e-5
    assert float(np.abs(tvol.numpy() - g["rollout_temporal"]).max()) < 2e-5
    heads = sp[0].shape[1]
    rs = np.stack([np.stack([a[:, h].mean(dim=1).numpy() for a in sp]) for h in range(heads)])
    rt = np.stack([np.stack([a[:, h].mean(dim=1).numpy() for a in tp]) for h in range(heads)])
    assert float(np.abs(rs - g["rawattn_spatial"]).max()) < 1e-5 * float(g["rawattn_spatial"].max())
    assert float(np.abs(rt - g["rawattn_temporal"]).max()) < 1e-5 * float(g["rawattn_temporal"].max())
    # the normalised, rotated per-head volumes the grid GIF renders
    raw = O.raw_attention_maps(sp[:1], "spatial")
    rec = torch.from_numpy(g["rawattn_spatial"][0, 0]).reshape(24, 24, 24)
    ref = np.rot90(O.norm_minmax_max(rec).numpy(), k=-1, axes=(0, 1))
    assert float(np.abs(raw[0, 0].numpy() - ref).max()) < 1e-4


def test_occlusion_accumulate_and_finalize_match_reference(golden_dir):
    """_compute_occlusion's float64 accumulation + normalisation + rot90 (:390-424) on the reference's own scores."""
    g = np.load(golden_dir / "full_attrib.npz")
    ps, st = tuple(int(v) for v in g["occ_patch"]), tuple(int(v) for v in g["occ_stride"])
    windows = O.occlusion_windows((240, 480, 480), ps, st)
    assert len(windows) == len(g["occ_scores"]) == 8
    h64, c64 = O.occlusion_accumulate((240, 480, 480), windows, ps, float(g["occ_orig"]), list(g["occ_scores"]))
    heat = O.occlusion_finalize(h64, c64)                       # rot90'd like the reference's return value
    sub = np.rot90(heat, k=1, axes=(1, 2))[5::10, 10::20, 10::20]
    assert np.array_equal(sub.astype(np.float32), g["occ_heat_sub"])
