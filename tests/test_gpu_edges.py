"""GPU edge cases of the attribution path at the TINY config: empty / degenerate sweeps, ragged batches, windows
off the token grid, volumes that are all padding.  The reference has no tests of its own; the expected
values follow from its code (visualizations.py:335-424, 851-901) and are checked against the oracle where it applies."""
import numpy as np
import pytest
import torch

from oracle import ctclip_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


@pytest.fixture(scope="module")
def tiny():
    from test_gpu_model import make_engine
    cfg = O.TINY
    eng, sd = make_engine(cfg)
    vol = O.synthetic_volume(cfg, 2).to(DEV)
    tl = eng.text_latents(O.synthetic_text_embeds(cfg, 7).to(DEV))
    return cfg, eng, sd, vol, tl


def test_empty_sweep(tiny):
    """A window larger than the volume yields no windows (the ranges of visualizations.py:340-349 are empty):
    the heat map is all zeros (count 0 -> 1, (h - min)/(max - min + 1e-8) of a zero map), nothing is evaluated."""
    from ctclip_b200 import attribution as A
    cfg, eng, _, vol, tl = tiny
    D, H, W = vol.shape[-3:]
    orig, scores = A.occlusion_scores(eng, vol, tl, [], (2, 4, 4))
    assert scores.numel() == 0 and np.isfinite(orig)
    heat, aux = A.occlusion_sensitivity(eng, vol, tl, (D + 2, H, W), (2, 4, 4))
    assert aux["windows"] == [] and heat.shape == (D, W, H)
    assert float(heat.abs().max()) == 0.0
    ref = O.occlusion_finalize(*O.occlusion_accumulate((D, H, W), [], (D + 2, H, W), orig, []))
    assert np.array_equal(heat.cpu().numpy(), ref)


def test_single_window_covering_the_volume(tiny):
    from ctclip_b200 import attribution as A
    cfg, eng, _, vol, tl = tiny
    D, H, W = vol.shape[-3:]
    heat, aux = A.occlusion_sensitivity(eng, vol, tl, (D, H, W), (D, H, W))
    assert len(aux["windows"]) == 1
    # a constant map normalises to 0 everywhere: (h - min) = 0
    assert float(heat.abs().max()) == 0.0
    # the one score is the logit of the all -1 volume
    allair = torch.full_like(vol, -1.0)
    assert float(aux["scores"][0]) == float(eng.forward(allair, tl).sim[0, 0])


def test_all_padding_volume_skips_every_window(tiny):
    from ctclip_b200 import attribution as A
    cfg, eng, _, vol, tl = tiny
    air = torch.full_like(vol, -1.0)
    stats = {}
    windows = A.occlusion_windows(tuple(air.shape[-3:]), (4, 8, 8), (2, 4, 4))
    orig, scores = A.occlusion_scores(eng, air, tl, windows, (4, 8, 8), stats=stats)
    assert stats == {"evaluated": 0, "noop": len(windows)}
    assert bool((scores == orig).all())
    _, dense = A.occlusion_scores(eng, air, tl, windows[:5], (4, 8, 8), reuse=False)
    assert bool((dense == orig).all())


def test_windows_off_the_token_grid_take_the_dense_path(tiny):
    from ctclip_b200 import attribution as A
    cfg, eng, _, vol, tl = tiny
    ps = (3, 5, 6)
    windows = [(1, 2, 3), (0, 0, 0), (9, 19, 18)]
    with pytest.raises(ValueError, match="aligned"):
        A.occlusion_scores(eng, vol, tl, windows, ps, reuse=True)
    orig, scores = A.occlusion_scores(eng, vol, tl, windows, ps, batch=2)          # ragged last batch (2 + 1)
    for (d, h, w), s in zip(windows, scores.tolist()):
        masked = O.occlusion_mask_apply(vol, (d, h, w), ps)                        # visualizations.py:380-381
        assert float(eng.forward(masked.contiguous(), tl).sim[0, 0]) == s          # fused mask == materialised mask
    assert orig == float(eng.forward(vol, tl).sim[0, 0])


@pytest.mark.parametrize("steps,batch", [(1, 5), (3, 2), (7, 7)])
def test_integrated_gradients_ragged_batches(tiny, steps, batch):
    """steps = 1 is alpha = [0] (torch.linspace(0, 1, 1), visualizations.py:861); batch sizes that do not divide the
    step count must give the same map as one step per batch."""
    from ctclip_b200 import attribution as A
    cfg, eng, _, vol, tl = tiny
    ig, aux = A.integrated_gradients(eng, vol, tl, steps=steps, batch=batch, rot90=False)
    ig1, aux1 = A.integrated_gradients(eng, vol, tl, steps=steps, batch=1, rot90=False)
    assert aux["scores"].shape == (steps,) and torch.equal(aux["scores"], aux1["scores"])
    g, g1 = aux["gsum"], aux1["gsum"]
    assert float((g - g1).abs().max()) <= 1e-5 * float(g1.abs().max())
    assert float((ig - ig1).abs().max()) < 1e-4
    assert torch.isfinite(ig).all() and float(ig.min()) >= 0.0 and float(ig.max()) <= 1.0 + 1e-6
    if steps == 1:                                                                 # alpha = 0: the all-ones baseline
        ones = torch.ones_like(vol)
        assert float(aux["scores"][0]) == float(eng.forward(ones, tl).sim[0, 0])
