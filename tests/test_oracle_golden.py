"""CPU: the oracle restatement (oracle/ctclip_oracle.py) against fixtures produced by the
UNMODIFIED reference classes (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import ctclip_oracle as O


def _sd_checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.dtype.is_floating_point))


@pytest.fixture(scope="module")
def tiny(golden_dir):
    g = np.load(golden_dir / "tiny_model.npz")
    sd = O.init_state_dict(O.TINY, 42)
    assert abs(_sd_checksum(sd) - float(g["sd_checksum"])) < 1e-6 * float(g["sd_checksum"])
    return g, sd


def test_tiny_cpb_bias(tiny):
    g, sd = tiny
    b = O.cpb_bias(sd, "visual_transformer.spatial_rel_pos_bias.", O.TINY.h, O.TINY.w)
    np.testing.assert_allclose(b.numpy(), g["cpb_bias"], rtol=1e-5, atol=1e-6)


def test_tiny_patch_embed_and_encode(tiny):
    g, sd = tiny
    img = O.synthetic_volume(O.TINY, 0, batch=2)
    pe = O.patch_embed(img, sd, O.TINY)
    np.testing.assert_allclose(pe.numpy(), g["patch_emb"], rtol=1e-4, atol=2e-5)
    enc = O.encode(pe, sd, O.TINY)
    np.testing.assert_allclose(enc.numpy(), g["encoded"], rtol=1e-4, atol=5e-5)


def test_tiny_forward_and_input_grad(tiny):
    g, sd = tiny
    img = O.synthetic_volume(O.TINY, 0, batch=2).requires_grad_()
    txt = O.synthetic_text_embeds(O.TINY, 7, batch=2)
    sim, il, tl, temp, tokens, ind = O.ctclip_forward(img, txt, sd, O.TINY)
    assert (ind.numpy() == g["indices"].reshape(ind.shape)).all()
    np.testing.assert_allclose(tokens.detach().numpy(), g["tokens"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sim.detach().numpy(), g["sim"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(il.detach().numpy(), g["image_latents"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(tl.detach().numpy(), g["text_latents"], rtol=1e-4, atol=1e-6)
    (gi,) = torch.autograd.grad(sim[0, 0] + 0.5 * sim[1, 1], img)
    ref = g["grad_image"]
    err = np.abs(gi.numpy() - ref).max() / np.abs(ref).max()
    assert err < 1e-3, err


def test_window_enumeration_matches_reference_count():
    # visualization.ipynb:362 : 12 167 windows at the defaults
    w = O.occlusion_windows((240, 480, 480))
    assert len(w) == 23 ** 3 == 12167
    assert w[0] == (0, 0, 0) and w[1] == (0, 0, 20) and w[23] == (0, 20, 0)
    # remainder drop 1/3/7 at world 2/4/8 (SURVEY a18 [probe])
    for world, dropped in ((2, 1), (4, 3), (8, 7)):
        tot = sum(len(O.shard_windows(w, r, world)) for r in range(world))
        assert len(w) - tot == dropped
