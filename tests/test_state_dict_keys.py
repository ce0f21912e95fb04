"""Drop-in boundary: the module mirrors must accept the reference checkpoint — every state-dict key of the
unmodified reference CTCLIP(CTViT) at the benchmark configuration exists in the mirror with the same shape
(fixture: tests/golden/make_golden_keys.py run against /root/reference)."""
import json

import torch


def build():
    from ctclip_b200.modules import CTCLIP, CTViT
    with torch.device("meta"):
        vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
                    spatial_depth=4, temporal_depth=4, dim_head=32, heads=8)
        return CTCLIP(text_encoder=torch.nn.Identity(), image_encoder=vit, dim_text=768, dim_image=294912,
                      dim_latent=512)


def test_reference_checkpoint_keys_and_shapes(golden_dir):
    ref = json.loads((golden_dir / "state_dict_keys.json").read_text())
    assert len(ref) == 156
    ours = {k: list(v.shape) for k, v in build().state_dict().items()}
    assert [k for k in ref if k not in ours] == []
    assert [(k, ref[k], ours[k]) for k in ref if ref[k] != ours[k]] == []
    # the mirror additionally carries the EMA buffer real vector_quantize_pytorch checkpoints hold (unused in eval)
    assert set(ours) - set(ref) <= {"visual_transformer.vq._codebook.embed_avg"}


def test_reference_attribute_surface():
    """Attributes the reference's callers read (ctvit.py:28-33, ctclip.py:45-68, visualizations.py:242-263)."""
    clip = build()
    vit = clip.visual_transformer
    assert (vit.image_size, vit.patch_size, vit.temporal_patch_size) == (480, 20, 10)
    assert (vit.patch_height, vit.patch_width) == (24, 24)
    for name in ("spatial_rel_pos_bias", "to_patch_emb", "to_patch_emb_first_frame", "enc_spatial_transformer",
                 "enc_temporal_transformer", "vq"):
        assert hasattr(vit, name), name
    for name in ("text_transformer", "visual_transformer", "to_text_latent", "to_visual_latent", "temperature"):
        assert hasattr(clip, name), name
    layer = vit.enc_spatial_transformer.layers[0]
    assert len(vit.enc_spatial_transformer.layers) == 4 and len(layer) == 4 and layer[2] is None   # [PEG, Attention, None, FF]
