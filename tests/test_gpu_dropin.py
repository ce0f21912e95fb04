"""GPU: the drop-in module surface end to end — `utils.CTClipInference.CTClipInference` +
`utils.visualizations.Visualizations` imported the way `src/inference_ctclip.py:5-7` imports them, driven with a
stand-in text tower / tokenizer / dataset, must run all five attribution methods and write the reference's `.npy`
names (visualizations.py:638-639, 848-849, 906, 1021-1026, 1082)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ctclip_oracle as O

pytestmark = pytest.mark.gpu


class FakeTokens(dict):
    def to(self, device):
        return FakeTokens({k: v.to(device) for k, v in self.items()})


class FakeTokenizer:
    def __call__(self, texts, **kw):
        n = len(texts) if isinstance(texts, (list, tuple)) else 1
        return FakeTokens(input_ids=torch.arange(n * 4).view(n, 4))


class FakeTextTower(torch.nn.Module):
    def __init__(self, dim_text):
        super().__init__()
        self.emb = torch.nn.Embedding(64, dim_text)

    def forward(self, input_ids):
        return SimpleNamespace(last_hidden_state=self.emb(input_ids % 64))


def test_inference_entry_point_writes_reference_outputs(tmp_path):
    from models.ctclip import CTCLIP                       # the import names of src/inference_ctclip.py:5-7
    from utils.ctvit import CTViT
    from utils.CTClipInference import CTClipInference
    cfg = O.TINY
    torch.manual_seed(0)
    vit = CTViT(dim=cfg.dim, codebook_size=cfg.codebook_size, image_size=cfg.image_size, patch_size=cfg.patch_size,
                temporal_patch_size=cfg.temporal_patch_size, spatial_depth=cfg.spatial_depth,
                temporal_depth=cfg.temporal_depth, dim_head=cfg.dim_head, heads=cfg.heads)
    T, H = cfg.depth_voxels // cfg.temporal_patch_size, cfg.image_size // cfg.patch_size
    clip = CTCLIP(text_encoder=FakeTextTower(cfg.dim_text), image_encoder=vit, dim_text=cfg.dim_text,
                  dim_image=H * H * cfg.dim, dim_latent=cfg.dim_latent)
    vol = O.synthetic_volume(cfg, 0)[0]                    # [1, D, H, W] like the dataset's samples
    labels = torch.zeros(18)
    sample = (vol, "no acute findings", labels, "scan0", "scan0.nii.gz")
    dataset = [sample]
    loader = [(vol.unsqueeze(0), ["no acute findings"], labels.unsqueeze(0), ["scan0"], ["scan0.nii.gz"])]
    inf = CTClipInference(clip, batch_size=1, dataset=dataset, dataloader=loader, tokenizer=FakeTokenizer(),
                          results_folder=tmp_path)
    vis = inf.vis
    vis.visualize(raw_attention_maps=True, attention_rollout=True, integrated_gradients=True, grad_cam=True)
    D, Hh, Ww = vol.shape[-3:]
    tokens = FakeTokenizer()(["no acute findings"]).to("cuda")
    vis.visualize_occlusion_sensitivity(vol.unsqueeze(0).cuda(), tokens, labels, "scan0", "scan0.nii.gz",
                                        patch_size=(4, 8, 8), stride=(2, 4, 4))
    root = inf.results_folder

    def load(rel):
        f = list(root.glob(rel))
        assert len(f) == 1, (rel, [str(p) for p in root.rglob("*.npy")])
        return np.load(f[0], allow_pickle=True)

    assert load("raw_attention_grids/1/scan0_spatial_grid.npy").ndim >= 3
    assert load("raw_attention_grids/1/scan0_temporal_grid.npy").ndim >= 3
    for name in ("attention_rollout/1/scan0_spatial.npy", "attention_rollout/1/scan0_temporal.npy",
                 "integrated_gradients/1/scan0.npy", "occlusion/1/scan0__heatmap.npy",
                 *[f"grad_cam/1/scan0_{k}.npy" for k in ("spatial_ff", "temporal_ff", "spatial", "temporal", "combined", "vq")]):
        a = load(name)
        assert a.shape == (D, Ww, Hh) and a.dtype == np.float32 and np.isfinite(a).all(), (name, a.shape)
    # every map is normalised to [0, 1] by its reference recipe
    assert 0.0 <= float(load("occlusion/1/scan0__heatmap.npy").min()) and float(load("occlusion/1/scan0__heatmap.npy").max()) <= 1.0


def test_module_forward_backward_does_not_leak():
    """CTCLIP.forward + sim.backward() in a loop (the bench's end-to-end step, the reference's IG / Grad-CAM call
    pattern visualizations.py:862-872): device memory must be flat from the second step on WITHOUT relying on the
    cyclic garbage collector, and a second backward through a released forward must fail loudly."""
    import gc
    from models.ctclip import CTCLIP
    from utils.ctvit import CTViT
    cfg = O.TINY
    torch.manual_seed(0)
    vit = CTViT(dim=cfg.dim, codebook_size=cfg.codebook_size, image_size=cfg.image_size, patch_size=cfg.patch_size,
                temporal_patch_size=cfg.temporal_patch_size, spatial_depth=cfg.spatial_depth,
                temporal_depth=cfg.temporal_depth, dim_head=cfg.dim_head, heads=cfg.heads)
    H = cfg.image_size // cfg.patch_size
    clip = CTCLIP(text_encoder=torch.nn.Identity(), image_encoder=vit, dim_text=cfg.dim_text,
                  dim_image=H * H * cfg.dim, dim_latent=cfg.dim_latent).cuda()
    vol = O.synthetic_volume(cfg, 0, batch=2).cuda()
    text = O.synthetic_text_embeds(cfg, 7).cuda()
    gc.collect()
    gc.disable()
    try:
        used, g0, same = [], None, []
        for _ in range(5):
            x = vol.clone().requires_grad_()
            sim, *_ = clip(None, x, text)
            sim[:, 0].sum().backward()
            if g0 is None:
                g0 = x.grad.clone()
            same.append(bool(torch.equal(g0, x.grad)))
            del x, sim, _
            torch.cuda.synchronize()
            used.append(torch.cuda.memory_allocated())
    finally:
        gc.enable()
    assert used[1] == used[2] == used[3] == used[4], used
    assert all(same)
    x = vol.clone().requires_grad_()
    sim, *_ = clip(None, x, text)
    sim[:, 0].sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="released"):
        sim[:, 0].sum().backward()


def test_bad_inputs_fail_loudly():
    """No CPU path and no silent reshaping: CPU tensors, wrong ranks and volumes that do not tile into patches are
    rejected with an exception before any kernel is launched."""
    from models.ctclip import CTCLIP
    from utils.ctvit import CTViT
    cfg = O.TINY
    vit = CTViT(dim=cfg.dim, codebook_size=cfg.codebook_size, image_size=cfg.image_size, patch_size=cfg.patch_size,
                temporal_patch_size=cfg.temporal_patch_size, spatial_depth=cfg.spatial_depth,
                temporal_depth=cfg.temporal_depth, dim_head=cfg.dim_head, heads=cfg.heads)
    H = cfg.image_size // cfg.patch_size
    clip = CTCLIP(text_encoder=torch.nn.Identity(), image_encoder=vit, dim_text=cfg.dim_text,
                  dim_image=H * H * cfg.dim, dim_latent=cfg.dim_latent)
    vol = O.synthetic_volume(cfg, 0)
    text = O.synthetic_text_embeds(cfg, 7)
    with pytest.raises(RuntimeError, match="CUDA"):
        clip(None, vol, text)                                   # model and volume on the CPU
    with pytest.raises(RuntimeError, match="CUDA"):
        vit(vol)
    clip = clip.cuda()
    eng = clip.engine()
    tl = eng.text_latents(text.cuda())
    with pytest.raises(RuntimeError, match="CUDA"):
        eng.forward(vol, tl)                                    # CPU volume into a CUDA engine
    with pytest.raises(ValueError, match="shape"):
        eng.forward(vol.cuda()[0], tl)                          # missing batch axis
    with pytest.raises(ValueError, match="tile"):
        eng.forward(vol.cuda()[:, :, :-1].contiguous(), tl)     # depth not a multiple of the temporal patch
    with pytest.raises(ValueError, match="tile"):
        eng.forward(vol.cuda()[..., :-4].contiguous(), tl)      # wrong in-plane size
    with pytest.raises(ValueError, match="volumes for a batch"):
        eng.forward(torch.cat([vol, vol]).cuda(), tl, batch=3)
    sim = clip(None, vol.cuda(), text.cuda())[0]                # and the good call still works
    assert sim.shape == (1, 1) and torch.isfinite(sim).all()
