"""Zero-shot scoring (reference src/utils/CTClipInference.py:133-201).

CPU: the oracle restatement against the fixture produced by the reference's own `zeroshot` loop
(tests/golden/make_golden_zeroshot.py), and the prompt order of the host mirror.
GPU: the pair-softmax kernel against the same fixture, the engine path against the oracle at the TINY config (one
image forward with Bt = 2P against the reference's P forwards), and the `CTClipInference.zeroshot()` entry point."""
import numpy as np
import pytest
import torch

from oracle import ctclip_oracle as O

REF_PATHOLOGIES = 18


def load_golden(golden_dir):
    g = np.load(golden_dir / "zero_shot.npz")
    return (torch.from_numpy(g["image_latents"]), torch.from_numpy(g["pair_latents"]), torch.from_numpy(g["temp"]),
            g["predictions"])


def test_oracle_matches_reference_loop(golden_dir):
    il, pair, temp, pred = load_golden(golden_dir)
    assert pred.shape == (il.shape[0], REF_PATHOLOGIES) and pred.dtype == np.float64
    for i in range(il.shape[0]):
        got = O.zero_shot_predictions(il[i:i + 1], pair, temp).numpy()
        assert np.abs(got - pred[i]).max() <= 1e-7, i
    # rank indexing of the gathered diagonal (:175-176): sample r of a gathered batch scores against ITS prompts
    both = O.zero_shot_predictions(il[:2], pair[:, [0, 1, 0, 1]], temp, rank=1).numpy()
    assert np.abs(both - pred[1]).max() <= 1e-7


def test_prompt_order_and_shim_surface():
    from ctclip_b200.attribution import PATHOLOGIES
    from ctclip_b200.zeroshot import zero_shot_prompts
    assert len(PATHOLOGIES) == REF_PATHOLOGIES
    p = zero_shot_prompts()
    assert len(p) == 2 * REF_PATHOLOGIES
    assert p[0] == "There is Medical material." and p[1] == "There is no Medical material."
    assert p[20] == "There is Lung opacity." and p[21] == "There is no Lung opacity."
    assert zero_shot_prompts([]) == []
    from utils.CTClipInference import CTClipInference
    assert callable(CTClipInference.zeroshot)


@pytest.mark.gpu
def test_pair_softmax_kernel(golden_dir):
    from ctclip_b200._lib import call, stream_ptr
    il, pair, temp, pred = load_golden(golden_dir)
    n, P = pred.shape
    # logits formed exactly as validate_prompts forms them (one [1,d] x [d,1] product each), so the kernel sees the
    # reference's own fp32 inputs
    sim = torch.tensor([[float(il[i:i + 1] @ pair[j, k:k + 1].t() * temp) for j in range(P) for k in (0, 1)]
                        for i in range(n)]).cuda().contiguous()
    out = torch.empty(n, P, dtype=torch.float64, device="cuda")
    call("ctc_pair_softmax", sim, n, P, out, stream_ptr())
    assert np.abs(out.cpu().numpy() - pred).max() <= 2e-7           # 2 ulp of the reference's fp32 softmax
    # saturated, equal and ragged inputs
    edge = torch.tensor([[80.0, -80.0, -80.0, 80.0, 3.0, 3.0, 1e4, 1e4 - 1]], device="cuda")
    o = torch.empty(1, 4, dtype=torch.float64, device="cuda")
    call("ctc_pair_softmax", edge, 1, 4, o, stream_ptr())
    ref = torch.softmax(edge.view(4, 2), dim=1)[:, 0].double()
    assert torch.equal(o.view(-1) == 1, ref == 1) and float((o.view(-1) - ref).abs().max()) < 1e-7
    call("ctc_pair_softmax", edge, 0, 4, o, stream_ptr())            # empty batch: no launch, no error
    call("ctc_pair_softmax", edge, 1, 0, o, stream_ptr())


def tiny_clip():
    from test_gpu_dropin import FakeTextTower
    from models.ctclip import CTCLIP
    from utils.ctvit import CTViT
    cfg = O.TINY
    torch.manual_seed(0)
    vit = CTViT(dim=cfg.dim, codebook_size=cfg.codebook_size, image_size=cfg.image_size, patch_size=cfg.patch_size,
                temporal_patch_size=cfg.temporal_patch_size, spatial_depth=cfg.spatial_depth,
                temporal_depth=cfg.temporal_depth, dim_head=cfg.dim_head, heads=cfg.heads)
    H = cfg.image_size // cfg.patch_size
    return cfg, CTCLIP(text_encoder=FakeTextTower(cfg.dim_text), image_encoder=vit, dim_text=cfg.dim_text,
                       dim_image=H * H * cfg.dim, dim_latent=cfg.dim_latent).cuda()


@pytest.mark.gpu
def test_engine_zero_shot_vs_oracle():
    from ctclip_b200.zeroshot import zero_shot_probabilities
    from test_gpu_model import make_engine
    cfg = O.TINY
    eng, sd = make_engine(cfg)
    P, B = 5, 2
    vol = O.synthetic_volume(cfg, 3, batch=B).cuda()
    # prompt pairs that share a direction, so the probabilities are not saturated
    g = torch.Generator().manual_seed(5)
    emb = (torch.randn(P, 1, cfg.dim_text, generator=g) + 0.3 * torch.randn(P, 2, cfg.dim_text, generator=g))
    emb = emb.reshape(2 * P, cfg.dim_text).cuda()
    tl = eng.text_latents(emb)
    got = zero_shot_probabilities(eng, vol, tl)
    assert got.shape == (B, P) and got.dtype == torch.float64
    ctx = eng.forward(vol, tl, save=False)                           # same codes for the oracle (hard arg-max)
    with torch.no_grad():
        _, il, tlr, temp, _, _ = O.ctclip_forward(vol, emb, sd, cfg, None, force_indices=ctx.indices)
    for b in range(B):
        ref = O.zero_shot_predictions(il[b:b + 1].cpu(), tlr.view(P, 2, -1).cpu(), temp.cpu())
        err = float((got[b].cpu() - ref).abs().max())
        print(f"\n[zero-shot] sample {b}: p = {got[b].cpu().numpy().round(4)}  max |dp| {err:.2e}")
        assert err < 2e-3                                            # |d logit| < 2e-3  =>  |dp| <= 1e-3
        assert float(got[b].min()) > 0.01 and float(got[b].max()) < 0.99
    assert zero_shot_probabilities(eng, vol, tl[:0]).shape == (B, 0)
    with pytest.raises(ValueError):
        zero_shot_probabilities(eng, vol, tl[:3])


@pytest.mark.gpu
def test_inference_zeroshot_entry_point(tmp_path):
    from test_gpu_dropin import FakeTokenizer
    from utils.CTClipInference import CTClipInference
    from ctclip_b200.zeroshot import pair_text_latents, zero_shot_probabilities
    cfg, clip = tiny_clip()
    vols = [O.synthetic_volume(cfg, i) for i in range(3)]
    labels = [(torch.arange(18) % (i + 2) == 0).float() for i in range(3)]
    loader = [(v, ["report"], [l], ["scan"], ["scan.nii.gz"]) for v, l in zip(vols, labels)]
    inf = CTClipInference(clip, batch_size=1, dataset=None, dataloader=loader, tokenizer=FakeTokenizer(),
                          results_folder=tmp_path, zero_shot=True, visualize=False)
    inf.infer()
    pred = np.load(inf.results_folder / "zero_shot_predictions.npy")
    targ = np.load(inf.results_folder / "zero_shot_targets.npy")
    assert pred.shape == (3, 18) and pred.dtype == np.float64 and targ.shape == (3, 18)
    assert np.array_equal(targ, torch.stack(labels).numpy())
    assert ((pred > 0) & (pred < 1)).all()
    # batched scoring == the per-volume loop, bit for bit (rows of a batch are independent)
    tl = pair_text_latents(clip, FakeTokenizer(), device=torch.device("cuda"))
    batched = zero_shot_probabilities(clip.engine(), torch.cat(vols).cuda(), tl).cpu().numpy()
    assert np.array_equal(batched, pred)
