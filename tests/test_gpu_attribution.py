"""GPU parity of the attribution methods against (a) the golden fixtures produced by the UNMODIFIED
reference `Visualizations` class (tests/golden/full_attrib.npz) and (b) the oracle on the same device."""
import math

import numpy as np
import pytest
import torch

from oracle import ctclip_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


def pearson(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    a, b = a - a.mean(), b - b.mean()
    return float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))


@pytest.fixture(scope="module")
def setup(golden_dir):
    from ctclip_b200.engine import Engine
    from ctclip_b200.plan import Config, Plan
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = O.init_state_dict(O.FULL, 42)
    eng = Engine(Plan(sd, Config(), DEV))
    vol = O.synthetic_volume(O.FULL, 0).to(DEV)
    tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(DEV))
    return eng, vol, tl, np.load(golden_dir / "full_attrib.npz"), O.to_device(sd, DEV)


def test_grad_cam_vs_reference_golden(setup):
    from ctclip_b200 import attribution as A
    eng, vol, tl, gold, _ = setup
    maps = A.grad_cam(eng, vol, tl)
    for k in ("spatial", "spatial_ff", "temporal", "temporal_ff", "combined", "vq"):
        ref = gold["gradcam_" + k]
        r = pearson(maps[k], ref)
        err = float(np.abs(maps[k].cpu().numpy() - ref).max())
        print(f"[grad-cam {k}] pearson {r:.5f} max abs err {err:.3e}")
        # random-codebook checkpoint: the VQ-CAM is relu(codebook rows . w), i.e. it moves with every flipped near-tie
        # code (0.7 % of the tokens here); the north-star 0.999 is asserted on the fitted-codebook checkpoint in
        # tests/test_gpu_parity.py, where the arg-max is as well conditioned as in a trained model
        assert r > (0.99 if k in ("vq", "combined", "spatial") else 0.999), (k, r)


def test_rollout_and_raw_attention_vs_reference_golden(setup):
    from ctclip_b200 import attribution as A
    eng, vol, tl, gold, _ = setup
    sp, tp = A.attention_rollout_maps(eng, vol, tl)
    assert tuple(sp.shape) == (96, 24, 24) and tuple(tp.shape) == (24, 24, 24)
    for name, mine, ref in (("spatial", sp, gold["rollout_spatial"]), ("temporal", tp, gold["rollout_temporal"])):
        r = pearson(mine, ref)
        err = float(np.abs(mine.cpu().numpy() - ref).max())
        print(f"[rollout {name}] pearson {r:.6f} max abs err {err:.3e}")
        assert r > 0.999 and err < 2e-2
    rs, rt = A.raw_attention_maps(eng, vol, tl)            # [heads, layers, D, H, W], normalised
    gs = torch.from_numpy(gold["rawattn_spatial"])         # [heads, layers, 24, 576] raw query-means
    gt = torch.from_numpy(gold["rawattn_temporal"])        # [heads, layers, 576, 24]
    for h in range(8):
        for l in range(4):
            a = O.norm_minmax_max(gs[h, l].reshape(24, 24, 24))
            b = O.norm_minmax_max(gt[h, l].reshape(24, 24, 24).permute(2, 0, 1))
            assert float((rs[h, l].cpu() - a).abs().max()) < 3e-2
            assert float((rt[h, l].cpu() - b).abs().max()) < 3e-2


def test_integrated_gradients_vs_reference_golden(setup):
    from ctclip_b200 import attribution as A
    eng, vol, tl, gold, _ = setup
    steps = int(gold["ig_steps"])
    out, aux = A.integrated_gradients(eng, vol, tl, steps=steps, batch=steps, shard_steps=False)
    pre = aux["pre_threshold"]
    sub = pre[::4, ::8, ::8].cpu().numpy()
    r = pearson(sub, gold["ig_pre_sub"])
    # the exact device quantile agrees with numpy on OUR map
    q_np = float(np.quantile(pre.cpu().numpy(), 0.90))
    aux["q90"] = float(aux["q90"])                 # a device scalar: the post-processing chain never syncs with the host
    print(f"[IG] pre-threshold pearson {r:.5f}; q90 ours {aux['q90']:.6e} numpy-on-ours {q_np:.6e} "
          f"reference {float(gold['ig_q90']):.6e}; nonzero {int((out > 0).sum())} vs {int(gold['ig_final_nonzero'])}")
    assert aux["q90"] == q_np                      # exact device quantile == numpy, bit for bit
    assert r > 0.999
    nz = int((out > 0).sum())
    assert abs(nz - int(gold["ig_final_nonzero"])) < 0.02 * int(gold["ig_final_nonzero"])
    # final map == reference post-processing applied to OUR pre-threshold map (bit-level formula check)
    ref_post = O.integrated_gradients_post(torch.relu((vol[0, 0] - 1) * (aux["gsum"] * (1.0 / steps))).cpu())
    assert float(np.abs(out.cpu().numpy() - ref_post).max()) < 1e-5


def test_occlusion_coarse_vs_reference_golden(setup):
    """Eight coarse windows.  Window list / masks are bit-exact; the logits are compared (a) with the oracle
    conditioned on OUR code assignments (tight) and (b) with the reference golden values (loose: a random-init
    model with ~30 % constant 'air' tokens flips near-tie VQ codes in a correlated way under any rounding
    change — two fp32 runs on different hardware already differ at this level, SURVEY §7 hard part 1)."""
    from ctclip_b200 import attribution as A
    eng, vol, tl, gold, sd = setup
    ps, st = tuple(int(x) for x in gold["occ_patch"]), tuple(int(x) for x in gold["occ_stride"])
    heat, aux = A.occlusion_sensitivity(eng, vol, tl, ps, st, batch=4)
    windows = aux["windows"]
    assert windows == O.occlusion_windows((240, 480, 480), ps, st)        # bit-exact window list
    scores = aux["scores"].cpu().numpy()
    print(f"[occlusion] orig {aux['orig']:.6f} vs {float(gold['occ_orig']):.6f}\n  ours {scores}\n  ref  {gold['occ_scores']}")
    assert abs(aux["orig"] - float(gold["occ_orig"])) < 2e-2
    assert np.abs(scores - gold["occ_scores"]).max() < 3e-2
    # (a) same code assignments -> tight agreement of every logit, fused cube mask == materialised mask
    txt = O.synthetic_text_embeds(O.FULL, 7).to(DEV)
    wins = torch.tensor([[d, h, w, *ps] for (d, h, w) in windows], dtype=torch.int32, device=DEV)
    ctx = eng.forward(vol, tl, batch=len(windows), occl=wins)
    ind = ctx.indices.view(len(windows), -1)
    agree = []
    with torch.no_grad():
        for i, win in enumerate(windows):
            occ = O.occlusion_mask_apply(vol, win, ps)
            s_forced = O.ctclip_forward(occ, txt, sd, O.FULL, None, force_indices=ind[i])[0]
            s_free, *_r, ind_o = O.ctclip_forward(occ, txt, sd, O.FULL)
            agree.append(float((ind_o.reshape(-1) == ind[i].long()).float().mean()))
            assert abs(float(s_forced) - float(ctx.sim[i, 0])) < 2e-3, (i, float(s_forced), float(ctx.sim[i, 0]))
    print(f"  VQ code agreement per window: {np.round(agree, 4)}")
    assert min(agree) > 0.9
    # heat map assembled from OUR scores equals the reference assembly of the same scores
    h64, c64 = O.occlusion_accumulate((240, 480, 480), windows, ps, aux["orig"], scores)
    ref = O.occlusion_finalize(h64, c64)
    assert float(np.abs(heat.cpu().numpy() - ref).max()) < 1e-5


def test_occlusion_reuse_equals_dense_forward(setup):
    """The frame-reuse fast path (Engine.forward_occluded) must give the same logits and the same VQ codes as
    a full forward with the cube fused into the patch-embedding load, for reference-sized windows
    (20,40,40) at the volume corners, the centre and the last frames (clipped changed-frame sets)."""
    from ctclip_b200 import attribution as A
    eng, vol, tl, _, _ = setup
    ps = (20, 40, 40)
    windows = [(0, 0, 0), (220, 440, 440), (100, 200, 220), (210, 0, 440), (10, 20, 20), (150, 440, 0), (220, 0, 0)]
    stats = {}
    o_fast, s_fast = A.occlusion_scores(eng, vol, tl, windows, ps, reuse=True, reuse_batch=4, stats=stats)
    o_dense, s_dense = A.occlusion_scores(eng, vol, tl, windows, ps, reuse=False, batch=4)
    # windows lying entirely in the -1 border of the synthetic volume are recognised as no-ops and not evaluated;
    # evaluating them anyway gives the same bits
    assert stats == {"evaluated": 2, "noop": 5}, stats
    _, s_all = A.occlusion_scores(eng, vol, tl, windows, ps, reuse=True, reuse_batch=4, skip_noop=False)
    assert torch.equal(s_all, s_fast)
    print(f"[occlusion reuse] orig {o_fast:.7f} / {o_dense:.7f}\n  fast  {s_fast.cpu().numpy()}\n  dense {s_dense.cpu().numpy()}")
    assert o_fast == o_dense
    assert float((s_fast - s_dense).abs().max()) < 1e-6
    cache = eng.occlusion_baseline(vol, tl)
    cubes = [[d // 10, h // 20, w // 20] for (d, h, w) in windows[:3]]
    ctx_f = eng.forward_occluded(cache, cubes, (2, 2, 2), tl)
    wins = torch.tensor([[d, h, w, *ps] for (d, h, w) in windows[:3]], dtype=torch.int32, device=DEV)
    ctx_d = eng.forward(vol, tl, batch=3, occl=wins)
    assert torch.equal(ctx_f.indices, ctx_d.indices)
    assert torch.equal(ctx_f.x_pre_vq, ctx_d.x_pre_vq)


def test_occlusion_multi_prompt_sweep_equals_single_prompt_sweeps(setup):
    """One sweep scoring P prompts == P independent sweeps (the reference's loop over positive
    pathologies, visualizations.py:1037-1044): bit-identical scores and heat maps."""
    from ctclip_b200 import attribution as A
    eng, vol, _, _, _ = setup
    emb = torch.randn(3, 768, generator=torch.Generator().manual_seed(11)).to(DEV)
    tl = eng.text_latents(emb)
    ps = st = (120, 240, 240)
    heats, aux = A.occlusion_sensitivity_multi(eng, vol, tl, ps, st)
    assert len(heats) == 3 and tuple(aux["scores"].shape) == (8, 3)
    for j in range(3):
        h1, a1 = A.occlusion_sensitivity(eng, vol, tl[j:j + 1].contiguous(), ps, st)
        assert a1["orig"] == float(aux["orig"][j])
        assert torch.equal(a1["scores"], aux["scores"][:, j])
        assert torch.equal(h1, heats[j])


def test_occlusion_heatmap_default_grid_vs_oracle():
    """12 167-window grid, random scores, shard with dropped remainder: bit-exact masks / counts."""
    from ctclip_b200 import attribution as A
    shape, ps, st = (240, 480, 480), (20, 40, 40), (10, 20, 20)
    windows = O.occlusion_windows(shape, ps, st)
    g = torch.Generator().manual_seed(3)
    scores = torch.rand(len(windows), generator=g) * 0.02
    orig = 0.012
    kept = len(windows) // 8 * 8
    inc = torch.zeros(len(windows), dtype=torch.uint8)
    inc[:kept] = 1
    heat = A.occlusion_heatmap(orig, scores.to(DEV), inc.to(DEV), shape, ps, st, rot90=True)
    h64, c64 = O.occlusion_accumulate(shape, windows[:kept], ps, orig, scores[:kept].numpy())
    ref = O.occlusion_finalize(h64, c64)
    assert heat.shape == ref.shape
    assert float(np.abs(heat.cpu().numpy() - ref).max()) < 2e-6


def test_quantile_minmax_normalize_kernels():
    from ctclip_b200 import attribution as A
    g = torch.Generator().manual_seed(1)
    x = torch.rand(96, 100, 120, generator=g).pow(3)
    x[x < 0.2] = 0                                    # many exact zeros, like relu'd IG
    xd = x.to(DEV)
    for q in (0.9, 0.5, 0.999):
        assert A.quantile_linear(xd, q) == float(np.quantile(x.numpy(), q))
    y = torch.randn(24, 24, 24, generator=g)
    for mode, fn in ((0, O.norm_minmax_max), (1, O.norm_minmax_range)):
        assert float((A.normalize(y.to(DEV), mode).cpu() - fn(y)).abs().max()) < 1e-6
    r = A.normalize(y.to(DEV), 2, rot90=True).cpu().numpy()
    assert np.abs(r - O.rot90((y / (y.max() + 1e-8)).numpy())).max() < 1e-6
    up = A.upsample(torch.rand(96, 24, 24, generator=g).to(DEV), (240, 480, 480), rot90=False)
    assert tuple(up.shape) == (240, 480, 480)


def test_module_api_autograd_matches_engine(setup):
    """CTCLIP.forward + sim.backward() (the call the reference's IG makes) == engine forward/backward."""
    from ctclip_b200.modules import CTCLIP, CTViT
    eng, vol, tl, gold, sd = setup
    vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
                spatial_depth=4, temporal_depth=4, dim_head=32, heads=8)
    clip = CTCLIP(text_encoder=torch.nn.Identity(), image_encoder=vit, dim_text=768, dim_image=294912, dim_latent=512)
    clip.load_state_dict(O.init_state_dict(O.FULL, 42), strict=False)
    x = vol.clone().requires_grad_()
    txt = O.synthetic_text_embeds(O.FULL, 7).to(DEV)
    sim, il, tlat, temp, tokens = clip(None, x, txt)
    assert tokens.shape == (1, 24, 24, 24, 512) and sim.shape == (1, 1)
    sim[0, 0].backward()
    ctx = eng.forward(vol, tl, save=True)
    g = eng.backward(ctx)
    assert float((sim.detach() - ctx.sim).abs().max()) < 1e-5
    assert float((x.grad - g).abs().max()) <= 1e-5 * float(g.abs().max())
    ids = vit(vol, return_only_codebook_ids=True)
    assert ids.shape == (1, 24, 24, 24) and torch.equal(ids.flatten().int(), ctx.indices)
