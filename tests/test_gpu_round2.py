"""GPU: kernels and host paths added in round 2 — head-fused probability emission, the generic device
attention_rollout (every option of visualizations.py:707-743), the synchronisation-free IG quantile chain, the
deterministic reductions (IG batch sum, Grad-CAM channel weights, rollout), run-to-run bit reproducibility."""
import numpy as np
import pytest
import torch

from oracle import ctclip_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


@pytest.fixture(scope="module")
def setup():
    from ctclip_b200.engine import Engine
    from ctclip_b200.plan import Config, Plan
    torch.backends.cuda.matmul.allow_tf32 = False
    eng = Engine(Plan(O.init_state_dict(O.FULL, 42), Config(), DEV))
    vol = O.synthetic_volume(O.FULL, 0).to(DEV)
    tl = eng.text_latents(O.synthetic_text_embeds(O.FULL, 7).to(DEV))
    return eng, vol, tl


@pytest.mark.parametrize("n,heads", [(24, 8), (576, 8), (100, 3)])
@pytest.mark.parametrize("fusion", ["mean", "max"])
@pytest.mark.parametrize("discard,residual", [(0.0, True), (0.3, True), (0.9, False), (0.5, False)])
def test_generic_attention_rollout_matches_reference_restatement(n, heads, fusion, discard, residual):
    """Visualizations.attention_rollout with head_fusion / discard_ratio / use_residual on the device kernels vs the
    oracle's restatement of visualizations.py:720-741 (pinned to the real method by the CPU golden tests)."""
    from ctclip_b200 import attribution as A
    g = torch.Generator().manual_seed(n + heads)
    layers = [torch.softmax(3 * torch.randn(heads, n, n, generator=g), dim=-1).to(DEV) for _ in range(3)]
    ours = A.attention_rollout(layers, head_fusion=fusion, discard_ratio=discard, use_residual=residual)
    ref = O.attention_rollout(layers, fusion, discard, residual)
    err = float((ours - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err
    with pytest.raises(ValueError):
        A.attention_rollout(layers, head_fusion="median")


def test_fused_probability_emission_equals_reductions_of_materialised_probs(setup):
    """ctc_attention_fused_probs (head mean / head max / per-head query mean without materialising attn[b,h,n,n])
    vs the same reductions of ctc_attention_probs."""
    eng, vol, tl = setup
    ctx = eng.forward(vol, tl, keep_attn=True)
    for kind, layer in (("spatial", 0), ("spatial", 3), ("temporal", 1)):
        probs = eng.attention_probs(ctx, kind, layer)                 # [n_seq, heads, n, n]
        fm, cm = eng.attention_fused(ctx, kind, layer, fused=True, colmean=True)
        fx, _ = eng.attention_fused(ctx, kind, layer, fused=True, fusion="max")
        assert float((fm - probs.mean(dim=1)).abs().max()) < 2e-6
        assert float((fx - probs.max(dim=1).values).abs().max()) < 2e-6
        assert float((cm - probs.mean(dim=2)).abs().max()) < 2e-6
        assert float((fm.sum(-1) - 1).abs().max()) < 2e-3             # rows of the head mean sum to one
        fm2, cm2 = eng.attention_fused(ctx, kind, layer, fused=True, colmean=True)
        assert torch.equal(fm, fm2) and torch.equal(cm, cm2)          # no atomics: bit-reproducible
        del probs


def test_device_quantile_chain_is_numpy_exact():
    from ctclip_b200 import attribution as A
    g = torch.Generator().manual_seed(1)
    x = torch.rand(96, 100, 120, generator=g).pow(3)
    x[x < 0.2] = 0                                    # many exact zeros, like a relu'd IG map
    xd = x.to(DEV)
    for q in (0.9, 0.5, 0.999, 0.0, 1.0):
        assert float(A.quantile_linear_dev(xd, q)) == float(np.quantile(x.numpy(), q)), q
    srt = np.sort(x.numpy().ravel())
    for k in (0, 1, 12345, x.numel() // 2, x.numel() - 1):
        assert float(A.kth_value_dev(xd, k)) == float(srt[k])
    assert A.quantile_linear(xd, 0.9) == float(A.quantile_linear_dev(xd, 0.9))


def test_batch_sum_and_colmean_are_ordered_sums():
    from ctclip_b200 import _lib
    from ctclip_b200._lib import call, stream_ptr
    g = torch.Generator().manual_seed(2)
    src = torch.randn(7, 4096 * 3, generator=g).to(DEV)
    dst = torch.randn(4096 * 3, generator=g).to(DEV)
    ref = dst.clone()
    for b in range(7):
        ref = ref + src[b] * 0.5
    call("ctc_batch_sum", src, 7, src.shape[1], 0.5, 1, dst, stream_ptr())
    assert torch.equal(dst, ref)                                      # same order of additions -> same bits
    x = torch.randn(13824, 512, generator=g).to(DEV)
    w = torch.empty(512, device=DEV)
    ws = torch.empty(_lib.load().ctc_colmean_ws_floats(13824, 512), device=DEV)
    call("ctc_colmean", x, 13824, 512, w, ws, stream_ptr())
    w2 = torch.empty(512, device=DEV)
    call("ctc_colmean", x, 13824, 512, w2, ws, stream_ptr())
    assert torch.equal(w, w2)
    assert float((w - x.double().mean(0).float()).abs().max()) < 1e-6


def test_attribution_is_bit_reproducible_run_to_run(setup):
    """IG with batch > 1 (sum over the batch rows), Grad-CAM weights and rollout used floating-point atomics.
    All reductions are ordered now: two runs give identical bits."""
    from ctclip_b200 import attribution as A
    eng, vol, tl = setup
    a, aux_a = A.integrated_gradients(eng, vol, tl, steps=4, batch=4, shard_steps=False)
    b, aux_b = A.integrated_gradients(eng, vol, tl, steps=4, batch=4, shard_steps=False)
    assert torch.equal(aux_a["gsum"], aux_b["gsum"]) and torch.equal(a, b)
    c, aux_c = A.integrated_gradients(eng, vol, tl, steps=4, batch=2, shard_steps=False)
    assert float((aux_c["gsum"] - aux_a["gsum"]).abs().max()) <= 1e-5 * float(aux_a["gsum"].abs().max())
    m1, m2 = A.grad_cam(eng, vol, tl), A.grad_cam(eng, vol, tl)
    assert all(torch.equal(m1[k], m2[k]) for k in m1)
    r1, r2 = A.attention_rollout_maps(eng, vol, tl), A.attention_rollout_maps(eng, vol, tl)
    assert torch.equal(r1[0], r2[0]) and torch.equal(r1[1], r2[1])


def test_to_host_returns_owned_copies():
    from ctclip_b200 import attribution as A
    a = torch.full((8, 16), 1.0, device=DEV)
    b = torch.full((8, 16), 2.0, device=DEV)
    ha = A.to_host(a)
    hb = A.to_host(b)                                                 # same size / dtype / slot: same staging buffer
    assert float(ha.mean()) == 1.0 and float(hb.mean()) == 2.0
    va = A.to_host(a, view=True)
    A.to_host(b, view=True)
    assert float(va.mean()) == 2.0                                    # the documented aliasing of view=True
